#!/bin/bash
# quick: training tests, default bench (train only), step timeline
set -u
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_sigmoid_fusion_gpu.py -m gpu -q -rf -k "not attention" > gpurun_out/pytest_${TAG}.log 2>&1; tail -5 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-eager --cpu-train-steps 0 --cpu-chunks 8 --skip-note-encoder > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"
timeout 300 python scripts/trace_step.py gpurun_out/trace_n1_${TAG}.csv > gpurun_out/trace_${TAG}.log 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print("==", {k: d.get(k) for k in ("value", "ms_per_step")}, d.get("roofline", {}).get("frac"), d.get("step_tensor", {}).get("frac_of_sustained_peak"), d.get("e2e", {}).get("value"))
print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
PY

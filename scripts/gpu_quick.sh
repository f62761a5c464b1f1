#!/bin/bash
# full -m gpu suite, default bench, config 3, step timeline
set -u
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_${TAG}.log 2>&1; tail -8 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-eager --cpu-train-steps 0 --cpu-chunks 8 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/bench_c3_${TAG}.json 2> gpurun_out/bench_c3_${TAG}.err; echo "config3 exit=$?"
timeout 600 python scripts/bench_attn_bwd.py > gpurun_out/attn_bwd_bench.log 2>&1; tail -16 gpurun_out/attn_bwd_bench.log
timeout 300 python scripts/trace_step.py gpurun_out/trace_n1_${TAG}.csv > gpurun_out/trace_${TAG}.log 2>&1
python - <<PY
import json
for f in ("bench_${TAG}", "bench_c3_${TAG}"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print("==", f, {k: d.get(k) for k in ("value", "ms_per_step")}, d.get("roofline", {}).get("frac"), d.get("step_tensor", {}).get("frac_of_sustained_peak"), d.get("e2e", {}).get("value"))
    print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
    n = d.get("note_encoder")
    if n:
        print({k: n.get(k) for k in ("value", "ms_per_step", "tensor_frac_of_sustained_peak")}, n["roofline"]["frac"])
        print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in n["kernels"].items()})
PY

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_kernels_gpu.py tests/test_note_encoder_gpu.py -k "attention or attn or golden or gemm" -m gpu -q -rf -x > gpurun_out/pytest_r02e.log 2>&1; tail -15 gpurun_out/pytest_r02e.log
timeout 600 python scripts/bench_attn_bwd.py > gpurun_out/attn_bwd_bench.log 2>&1; tail -16 gpurun_out/attn_bwd_bench.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-eager --cpu-train-steps 0 --cpu-chunks 8 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; echo "bench exit=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r02e.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step")}, d.get("roofline", {}).get("frac"), d.get("step_tensor", {}).get("frac_of_sustained_peak"))
print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
n = d["note_encoder"]; print({k: n.get(k) for k in ("value", "ms_per_step", "tensor_frac_of_sustained_peak")}, n["roofline"]["frac"])
print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in n["kernels"].items()})
PY

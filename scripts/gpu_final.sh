#!/bin/bash
# Round-end verification on one B200: smoke, full GPU test suite, the bench line (both arms), HBM kernel table.
set -u
TAG=${1:-r01d}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; echo "ref exit=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "tensor_frac_of_sustained_peak")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("clocks", d["clocks"])
print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
t = d.get("train"); print("train", t["value"], t["ms_per_step"], t["e2e"], t.get("cpu_baseline")); print("cpu", d.get("cpu_baseline"))
r = json.loads(open("gpurun_out/bench_ref_${TAG}.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["cpu_baseline"])
PY
timeout 300 python scripts/bench_hbm_kernels.py gpurun_out/hbm_kernels_${TAG}.json 2>&1 | grep -E "scaled|config 3" | cut -c1-150

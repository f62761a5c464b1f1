#!/bin/bash
# Round-2 trip A (one GPU): parity tests, the bench line (train headline + note encoder + eager baselines), config 3,
# the reference arm.  Outputs under gpurun_out/.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_${TAG}.log 2>&1; tail -8 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"
tail -3 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/bench_c3_${TAG}.json 2> gpurun_out/bench_c3_${TAG}.err; echo "config3 exit=$?"
tail -3 gpurun_out/bench_c3_${TAG}.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref exit=$?"
python - <<PY
import json
for f in ("bench_${TAG}", "bench_c3_${TAG}", "bench_ref_${TAG}"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print("==", f, {k: d.get(k) for k in ("metric", "value", "ms_per_step", "gpu_launches")})
    print("   e2e", d.get("e2e")); print("   roofline", d.get("roofline")); print("   step_tensor", d.get("step_tensor")); print("   clocks", d.get("clocks"))
    if "kernels" in d: print("   kernels", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
    n = d.get("note_encoder")
    if n:
        print("   note:", {k: n.get(k) for k in ("value", "ms_per_step", "tensor_frac_of_sustained_peak", "gpu_launches")}, n.get("e2e"), n.get("roofline"))
        if "kernels" in n: print("   note kernels", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in n["kernels"].items()})
        print("   note cpu", n.get("cpu_baseline"))
    print("   cpu", d.get("cpu_baseline")); print("   eager", d.get("torch_eager_b200"))
PY

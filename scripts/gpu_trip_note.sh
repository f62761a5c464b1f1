set -u
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_note_encoder_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python scripts/diag_kernels.py attn 2>&1 | grep time
timeout 600 python bench.py --cpu-chunks 0 --skip-train > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r01g.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "tensor_frac_of_sustained_peak")})
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"]); print("clocks", d["clocks"])
print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
PY

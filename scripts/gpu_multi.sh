#!/bin/bash
# Multi-GPU trip: data-parallel parity test, training-step bench and the config-5 sweep at N ranks.  usage: gpu_multi.sh N TAG
set -u
N=${1:-2}; TAG=${2:-r02d}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -s > gpurun_out/pytest_dp_${TAG}.log 2>&1; tail -25 gpurun_out/pytest_dp_${TAG}.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --skip-eager --cpu-train-steps 0 > gpurun_out/bench_n${N}_${TAG}.json 2> gpurun_out/bench_n${N}_${TAG}.err; echo "bench N=$N exit=$?"
tail -3 gpurun_out/bench_n${N}_${TAG}.err
timeout 900 $TR bench.py --gpus $N --config 5 --steps 3 --warmup 1 > gpurun_out/bench_c5_n${N}_${TAG}.json 2> gpurun_out/bench_c5_n${N}_${TAG}.err; echo "config5 N=$N exit=$?"
tail -3 gpurun_out/bench_c5_n${N}_${TAG}.err
python - <<PY
import json
for f in ("bench_n${N}_${TAG}", "bench_c5_n${N}_${TAG}"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print("==", f, {k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "gpu_launches", "scaling")})
    print("   e2e", d.get("e2e")); print("   config", d.get("config"))
    n = d.get("note_encoder")
    if n: print("   note:", {k: n.get(k) for k in ("value", "ms_per_step")})
PY

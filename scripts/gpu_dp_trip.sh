#!/bin/bash
# Multi-GPU trip for the data-parallel training step: parity (N ranks == one process on the concatenated batch),
# then the step time at N GPUs.  Usage: gpurun --gpus N -- bash scripts/gpu_dp_trip.sh N [check]
set -u
N=${1:-2}
mkdir -p gpurun_out
if [ "${2:-check}" = "check" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py 2>&1 | grep -E "rank|Error|error" | head -20
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --only-train 2>gpurun_out/train_n$N.err | tail -1 > gpurun_out/train_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/train_n$N.json").read())["train"]
print("N=$N train", round(d["value"], 1), "patients/s", round(d["ms_per_step"], 3), "ms/step  e2e", round(d["e2e"]["ms_per_step"], 3))
PY

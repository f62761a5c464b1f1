#!/bin/bash
# data-parallel training step against the SMs the lab tower's persistent kernels leave to the side stream + NCCL.  usage: N
N=${1:-2}
mkdir -p gpurun_out
for r in ${RS_LIST:-16 20 24 32}; do
  FAME_RESERVED_SMS_DP=$r timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 5 --skip-eager --cpu-train-steps 0 --skip-note-encoder > gpurun_out/sweep_rsdp_${N}_$r.json 2> gpurun_out/sweep_rsdp_${N}_$r.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/sweep_rsdp_${N}_$r.json").read().strip().splitlines()[-1])
print("N=$N reserved_sms_dp=$r", round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "patients/s")
PY
done

#!/bin/bash
# training-step time against the number of SMs the lab tower's persistent kernels leave to the demographic stream
mkdir -p gpurun_out
for r in 0 8 12 16 20 24 32; do
  FAME_RESERVED_SMS=$r timeout 300 python bench.py --steps 30 --warmup 5 --skip-eager --cpu-train-steps 0 --cpu-chunks 8 --skip-note-encoder > gpurun_out/sweep_rs_$r.json 2> gpurun_out/sweep_rs_$r.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/sweep_rs_$r.json").read().strip().splitlines()[-1])
print("reserved_sms=$r", round(d["ms_per_step"], 4), "ms/step", round(d["value"]), "patients/s; e2e", round(d["e2e"]["ms_per_step"], 4))
PY
done

#!/bin/bash
# One GPU trip: parity tests, the bench line, the ncu launch list of the same command, full captures of the
# dominant kernels, and the launch list of one training step.  Outputs under gpurun_out/ (summaries copied to
# profiles/ by scripts/summarize_profiles.py after reading them).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "tensor_frac_of_sustained_peak")})
print("e2e", d["e2e"]); print("roofline", d["roofline"]); print("clocks", d["clocks"])
print({k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()})
print("train", d.get("train")); print("cpu", d.get("cpu_baseline"))
PY
CMD="python bench.py --steps 2 --warmup 3 --skip-train --cpu-chunks 0"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "ncu launches exit=$?"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pair -s 12 -c 2 -o gpurun_out/prof_attn_${TAG} -f $CMD > gpurun_out/ncu_attn_${TAG}.log 2>&1
echo "ncu attn exit=$?"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 48 -c 4 -o gpurun_out/prof_gemm_${TAG} -f $CMD > gpurun_out/ncu_gemm_${TAG}.log 2>&1
echo "ncu gemm exit=$?"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:layernorm_bf16_rows2 -s 24 -c 1 -o gpurun_out/prof_ln_${TAG} -f $CMD > gpurun_out/ncu_ln_${TAG}.log 2>&1
echo "ncu ln exit=$?"
export FAME_NO_GRAPH=1
TCMD="python scripts/bench_train.py 32 542 2"
$TCMD > gpurun_out/plain_train_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train_${TAG}.csv $TCMD > gpurun_out/ncu_train_${TAG}.log 2>&1
echo "ncu train exit=$?"
unset FAME_NO_GRAPH
ls -la gpurun_out/*${TAG}*

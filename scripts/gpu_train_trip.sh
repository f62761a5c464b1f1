#!/bin/bash
# GPU trip for the training step: parity tests, step time (CUDA graph), per-kernel device times (ncu launch list).
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 300 python scripts/bench_train.py 32 542 20 2>&1 | tail -40
timeout 300 python scripts/bench_train.py 1024 542 3 2>&1 | head -12
export FAME_NO_GRAPH=1
CMD="python scripts/bench_train.py 32 542 2"
$CMD > gpurun_out/plain_train_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train_${TAG}.csv $CMD > gpurun_out/ncu_train_${TAG}.log 2>&1
echo "ncu exit=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_train_${TAG}.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
# keep the LAST step only: find the last clip_adamw launch and walk back to the previous one
names=[r[ix["Kernel Name"]] for r in rows[1:]]
ends=[i for i,n in enumerate(names) if "clip_adamw" in n]
lo, hi = ends[-2]+1, ends[-1]+1
agg=collections.OrderedDict()
for r in rows[1+lo:1+hi]:
    k=r[ix["Kernel Name"]][:70]; d=agg.setdefault(k,[0,0.0]); d[0]+=1; d[1]+=float(r[ix["Metric Value"]])
tot=sum(v[1] for v in agg.values())
print(f"one step: {hi-lo} launches, sum of device time {tot/1e6:.3f} ms")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{v[1]/1e3:9.1f} us x{v[0]:4d}  avg {v[1]/v[0]/1e3:7.1f} us  {k}")
PY

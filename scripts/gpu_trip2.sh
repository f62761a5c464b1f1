#!/bin/bash
# tests + smoke + bench + ncu launch list + ncu full capture of the GEMM
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit=$?"; cat gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cat gpurun_out/bench_ref.json
timeout 300 python bench.py --steps 2 --warmup 3 --cpu-chunks 0 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-chunks 0 > gpurun_out/ncu1.log 2>&1; echo "ncu1 exit=$?"
timeout 300 python bench.py --steps 1 --warmup 3 --cpu-chunks 0 > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 30 -c 4 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --cpu-chunks 0 > gpurun_out/ncu2.log 2>&1; echo "ncu2 exit=$?"
tail -3 gpurun_out/ncu2.log

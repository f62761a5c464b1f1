#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -k "attention" -m gpu -q -rf -x > gpurun_out/pytest_r02e.log 2>&1; tail -15 gpurun_out/pytest_r02e.log
timeout 600 python scripts/bench_attn_bwd.py > gpurun_out/attn_bwd_bench.log 2>&1; tail -14 gpurun_out/attn_bwd_bench.log

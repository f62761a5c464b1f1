#!/bin/bash
# Round-2 trip B (one GPU): full -m gpu suite with the complete log, then the default bench line.
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_${TAG}.log 2>&1; tail -12 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"
tail -3 gpurun_out/bench_${TAG}.err

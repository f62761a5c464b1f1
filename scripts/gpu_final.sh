#!/bin/bash
# Final trip of a round (one GPU): the full -m gpu suite, the default bench line, the reference arm, configs 3 and 5, the
# HBM kernel table, the ncu launch lists of the bench commands and full captures of the dominant kernels.
# Outputs under gpurun_out/; scripts/summarize_profiles.py turns them into the tracked summaries under profiles/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_${TAG}.log 2>&1; tail -4 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit=$?"; tail -2 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref exit=$?"
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/bench_c3_${TAG}.json 2> gpurun_out/bench_c3_${TAG}.err; echo "config3 exit=$?"
timeout 900 python bench.py --config 5 --steps 2 --warmup 1 > gpurun_out/bench_c5_${TAG}.json 2> gpurun_out/bench_c5_${TAG}.err; echo "config5 exit=$?"
timeout 600 python scripts/bench_hbm_kernels.py gpurun_out/hbm_kernels_${TAG}.json > gpurun_out/hbm_kernels_${TAG}.log 2>&1; echo "hbm exit=$?"; tail -12 gpurun_out/hbm_kernels_${TAG}.log
timeout 300 python scripts/bench_lab_small_kernels.py gpurun_out/lab_small_kernels_${TAG}.json > gpurun_out/lab_small_kernels_${TAG}.log 2>&1; echo "small kernels exit=$?"
# ---- ncu: launch lists (same commands, after each exited 0 without ncu), then full captures of the dominant kernels
TCMD="python bench.py --steps 2 --warmup 3 --skip-note-encoder --skip-eager --cpu-train-steps 0"
$TCMD > gpurun_out/plain_train_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train_${TAG}.csv $TCMD > gpurun_out/ncu_train_${TAG}.log 2>&1
echo "ncu train launches exit=$?"
$TCMD > gpurun_out/plain_train_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 72 -c 3 -o gpurun_out/prof_gemm_train_${TAG} -f $TCMD > gpurun_out/ncu_gemm_train_${TAG}.log 2>&1
echo "ncu train gemm exit=$?"
$TCMD > gpurun_out/plain_train_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 6 -c 2 -o gpurun_out/prof_attn_bwd_${TAG} -f $TCMD > gpurun_out/ncu_attn_bwd_${TAG}.log 2>&1
echo "ncu attn bwd exit=$?"
NCMD="python bench.py --config 2 --steps 2 --warmup 3 --cpu-chunks 0"
$NCMD > gpurun_out/plain_note_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_note_${TAG}.csv $NCMD > gpurun_out/ncu_note_${TAG}.log 2>&1
echo "ncu note launches exit=$?"
$NCMD > gpurun_out/plain_note_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 48 -c 2 -o gpurun_out/prof_gemm_note_${TAG} -f $NCMD > gpurun_out/ncu_gemm_note_${TAG}.log 2>&1
echo "ncu note gemm exit=$?"
$NCMD > gpurun_out/plain_note_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pair -s 12 -c 2 -o gpurun_out/prof_attn_fwd_${TAG} -f $NCMD > gpurun_out/ncu_attn_fwd_${TAG}.log 2>&1
echo "ncu note attn exit=$?"
du -sh gpurun_out; ls -la gpurun_out/*${TAG}* | head -40

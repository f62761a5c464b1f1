#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 2 -c 2 -f -o gpurun_out/prof_attn_bwd_fused_r02 python scripts/prof_attn_bwd.py > gpurun_out/ncu_attn_bwd_r02.log 2>&1; tail -5 gpurun_out/ncu_attn_bwd_r02.log

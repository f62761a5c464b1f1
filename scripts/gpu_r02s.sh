#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fame_model_gpu.py tests/test_metric_wrappers_gpu.py tests/test_eddi_fusion_gpu.py -m gpu -q -rf > gpurun_out/pytest_r02s.log 2>&1; tail -4 gpurun_out/pytest_r02s.log
timeout 600 python scripts/bench_hbm_kernels.py gpurun_out/hbm_kernels_r02s.json > gpurun_out/hbm_kernels_r02s.log 2>&1; grep -E "eval_counts|loss_stats" gpurun_out/hbm_kernels_r02s.log

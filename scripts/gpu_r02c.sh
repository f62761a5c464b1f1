#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eddi_fusion_gpu.py tests/test_average_fusion_gpu.py "tests/test_train_gpu.py::test_full_step_parity_at_bench_shape" tests/test_train_gpu.py::test_train_step_with_dropout -m gpu -q -rf -s > gpurun_out/pytest_r02c.log 2>&1; tail -12 gpurun_out/pytest_r02c.log
timeout 600 python scripts/diag_dropout_grad.py > gpurun_out/diag_dropout.log 2>&1; tail -20 gpurun_out/diag_dropout.log

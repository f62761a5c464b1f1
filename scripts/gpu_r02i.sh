#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_fame_model_gpu.py tests/test_metric_wrappers_gpu.py -k "attention or rank or metric or evaluate or calibrate" -m gpu -q -rf > gpurun_out/pytest_r02i.log 2>&1; tail -8 gpurun_out/pytest_r02i.log

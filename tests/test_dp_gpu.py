"""Data-parallel parity on real GPUs (needs >= 2 B200s on the box; skipped otherwise): scripts/dp_check.py under torchrun
at world size 2 -- N ranks == the single-process step on the concatenated batch (loss, bit-exact statistics, gradients,
the sharded optimizer's step + parameter sync, CUDA-graph replay with captured NCCL collectives, ablation 09)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_two_rank_step_equals_single_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu)")
    import __graft_entry__ as entry
    entry.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=540)
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and "FAIL" not in r.stdout and r.stdout.count("OK") >= 6

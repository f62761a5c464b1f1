"""Communication floor of the data-parallel step: NCCL SUM all-reduce of the gradient buckets alone (no compute),
timed on the device, max over ranks.  torchrun --nproc-per-node N scripts/nccl_ar_bench.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for mb, pieces in ((335, 1), (335, 8), (46, 1), (392, 1)):
    n = mb * 1000 * 1000 // 4
    buf = torch.ones(n, device=dev)
    cuts = [n * i // pieces for i in range(pieces + 1)]
    for _ in range(3):
        for a, b in zip(cuts, cuts[1:]):
            dist.all_reduce(buf[a:b])
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for a, b in zip(cuts, cuts[1:]):
            dist.all_reduce(buf[a:b])
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"all-reduce {mb} MB fp32 in {pieces} piece(s), {world} GPUs: {t.item():.3f} ms  "
              f"algbw {mb / t.item():.1f} GB/s  busbw {mb / t.item() * 2 * (world - 1) / world:.1f} GB/s", flush=True)
dist.barrier(); torch.cuda.synchronize()
parallel.shutdown(exit_code=0)

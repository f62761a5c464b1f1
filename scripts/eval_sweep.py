"""BASELINE configs[4] as a stand-alone run: fairmultimodal_b200.sweep.EvalSweep timed stage by stage (device time, max
over ranks).  bench.py --config 5 prints the same workload in the bench JSON contract.

    python scripts/eval_sweep.py [patients] [out.json]                      # one GPU (a per-GPU share of the cohort)
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/eval_sweep.py 46000
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import parallel, sweep  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD if world > 1 else None
P = int(sys.argv[1]) if len(sys.argv) > 1 else 5750
out_path = sys.argv[2] if len(sys.argv) > 2 else None

sw = sweep.EvalSweep(P, world, rank, dev, group)
sw.warm()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
res, ms = sw.run(resident=False)
t = torch.tensor([ms[k] for k in sorted(ms)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = dict(zip(sorted(ms), t.tolist()))
if rank == 0:
    out = {"workload": f"eval sweep: {P} patients x U{{1..16}} chunks = {sw.chunks_total} chunks of 512 tokens, {world} GPU(s), "
                       f"patients sharded by chunk count", "n_gpus": world, "patients": P, "chunks": sw.chunks_total,
           "ms": ms, "chunks_per_s": sw.chunks_total / (ms["note_encoder_and_pool"] * 1e-3),
           "patients_per_s_end_to_end": P / (ms["total"] * 1e-3), **res}
    print(json.dumps(out), flush=True)
    if out_path:
        json.dump(out, open(out_path, "w"), indent=1)
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
    parallel.shutdown(exit_code=0)

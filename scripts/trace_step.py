"""Device timeline of the training step (CUDA-graph replay) from torch.profiler / CUPTI: one CSV line per kernel with
start, duration and stream, to see what overlaps what (two tower streams, NCCL).  Works alone or under torchrun.
    python scripts/trace_step.py [out.csv]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, parallel, synth, train  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_step.csv"
B, L = 32, 542
torch.manual_seed(0)
model = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), dev).to(dev)
co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=1 + rank)
co["text"] = np.random.default_rng(rank).standard_normal((B, 768)).astype(np.float32)
keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features",
        "text", "labels")
batch = [torch.from_numpy(co[k]).to(dev) for k in keys]
pw = torch.from_numpy(synth.pos_weight(co["labels"])).to(dev)
hp = dict(lr=1e-5, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
group = dist.group.WORLD if world > 1 else None
model.train()
for _ in range(6):
    train.optimisation_step(model, batch, pw, 0.8, 0.01, (0.33, 0.33, 0.33), hp, group=group)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        train.optimisation_step(model, batch, pw, 0.8, 0.01, (0.33, 0.33, 0.33), hp, group=group)
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start if ev else 0
    with open(out, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for e in ev:
            f.write(f"{e.time_range.start - t0:.1f},{e.time_range.end - e.time_range.start:.1f},"
                    f"{getattr(e, 'stream', getattr(e, 'device_index', 0))},\"{e.name[:100]}\"\n")
    print(f"{len(ev)} device events -> {out}")
train.release_graphs(model)
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
    parallel.shutdown(exit_code=0)

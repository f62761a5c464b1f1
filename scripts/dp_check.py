"""Data-parallel parity check (run under torchrun on N GPUs):  N ranks x B patients must reproduce the single-process
step on the concatenated N*B batch: same global loss, same statistics (counts bit-exact), same gradients.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, ops, parallel, synth, train  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, L = 8, 40
keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features",
        "text", "labels")
co = synth.make_cohort(B * world, lab_tokens=L, chunks=0, with_tokens=False, seed=5)
co["text"] = (np.random.default_rng(3).standard_normal((B * world, 768)) * 0.5).astype(np.float32)
pw = torch.tensor([3.0, 1.2, 0.6], device=dev)
w = (0.41, 0.27, 0.32)


def build():
    torch.manual_seed(0)
    m = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.fame_shapes(lab_tokens=L), 4).items()}
    m.load_state_dict(sd)
    modules.set_dropout(m, 0.0)      # parity check: N ranks == one process is only defined without random masks
    return m.to(dev).train()


lo, hi = parallel.shard_range(B * world, rank, world)
shard = [torch.from_numpy(co[k][lo:hi]).to(dev) for k in keys]
full = [torch.from_numpy(co[k]).to(dev) for k in keys]

m_dp = build()
loss_dp, _ = train.forward_backward(m_dp, shard, pw, 0.8, 0.01, w, group=dist.group.WORLD)
st_dp = train.get_state(m_dp)
plan = st_dp.shard_plan(dist.group.WORLD)
g_dp = st_dp.g.clone()
m_1 = build()
loss_1, _ = train.forward_backward(m_1, full, pw, 0.8, 0.01, w, group=None)
g_1 = train.get_state(m_1).g.clone()
torch.cuda.synchronize()
dl = (loss_dp - loss_1).abs().max().item()
# with the sharded optimizer a rank holds the reduced gradient of ITS slice of every sharded bucket (reduce-scatter)
# and of the whole replicated tail: compare exactly those ranges with the single-process gradient
num = den = 0.0
for key in st_dp.grad_buckets():
    a, b = st_dp.my_range(key, plan)
    num += (g_dp[a:b] - g_1[a:b]).double().pow(2).sum().item()
    den += g_1[a:b].double().pow(2).sum().item()
rel = (num / den) ** 0.5
# one real optimizer step on both, then sync: Adam's first moment is linear in the clipped gradient
st_dp.clip_and_step(1e-4, 0.01)
train.get_state(m_1).clip_and_step(1e-4, 0.01)
train.sync_parameters(m_dp, dist.group.WORLD)
torch.cuda.synchronize()
st_1 = train.get_state(m_1)
rel_m = ((st_dp.m - st_1.m).norm() / st_1.m.norm()).item()
dnorm = abs(st_dp.grad_norm.item() - st_1.grad_norm.item()) / st_1.grad_norm.item()
moved = (st_dp.p - st_1.p).abs().max().item()           # both moved by ~lr per element from identical weights
# statistics: all-reduced shard statistics vs statistics of the whole batch
attrs_s = [shard[i] for i in (2, 4, 5)]
attrs_f = [full[i] for i in (2, 4, 5)]
z = torch.randn(B * world, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
st_s = ops.loss_stats(z[lo:hi].contiguous(), shard[8], attrs_s, pw)
dist.all_reduce(st_s)
st_f = ops.loss_stats(z, full[8], attrs_f, pw)
counts_equal = bool(torch.equal(st_s[78:103], st_f[78:103]))
sums_close = int((st_s[:78] - st_f[:78]).abs().max().item())
ok = dl < 1e-4 and rel < 2e-2 and counts_equal and sums_close == 0 and rel_m < 2e-2 and dnorm < 1e-3 and moved <= 2.5e-4
print(f"[rank {rank}] world={world} sharded={plan is not None} |loss_dp - loss_1|={dl:.2e} grad rel diff={rel:.3e} "
      f"adam-m rel diff after step+sync={rel_m:.3e} grad-norm rel diff={dnorm:.2e} max|p_dp - p_1|={moved:.2e} "
      f"counts_equal={counts_equal} fixed-point sum diff={sums_close} -> {'OK' if ok else 'FAIL'}", flush=True)
# ---- optimisation steps: CUDA graph with captured NCCL collectives (bucketed, overlapped gradient all-reduce) against
# eager steps, lr = 0 so that both see identical weights at every step
hp = dict(lr=0.0, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
res = {}
for mode in (True, False):
    m = build()
    st = train.get_state(m)
    losses = []
    for i in range(5):
        losses.append(train.optimisation_step(m, shard, pw, 0.8, 0.01, w, hp, group=dist.group.WORLD, use_graph=mode).clone())
    torch.cuda.synchronize()
    res[mode] = (torch.stack(losses), st.g.clone(), st.m.clone(),
                 any(e.get("graph") is not None for e in st.graphs.values()))
    train.release_graphs(m)
dl2 = (res[True][0] - res[False][0]).abs().max().item()
dg2 = ((res[True][1] - res[False][1]).norm() / res[False][1].norm()).item()
dm2 = ((res[True][2] - res[False][2]).norm() / res[False][2].norm()).item()
dl3 = (res[True][0][-1] - loss_1).abs().max().item()
ok2 = res[True][3] and dl2 < 1e-5 and dg2 < 1e-3 and dm2 < 1e-3 and dl3 < 1e-4
print(f"[rank {rank}] graph+NCCL vs eager: captured={res[True][3]} |dloss|={dl2:.2e} grad rel={dg2:.2e} adam-m rel={dm2:.2e} "
      f"|loss - single process|={dl3:.2e} -> {'OK' if ok2 else 'FAIL'}", flush=True)
ok = ok and ok2
# ---- ablation 09 (sigmoid_fusion): focal loss over the GLOBAL batch, gradients SUM-reduced == single process
from fairmultimodal_b200 import sigmoid_fusion as SF  # noqa: E402


def build_sf():
    torch.manual_seed(0)
    m = SF.MultimodalTransformer(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), dev)
    modules.set_dropout(m, 0.0)
    return m.to(dev).train()


m_a, m_b = build_sf(), build_sf()
m_b.load_state_dict(m_a.state_dict())
l_dp, _ = SF.forward_backward(m_a, shard[:8], shard[8], pw, gamma=1.0, group=dist.group.WORLD)
l_1, _ = SF.forward_backward(m_b, full[:8], full[8], pw, gamma=1.0, group=None)
sa, sb = SF.get_state(m_a), SF.get_state(m_b)
pl = sa.shard_plan(dist.group.WORLD)
num = den = 0.0
for key in sa.grad_buckets():
    a, b = sa.my_range(key, pl)
    num += (sa.g[a:b] - sb.g[a:b]).double().pow(2).sum().item()
    den += sb.g[a:b].double().pow(2).sum().item()
dl4, rel4 = abs(l_dp.item() - l_1.item()), (num / max(den, 1e-30)) ** 0.5
ok4 = dl4 < 1e-4 and rel4 < 2e-2
print(f"[rank {rank}] sigmoid-fusion (09) data parallel: |loss_dp - loss_1|={dl4:.2e} grad rel diff={rel4:.3e} -> {'OK' if ok4 else 'FAIL'}",
      flush=True)
ok = ok and ok4
sys.stdout.flush()
dist.barrier()
torch.cuda.synchronize()
parallel.shutdown(exit_code=0 if ok else 1)
sys.exit(0 if ok else 1)

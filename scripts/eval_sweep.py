"""BASELINE configs[4]: large-cohort evaluation sweep -- P patients x U{1..16} note chunks -> note encoder (256-chunk
batches) -> chunk->patient pool -> FAME model forward -> threshold calibration (101-point F1 sweep) -> AUROC / AUPRC /
F1 / EDDI / Equalized Odds over age / ethnicity / insurance subgroups.  Patients are sharded over the ranks by chunk
count (parallel.shard_patients_by_chunks); the only collectives are the int64 count all-reduce and the logit all-gather
for the exact rank metrics (parallel.evaluate_sharded).  Device time per stage, max over ranks.

    python scripts/eval_sweep.py [patients] [out.json]                      # one GPU (a per-GPU share of the cohort)
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/eval_sweep.py 46000
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import metrics, modules, ops, parallel, synth  # noqa: E402

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
L, CB, PB = 542, 256, 1024                       # lab tokens, chunks per encoder batch, patients per model batch

# ---- the cohort (every rank draws the same patient table; tokens only for its own shard)
meta = synth.make_cohort(P, lab_tokens=L, chunks="u1_16", with_tokens=False, seed=1234)
offs_all = meta["chunk_offsets"]
p_lo, p_hi = parallel.shard_patients_by_chunks(offs_all, world)[rank]
offs, (c_lo, c_hi) = parallel.rebase_offsets(offs_all, p_lo, p_hi)
C = c_hi - c_lo
rng = np.random.default_rng(99 + rank)
ids = rng.integers(1000, synth.VOCAB, (C, 512)).astype(np.int64)
n_per = np.diff(offs)
valid = np.full(C, 512, dtype=np.int64)
valid[offs[1:][n_per > 0] - 1] = rng.integers(16, 513, int((n_per > 0).sum()))
pos = np.arange(512)[None, :]
ids[:, 0] = 101
ids[np.arange(C), valid - 1] = 102
ids = np.where(pos < valid[:, None], ids, 0)
mask = (pos < valid[:, None]).astype(np.int64)
ids_h, mask_h = torch.from_numpy(ids).pin_memory(), torch.from_numpy(mask).pin_memory()

sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), 7).items()}
enc = modules.BioClinicalBERT_FT.from_state_dict(sd).to(dev)
del sd
torch.manual_seed(0)
fame = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), dev)
fame.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.fame_shapes(lab_tokens=L), 4).items()})
fame = fame.to(dev).eval()
keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features")
shard = {k: torch.from_numpy(meta[k][p_lo:p_hi]).to(dev) for k in keys + ("labels",)}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return r, t.item()


def encode_and_pool():
    cls = torch.empty((C, 768), device=dev, dtype=torch.float32)
    for s in range(0, C, CB):
        e = min(C, s + CB)
        cls[s:e] = enc(ids_h[s:e].to(dev, non_blocking=True), mask_h[s:e].to(dev, non_blocking=True))
    return modules.pool_chunks(cls, torch.from_numpy(offs).to(dev))


def model_forward(text):
    outs = []
    n = p_hi - p_lo
    with torch.no_grad():
        for s in range(0, n, PB):
            e = min(n, s + PB)
            o = fame(*[shard[k][s:e] for k in keys], text[s:e])
            outs.append(o["fused_logits"])
    return torch.cat(outs)


def metrics_pass(logits):
    attrs = [shard["age_ids"], shard["ethnicity_ids"], shard["insurance_ids"]]
    labels = shard["labels"].float()
    sweep = torch.from_numpy(metrics._SWEEP).to(dev)
    vec = ops.eval_counts(logits, labels, attrs, (0.5, 0.5, 0.5), sweep=sweep)
    parallel.all_reduce_sum_(vec, group)
    th = metrics.thresholds_from_hist(metrics.Counts(vec).hist)                 # calibrate_thresholds on the cohort
    return th, parallel.evaluate_sharded(logits, labels, attrs, th, group=group, verbose=False)


encode_and_pool() if C <= 4 * CB else enc(ids_h[:CB].to(dev), mask_h[:CB].to(dev))   # warm-up
text, t_enc = timed(encode_and_pool)
logits, t_model = timed(lambda: model_forward(text))
(th, (m, fair, eddi)), t_metrics = timed(lambda: metrics_pass(logits))
c_total = int(offs_all[-1])
if rank == 0:
    res = {"workload": f"eval sweep: {P} patients x U{{1..16}} chunks = {c_total} chunks of 512 tokens, {world} GPU(s), "
                       f"patients sharded by chunk count", "n_gpus": world, "patients": P, "chunks": c_total,
           "ms": {"note_encoder_and_pool (H2D of ids/mask included)": t_enc, "fame_model_forward": t_model,
                  "thresholds_and_metrics (counts all-reduce, logit all-gather, AUROC/AP ranks, EDDI, EO)": t_metrics},
           "chunks_per_s": c_total / (t_enc * 1e-3), "patients_per_s_end_to_end": P / ((t_enc + t_model + t_metrics) * 1e-3),
           "thresholds": th, "auroc": {k: v["aucroc"] for k, v in m.items()}, "eddi_overall": eddi["overall"],
           "eo_overall": {k: v["overall_eo"] for k, v in fair.items()}}
    print(json.dumps(res), flush=True)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)
if world > 1:
    barrier()
    parallel.shutdown(exit_code=0)

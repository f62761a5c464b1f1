"""Data-parallel plumbing for the FAME hot path: one process per GPU, torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The reference is single-process; the semantics defined here are "N ranks ==
the single-process reference on the concatenated global batch" (SURVEY.md section 8e):

  * patients are split into contiguous ranges, a patient's note chunks travel with it (CSR offsets re-based), so
    the chunk->patient pooling never crosses a rank and needs no collective;
  * the only data-path collectives are SUM all-reduces of (a) the 104 int64 loss statistics, (b) the flat fp32
    gradient buffer, (c) the integer evaluation counts, plus an all-gather of logits for the exact AUROC / AP ranks.
"""
from __future__ import annotations

import numpy as np
import torch


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n items for `rank`; sizes differ by at most one, order preserved."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_patients_by_chunks(offsets, world: int):
    """Contiguous patient ranges balanced by chunk count (the note encoder's cost is per chunk, not per patient).
    offsets: int array [P+1] (CSR).  Returns [(p_lo, p_hi)] * world covering 0..P without gaps."""
    offsets = np.asarray(offsets, dtype=np.int64)
    P = len(offsets) - 1
    total = int(offsets[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        # first patient whose starting chunk index reaches the target, never moving backwards
        p = int(np.searchsorted(offsets[:-1], target, side="left"))
        cuts.append(min(max(p, cuts[-1]), P))
    cuts.append(P)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def rebase_offsets(offsets, p_lo: int, p_hi: int):
    """CSR offsets of patients [p_lo, p_hi) re-based to start at 0, and the chunk range they cover."""
    offsets = np.asarray(offsets)
    c_lo, c_hi = int(offsets[p_lo]), int(offsets[p_hi])
    return (offsets[p_lo:p_hi + 1] - c_lo).astype(np.int32), (c_lo, c_hi)


def all_reduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_gather_rows(t: torch.Tensor, sizes, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks of different lengths (sizes[r] rows on rank r) in rank order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)])


def evaluate_sharded(logits, labels, attrs, thresholds, group=None, verbose=False):
    """Evaluation of a patient-sharded test set: integer counts are SUM-all-reduced, logits / labels all-gathered
    for the rank statistics (each rank ranks its own slice against all), then the host formulas run identically on
    every rank.  Returns the same triple as metrics.evaluate_from_logits on the concatenated data."""
    import torch.distributed as dist
    from . import metrics, ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return metrics.evaluate_from_logits(logits, labels, attrs, thresholds, verbose=verbose)
    n_local = torch.tensor([logits.shape[0]], device=logits.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    g_logits = all_gather_rows(logits, sizes, group)
    g_labels = all_gather_rows(labels, sizes, group)
    g_attrs = [all_gather_rows(a, sizes, group) for a in attrs]
    th = [thresholds[n] if isinstance(thresholds, dict) else thresholds for n in metrics.OUTCOMES]
    # counts from the LOCAL shard, summed over ranks == counts of the whole cohort (integers: exact)
    vec = ops.eval_counts(logits, labels, attrs, th)
    dist.all_reduce(vec, group=group)
    local = metrics.evaluate_from_logits(g_logits, g_labels, g_attrs, thresholds, verbose=verbose, counts=vec,
                                         rank_group=group)
    return local


def shutdown(timeout_s: float = 30.0, exit_code: int = 0):
    """Tear the default process group down without risking a hang at exit: destroy_process_group runs in a helper
    thread; if NCCL does not return within `timeout_s` (seen when CUDA graphs that captured its kernels were still
    alive) the process exits anyway.  Call after every result has been printed and flushed."""
    import os
    import sys
    import threading
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout_s)
    if t.is_alive():
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(exit_code)

"""FAME training step on the B200 (drop-in for ``train_step``, 10_FAME.py:401-449).

The reference runs  forward -> BCE + 10*lambda*LEDDI + L1 -> total_loss.backward() -> clip_grad_norm_(1.0) ->
AdamW.step()  through torch.autograd.  Here the same step is an explicit sequence of sm_100a kernels:

  forward   demo tower (12-layer BERT over one token, fp32 residual stream), lab tower (2 post-norm layers, flash
            attention head_dim 96), fusion head -- activations needed by the backward are kept
  loss      fame_loss_stats -> [SUM all-reduce of 104 int64 when data parallel] -> fame_loss_fwd_bwd (loss, dlogits)
  backward  fusion head (fp32), lab tower and demo tower: dgrad / wgrad tensor-core GEMMs (fame_gemm_ex with MN-major
            operands, no transposed copies), LayerNorm / GELU / softmax backward kernels, batched attention backward
  update    [all-reduce of the flat gradient buffer when data parallel] -> squared norm -> fused clip + AdamW

All trainable parameters live in ONE flat fp32 buffer (``FlatTrainState``): the nn.Parameters of the model are views
into it, their ``.grad`` are views into a flat gradient buffer, AdamW state is two more flat buffers and a bf16 shadow
feeds the tensor cores.  Parameters that receive no gradient in the reference (the three modality classifiers and
the BERT pooler: they are not on the loss path, 10_FAME.py:444) are left out, exactly as torch skips ``grad is None``.

Dropout: as in the reference's train() mode -- BERT embedding / hidden / attention-probability dropout in the
demographic tower (HF:111,205,297,355), dropout / dropout1 / dropout2 / attention dropout in the lab tower's
TransformerEncoderLayers, fusion_mlp[2] -- with the probabilities read from the model (``modules.set_dropout(model, 0)``
gives the parity configuration of SURVEY.md 7.2).  Masks come from a counter-based hash evaluated inside the kernels
(csrc/dropout.cuh): nothing is stored, the backward regenerates them, and a device-side step counter feeds the seed
so that CUDA-graph replays draw fresh masks.
"""
from __future__ import annotations

import math

import torch

from . import ops
from . import ops_train as T

import os

TWO_STREAMS = os.environ.get("FAME_TWO_STREAMS", "1") != "0"   # demographic and lab towers on concurrent streams
# SMs left free by the lab tower's persistent kernels while the demographic stream (and NCCL) runs beside it
RESERVED_SMS = int(os.environ.get("FAME_RESERVED_SMS", "16"))
RESERVED_SMS_DP = int(os.environ.get("FAME_RESERVED_SMS_DP", "32"))   # data parallel: the NCCL kernels need SMs too

NO_GRAD_PREFIXES = ("classifier_demo.", "classifier_lab.", "classifier_text.", "behrt_demo.bert.pooler.")
ATTR_IDX = (2, 4, 5)          # age_ids, ethnicity_ids, insurance_ids inside the 9-tensor batch (10_FAME.py:431)


# Regions of the flat buffers in layout order (= the order in which the backward completes their gradients):
#   QK     query / key projection WEIGHTS of the demographic BERT: with its length-1 sequences the softmax is identically
#          1, so their gradient is exactly zero on every rank (never written) -- never reduced
#   demo i the four large matrices (value, attention output, intermediate, output) of demographic layer i, 11 first
#   lab 1  the four large matrices of lab layers >= 1;   lab 0: those of lab layer 0
#   small  everything else: biases, LayerNorms, embeddings, the fusion head (read as fp32 by the kernels)
# The large matrices are only ever read through their bf16 shadows, so their fp32 masters and Adam moments can live on
# ONE rank each (sharded optimizer); "small" stays replicated.
_R_QK, _R_DEMO_TOP, _R_LAB1, _R_LAB0, _R_SMALL = 0, 1, 13, 14, 15
_R_NORED = _R_QK
_REGION_ALIGN = 64            # elements: every region (hence every bucket) splits evenly over 2 / 4 / 8 ranks of 8-element rows
_DEMO_BIG = ("attention.self.value.weight", "attention.output.dense.weight", "intermediate.dense.weight",
             "output.dense.weight")
_LAB_BIG = ("self_attn.in_proj_weight", "self_attn.out_proj.weight", "linear1.weight", "linear2.weight")


def _r_demo(i):
    return _R_DEMO_TOP + (11 - i)


FAME_NAMES = dict(demo="behrt_demo.", lab="behrt_lab.",
                  head=("demo_projector.", "lab_projector.", "text_projector.", "fusion_mlp."))


def _layout_key(name, names=FAME_NAMES):
    """Region of the flat buffers a parameter lives in (see FlatTrainState.__init__).  `names`: module prefixes of the
    demographic tower, the lab tower and the head (FAME_NAMES; sigmoid_fusion.py uses the reference's other names)."""
    enc = names["demo"] + "bert.encoder.layer."
    if name.startswith(enc):
        i = int(name[len(enc):].split(".")[0])
        if name.endswith(("attention.self.query.weight", "attention.self.key.weight")):
            return _R_QK
        if name.endswith(_DEMO_BIG) and 0 <= i <= 11:
            return _r_demo(i)
        return _R_SMALL
    lab = names["lab"] + "transformer_encoder.layers."
    if name.startswith(lab) and name.endswith(_LAB_BIG):
        return _R_LAB0 if int(name[len(lab):].split(".")[0]) == 0 else _R_LAB1
    return _R_SMALL


def demo_bucket_layers(n_layers=12):
    """A gradient bucket closes after each of these demographic-BERT layers (the backward runs n-1 -> 0 on the side
    stream next to the lab backward): two large buckets that cross NVLink under the lab tower's backward."""
    return (max(1, (7 * n_layers) // 12), 0) if n_layers > 1 else (0,)


def dropout_base_seed():
    """Seed of the dropout hash: follows torch.manual_seed (so runs / folds with different seeds draw different masks,
    as the reference's Philox stream does) and the data-parallel rank (the mask row index is rank-local: without the
    rank every shard of the global batch would draw the SAME masks)."""
    s = torch.initial_seed() & 0xFFFFFFFF
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            s ^= (dist.get_rank() * 0x9E3779B1) & 0xFFFFFFFF
    except Exception:
        pass
    return (s ^ 0x5EED) & 0xFFFFFFFF


class DropSites:
    """Dropout sites of one training step: name -> _lib.DropoutCfg (seed per site, shared device step counter)."""

    def __init__(self, model, step_dev, base_seed=None, lab_module=None, head_dropout=None, demo_module=None):
        """FAME model by default; with lab_module / head_dropout given: a model that has only the lab tower and one
        nn.Dropout in its head (behrt_combined.BEHRTModel_Combined).  base_seed: default = dropout_base_seed()."""
        from . import _lib
        base_seed = dropout_base_seed() if base_seed is None else base_seed
        self._lib, self.step_ptr, self.base, self.cache = _lib, step_dev.data_ptr(), base_seed, {}
        active = model.training
        if lab_module is None:
            cfg = model.behrt_demo.bert.config
            self.p_demo_hidden = float(getattr(cfg, "hidden_dropout_prob", 0.0)) if active else 0.0
            self.p_demo_attn = float(getattr(cfg, "attention_probs_dropout_prob", 0.0)) if active else 0.0
            lab_module, head_dropout = model.behrt_lab, model.fusion_mlp[2]
        else:
            self.p_demo_hidden = self.p_demo_attn = 0.0
            if demo_module is not None:
                cfg = demo_module.bert.config
                self.p_demo_hidden = float(getattr(cfg, "hidden_dropout_prob", 0.0)) if active else 0.0
                self.p_demo_attn = float(getattr(cfg, "attention_probs_dropout_prob", 0.0)) if active else 0.0
        self.lab = []
        for l in lab_module.transformer_encoder.layers:
            self.lab.append(dict(attn=float(l.self_attn.dropout) if active else 0.0, d1=float(l.dropout1.p) if active else 0.0,
                                 act=float(l.dropout.p) if active else 0.0, d2=float(l.dropout2.p) if active else 0.0))
        self.p_fusion = float(head_dropout.p) if (active and head_dropout is not None) else 0.0
        self.any = active and (self.p_demo_hidden > 0 or self.p_demo_attn > 0 or self.p_fusion > 0 or
                               any(v > 0 for d in self.lab for v in d.values()))

    def key(self):
        return (self.p_demo_hidden, self.p_demo_attn, self.p_fusion, tuple(tuple(sorted(d.items())) for d in self.lab))

    def site(self, name, p, group_shift=0):
        """DropoutCfg of site `name` (None when p == 0: the kernels then take their no-dropout instantiation)."""
        if p <= 0.0:
            return None
        if not 0.0 < p < 1.0:
            raise ValueError(f"dropout probability {p} of {name} outside (0, 1)")
        c = self.cache.get((name, p, group_shift))
        if c is None:
            import zlib
            c = self._lib.DropoutCfg()
            c.step = self.step_ptr
            c.seed = (zlib.crc32(name.encode()) ^ (self.base * 0x9E3779B1)) & 0xFFFFFFFF
            c.thresh16 = max(1, min(65535, int(round(p * 65536.0))))
            c.group_shift = group_shift
            self.cache[(name, p, group_shift)] = c
        return c

    @staticmethod
    def inv_keep(c):
        return 1.0 if c is None else 65536.0 / (65536.0 - c.thresh16)


def plan_layout(named_sizes, names=FAME_NAMES):
    """[(name, numel)] sorted by region -> (offsets {name: first element}, regions {region: (lo, hi)}, total)."""
    offsets, region, off, prev = {}, {}, 0, None
    for n, k in named_sizes:
        r = _layout_key(n, names)
        if r != prev:
            if prev is not None:                                 # close the previous region on a shardable boundary
                off = (off + _REGION_ALIGN - 1) // _REGION_ALIGN * _REGION_ALIGN
                region[prev] = (region[prev][0], off)
            prev = r
        offsets[n] = off
        lo, _ = region.get(r, (off, off))
        off += (k + 7) // 8 * 8                                  # 32-byte aligned segments
        region[r] = (lo, off)
    off = (off + _REGION_ALIGN - 1) // _REGION_ALIGN * _REGION_ALIGN
    if prev is not None:
        region[prev] = (region[prev][0], off)
    return offsets, region, off


# how a bucket of the flat gradient buffer is combined over data-parallel ranks
LOCAL, SHARDED, REPLICATED = "local", "sharded", "replicated"


def plan_buckets(region, total):
    """Contiguous gradient buckets over the regions of plan_layout, in the order the backward closes them:
    {key: (lo, hi, kind)}.  kind LOCAL: zero gradient everywhere, nothing to reduce (optimizer sharded); SHARDED:
    reduce-scatter -> AdamW on this rank's 1/N -> all-gather of the bf16 shadow; REPLICATED: all-reduce, AdamW on
    every rank (see FlatTrainState.grad_buckets)."""
    assert _R_SMALL in region
    demo_regions = sorted(r for r in region if _R_DEMO_TOP <= r < _R_LAB1)
    n_demo = (max(11 - (r - _R_DEMO_TOP) for r in demo_regions) + 1) if demo_regions else 0
    cuts = {}
    if _R_QK in region:
        cuts["qk"] = region[_R_QK] + (LOCAL,)
    if demo_regions:
        lo = region[demo_regions[0]][0]
        for i in demo_bucket_layers(n_demo):
            if _r_demo(i) in region and region[_r_demo(i)][1] > lo:
                cuts[("demo", i)] = (lo, region[_r_demo(i)][1], SHARDED)
                lo = region[_r_demo(i)][1]
    if _R_LAB1 in region:
        cuts[("lab", 1)] = region[_R_LAB1] + (SHARDED,)
    tail_lo = region[_R_LAB0][0] if _R_LAB0 in region else region[_R_SMALL][0]
    cuts["tail"] = (tail_lo, region[_R_SMALL][1], REPLICATED)
    # sanity: the buckets tile [0, total) exactly, each on a boundary that splits over 2 / 4 / 8 ranks
    spans = sorted((a, b) for a, b, _ in cuts.values())
    assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:])), spans
    assert all(a % _REGION_ALIGN == 0 and b % _REGION_ALIGN == 0 for a, b in spans), spans
    return cuts


class FlatTrainState:
    """Flat parameter / gradient / AdamW-state buffers for one MultimodalTransformer_EDDI_Sigmoid."""

    def __init__(self, model, no_grad_prefixes=NO_GRAD_PREFIXES, fame_layout=True, names=FAME_NAMES):
        """fame_layout: the region / bucket layout of MultimodalTransformer_EDDI_Sigmoid (below).  Other models
        (behrt_combined.BEHRTModel_Combined) use module order and a single all-reduce bucket."""
        self.model = model
        self.fame_layout = fame_layout
        self.names = names
        dev = next(model.parameters()).device
        named = [(n, p) for n, p in model.named_parameters() if not (no_grad_prefixes and n.startswith(no_grad_prefixes))]
        # Layout = the order in which the backward completes gradients, so that every reduction bucket is ONE
        # contiguous range (see the region table above _layout_key).
        if fame_layout:
            named.sort(key=lambda np_: _layout_key(np_[0], names))   # stable: module order inside each region
        self.offsets, self.region, self.n = plan_layout([(n, p.numel()) for n, p in named], names)
        off = self.n
        self.p = torch.zeros(off, device=dev, dtype=torch.float32)
        self.g = torch.zeros(off, device=dev, dtype=torch.float32)
        self.m = torch.zeros(off, device=dev, dtype=torch.float32)
        self.v = torch.zeros(off, device=dev, dtype=torch.float32)
        self.pb = torch.zeros(off, device=dev, dtype=torch.bfloat16)
        self.views, self.gviews, self.bviews = {}, {}, {}
        with torch.no_grad():
            for n, p in named:
                o, k = self.offsets[n], p.numel()
                self.p[o:o + k].copy_(p.detach().reshape(-1).float())
                p.data = self.p[o:o + k].view(p.shape)           # the module parameter now aliases the flat buffer
                p.grad = self.g[o:o + k].view(p.shape)
                self.views[n], self.gviews[n] = p.data, p.grad
                self.bviews[n] = self.pb[o:o + k].view(p.shape)
        self.step = 0
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.sumsq_scratch = torch.zeros(1, device=dev, dtype=torch.float64)
        self.active_plan, self.masters_stale = None, False
        self.grad_norm = torch.zeros(1, device=dev, dtype=torch.float32)
        # device-side step counter and {lr, weight_decay}: read by the AdamW kernel at run time, so a captured CUDA
        # graph of the step follows the schedule without being re-captured
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.hyper_dev = torch.zeros(2, device=dev, dtype=torch.float32)
        self._hyper_host = None
        # EDDI modality weights of the fusion head, read by the kernels at run time: update_dynamic_weights_all_tasks
        # changes them every epoch (10_FAME.py:805-830) and a captured step must not be re-captured for that
        self.w_mod_dev = torch.full((3,), 0.33, device=dev, dtype=torch.float32)
        self._w_mod_host = (0.33, 0.33, 0.33)
        self.last_stats = None
        self.graphs = {}
        self._build_transposed_shadows(dev)
        self.refresh_bf16()
        self._params = [p for _, p in named]
        self._view_ptrs = [self.views[n].data_ptr() for n, _ in named]
        self._versions = self._param_versions()

    # weights of the demographic tower whose data-gradient product runs on the <= 32-row skinny path: dX = dY . W reads
    # W^T rows, so a transposed bf16 shadow is kept next to the plain one (one table-driven transpose launch per step)
    _T_SUFFIXES = ("attention.self.value.weight", "attention.output.dense.weight", "intermediate.dense.weight",
                   "output.dense.weight")

    def _build_transposed_shadows(self, dev):
        import numpy as np
        enc = self.names["demo"] + "bert.encoder.layer."
        names = [n for n in self.offsets if n.startswith(enc) and n.endswith(self._T_SUFFIXES)]
        total = sum(self.views[n].numel() for n in names)
        self.pbT = torch.zeros(max(total, 8), device=dev, dtype=torch.bfloat16)
        self.tviews = {}
        rec = np.zeros(len(names), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("rows", "<i4"), ("cols", "<i4"),
                                                   ("tile0", "<i4"), ("tiles_x", "<i4")]))
        off = tiles = 0
        for i, n in enumerate(names):
            rows, cols = self.views[n].shape
            tv = self.pbT[off:off + rows * cols].view(cols, rows)
            self.tviews[n] = tv
            tx, ty = (cols + 63) // 64, (rows + 63) // 64
            rec[i] = (self.bviews[n].data_ptr(), tv.data_ptr(), rows, cols, tiles, tx)
            tiles += tx * ty
            off += rows * cols
        self._t_entries, self._t_tiles = len(names), tiles
        self._t_table = torch.from_numpy(rec.view(np.uint8).copy()).to(dev) if names else None

    def grad_buckets(self):
        """{key: (lo, hi, kind)}: contiguous ranges of the flat buffers keyed by the backward event that completes their
        gradient: ('demo', i) closes once layer i of the demographic BERT is done (its backward runs 11 -> 0 on the
        side stream), ('lab', 1) once the lab backward has passed layer 1, 'tail' = large matrices of lab layer 0 +
        every small tensor, complete when both tower streams have joined; 'qk' never carries a gradient.  Few large
        buckets: an NCCL collective over 335 MB takes 0.88 ms on 8 B200s in one piece and 1.67 ms in eight."""
        if getattr(self, "_buckets", None) is not None:
            return self._buckets
        cuts = plan_buckets(self.region, self.n) if self.fame_layout else {"tail": (0, self.n, REPLICATED)}
        self._buckets = cuts
        return cuts

    # ---- data-parallel optimizer sharding (ZeRO-1 over the large matrices) ------------------------------------
    def shard_plan(self, group):
        """(rank, world) when the optimizer state of the SHARDED / LOCAL buckets lives on one rank each, else None:
        needs a process group of 2, 4 or 8 ranks (bucket boundaries are multiples of 64 elements) and FAME_SHARDED_OPT."""
        if group is None or not SHARDED_OPT or not self.fame_layout:
            return None
        import torch.distributed as dist
        world = dist.get_world_size(group)
        if world not in (2, 4, 8):
            return None
        return dist.get_rank(group), world

    def my_range(self, key, plan):
        """This rank's slice of bucket `key` (the whole bucket when it is replicated or nothing is sharded)."""
        lo, hi, kind = self.grad_buckets()[key]
        if plan is None or kind == REPLICATED:
            return lo, hi
        r, w = plan
        s = (hi - lo) // w
        return lo + r * s, lo + (r + 1) * s

    def gather_shadows(self, group, plan):
        """All-gather of the bf16 shadows of the SHARDED buckets (each rank updated its 1/N in the previous step's
        AdamW), asynchronously on NCCL's stream in the order the forward needs them.  Returns {key: work}."""
        import torch.distributed as dist
        works = {}
        order = [k for k in (("lab", 1),) if k in self.grad_buckets()] + \
                sorted((k for k in self.grad_buckets() if isinstance(k, tuple) and k[0] == "demo"), key=lambda k: k[1])
        for key in order:
            lo, hi, kind = self.grad_buckets()[key]
            if kind != SHARDED or hi <= lo:
                continue
            a, b = self.my_range(key, plan)
            works[key] = dist.all_gather_into_tensor(self.pb[lo:hi], self.pb[a:b], group=group, async_op=True)
        return works

    def sync_masters(self, group):
        """After sharded steps the fp32 masters (and Adam moments) of a sharded bucket are current on their owner only:
        all-gather them so that model.state_dict() / evaluation / checkpoints see the same parameters on every rank
        (once per epoch: train_step calls it before returning)."""
        plan = self.shard_plan(group)
        if plan is None or not getattr(self, "masters_stale", False):
            return
        import torch.distributed as dist
        for key, (lo, hi, kind) in self.grad_buckets().items():
            if kind == REPLICATED or hi <= lo:
                continue
            a, b = self.my_range(key, plan)
            for buf in (self.p, self.m, self.v, self.pb):
                dist.all_gather_into_tensor(buf[lo:hi], buf[a:b], group=group)
        self.masters_stale = False
        self.invalidate_caches()

    def _param_versions(self):
        return tuple(p._version for p in self._params)

    def aliased(self):
        """False once the module's parameters stopped being views of the flat buffer (e.g. model.to(other_device) or a
        parameter was re-assigned): the state must then be rebuilt."""
        return all(p.data_ptr() == ptr for p, ptr in zip(self._params, self._view_ptrs))

    def sync_external_updates(self):
        """The kernels update the flat fp32 buffer behind torch's back (no version bump), so a changed `_version` means
        that somebody ELSE wrote the parameters in place -- load_state_dict(), an initialiser, p.copy_() under no_grad.
        The bf16 shadows the tensor cores read are then stale: refresh them (one cast pass) before the next step."""
        v = self._param_versions()
        if v != self._versions:
            self.refresh_bf16()
            self.invalidate_caches()
            self._versions = v

    def side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.p.device, priority=-1)   # short kernels first when an SM frees
        return self._side

    def wgrad_stream(self):
        """Branch for the weight gradients of the demographic tower: they are off the data-gradient chain (see
        _demo_backward)."""
        if getattr(self, "_wgs", None) is None:
            self._wgs = torch.cuda.Stream(device=self.p.device)
        return self._wgs

    def post_stream(self):
        if getattr(self, "_post", None) is None:
            self._post = torch.cuda.Stream(device=self.p.device, priority=-1)
        return self._post

    def refresh_transposed(self):
        if self._t_table is not None:
            T.transpose_bf16_table(self._t_table, self._t_entries, self._t_tiles)

    def refresh_bf16(self):
        T.cast_bf16(self.p, self.pb)
        self.refresh_transposed()

    def wt(self, name):         # transposed bf16 shadow [in_features, out_features], or None
        return self.tviews.get(name)

    def w(self, name):          # bf16 shadow (GEMM operand)
        return self.bviews[name]

    def f(self, name):          # fp32 master
        return self.views[name]

    def gr(self, name):         # fp32 gradient
        return self.gviews[name]

    def zero_grad(self):
        self.g.zero_()

    def set_hyper(self, lr, weight_decay):
        if self._hyper_host != (lr, weight_decay):
            self.hyper_dev.copy_(torch.tensor([lr, weight_decay], dtype=torch.float32))
            self._hyper_host = (lr, weight_decay)

    def set_w_mod(self, w_mod):
        w = tuple(float(x) for x in w_mod)
        if w != self._w_mod_host:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("modality weights changed during CUDA-graph capture")
            self.w_mod_dev.copy_(torch.tensor(w, dtype=torch.float32))
            self._w_mod_host = w

    def state_dict(self):
        """Optimizer state for checkpoints (the nn.Module's own state_dict() holds the parameters; a torch AdamW built
        on model.parameters() never steps here, so ITS state_dict is empty): Adam moments, step count, dropout step."""
        return {"m": self.m.clone(), "v": self.v.clone(), "step": int(self.step_dev.item()),
                "layout": dict(self.offsets), "n": int(self.n)}

    def load_state_dict(self, sd):
        if int(sd["n"]) != int(self.n) or dict(sd["layout"]) != dict(self.offsets):
            raise ValueError("optimizer state was saved for a different parameter layout")
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.step = int(sd["step"])
        self.step_dev.fill_(int(sd["step"]))
        qk = self.grad_buckets().get("qk") if self.fame_layout else None
        if qk is not None:      # moments of the zero-gradient region loaded from elsewhere: only then run the full AdamW on it
            self.qk_moments_zero = not bool(self.m[qk[0]:qk[1]].any().item() or self.v[qk[0]:qk[1]].any().item())

    def clip_and_step(self, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        """clip_grad_norm_(max_norm) + AdamW on the flat buffers (graph-capturable: lr / weight decay / step count are
        read from device memory).  The bf16 shadow is refreshed by the same kernel.  After a data-parallel
        forward_backward with a sharded optimizer every rank updates its slice of the sharded buckets and the whole
        replicated bucket (the next forward_backward all-gathers the shadows)."""
        if not torch.cuda.is_current_stream_capturing():
            self.set_hyper(lr, weight_decay)
        self.step += 1
        self.step_dev += 1
        plan = getattr(self, "active_plan", None)
        if not getattr(self, "sumsq_valid", False):       # forward_backward already summed it bucket by bucket
            if plan is not None:
                raise RuntimeError("sharded optimizer step without the bucket-wise gradient reduction of forward_backward")
            self.sumsq.zero_()
            T.grad_sumsq(self.g, self.sumsq)
        self.sumsq_valid = False
        if plan is None:
            ranges = [(0, self.n)]
        else:
            ranges = [self.my_range(k, plan) for k in self.grad_buckets()]
            self.masters_stale = True
        # the 'qk' bucket (query / key weights of the length-1 demographic BERT) has gradient exactly 0 in every step, so
        # its Adam moments stay 0 and its AdamW step is the weight decay alone (fame_decay_only: 10 instead of 32 bytes
        # per parameter for 14.2 M of the 97.9 M parameters) -- unless moments were loaded from elsewhere (qk_moments_zero)
        qk = self.grad_buckets().get("qk") if (self.fame_layout and getattr(self, "qk_moments_zero", True)) else None
        for a, b in ranges:
            if qk is not None and a < qk[1] and b > qk[0]:
                lo, hi = max(a, qk[0]), min(b, qk[1])
                if a < lo:
                    self._adamw(a, lo, max_norm, lr, betas, eps, weight_decay)
                T.decay_only(self.p[lo:hi], lr, weight_decay, hyper_dev=self.hyper_dev, p_bf16=self.pb[lo:hi])
                a = hi
            if b > a:
                self._adamw(a, b, max_norm, lr, betas, eps, weight_decay)
        # the transposed shadows are refreshed by the next forward_backward (beside its forward pass, not here)
        self.invalidate_caches()

    def _adamw(self, a, b, max_norm, lr, betas, eps, weight_decay):
        T.clip_adamw(self.p[a:b], self.g[a:b], self.m[a:b], self.v[a:b], self.sumsq, max_norm, lr, betas[0], betas[1], eps,
                     weight_decay, 0, self.grad_norm, step_dev=self.step_dev, hyper_dev=self.hyper_dev, p_bf16=self.pb[a:b])

    def invalidate_caches(self):
        """The kernels update the flat buffer behind torch's back (no version bump): drop the inference-path
        packed-weight caches so the next eval forward re-reads the parameters."""
        for m in self.model.modules():
            if getattr(m, "_packed", None) is not None:
                m._packed = None


def sync_parameters(model, group=None):
    """After data-parallel optimisation steps with the sharded optimizer: all-gather the fp32 masters / Adam moments of
    the sharded buckets so that model.state_dict(), evaluation and checkpoints are identical on every rank.  train_step
    calls it once per epoch; call it yourself after driving optimisation_step directly.  No-op otherwise."""
    st = getattr(model, "_fame_train_state", None)
    if st is not None and group is not None:
        st.sync_masters(group)


def describe_parallel(model, group):
    """One line for logs / the bench config: how the step is parallelised over `group`."""
    if group is None:
        return "single process"
    st = get_state(model)
    plan = st.shard_plan(group)
    mb = lambda k: sum((hi - lo) * 4 for key, (lo, hi, kind) in st.grad_buckets().items() if kind == k) / 1e6
    stat = "all-reduce(SUM) of 104 int64 loss statistics"
    if plan is None:
        return f"{stat} + all-reduce(SUM) of the flat fp32 gradient buffer ({mb(SHARDED) + mb(REPLICATED):.0f} MB), AdamW replicated"
    return (f"{stat}; ZeRO-1 over the large matrices: reduce-scatter of {mb(SHARDED):.0f} MB fp32 gradients in "
            f"{sum(1 for *_, k in st.grad_buckets().values() if k == SHARDED)} buckets -> AdamW on 1/{plan[1]} -> all-gather of the bf16 "
            f"shadows at the start of the next step; all-reduce of the {mb(REPLICATED):.0f} MB replicated tail; "
            f"{mb(LOCAL):.0f} MB of query/key weights never reduced")


def release_graphs(model):
    """Drop the captured step graphs of `model` (they hold the NCCL kernels of the process group that was current
    at capture time: destroy them before the process group, or destroy_process_group can block)."""
    st = getattr(model, "_fame_train_state", None)
    if st is not None:
        st.graphs.clear()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def get_state(model) -> FlatTrainState:
    st = getattr(model, "_fame_train_state", None)
    if st is None or st.model is not model or not st.aliased():
        if st is not None and st.model is model:
            import warnings
            warnings.warn("the model's parameters no longer alias the flat training buffers (model.to(...) or a "
                          "re-assigned parameter): rebuilding them -- Adam moments, step count and captured graphs "
                          "restart from zero (save / restore them with FlatTrainState.state_dict())")
        st = FlatTrainState(model)
        object.__setattr__(model, "_fame_train_state", st)
    return st


# ------------------------------------------------------------------------------------------------ demo tower
DEMO_TABLES = ("age", "gender", "ethnicity", "insurance")


def _demo_forward(st, model, ids, age, gender, eth, ins, ds=None, demo_module=None, dpre="behrt_demo.", codes=None,
                  table_names=DEMO_TABLES):
    """BEHRTModel_Demo.forward for training (sequence length 1): returns (demo_emb f32 [B,768], saved).  codes /
    table_names: other sets of code tables (ablation 07: seven of them); default = the four of 10_FAME.py."""
    demo_module = model.behrt_demo if demo_module is None else demo_module
    pre = dpre + "bert."
    B, S = ids.shape
    if S != 1:
        raise NotImplementedError("training path of the demographic encoder handles the reference's length-1 input")
    H = 768
    eps = demo_module.bert.config.layer_norm_eps
    dev = ids.device
    ph = ds.p_demo_hidden if ds is not None else 0.0
    pa = ds.p_demo_attn if ds is not None else 0.0
    site = (lambda n, p, g=0: ds.site(n, p, g)) if ds is not None else (lambda n, p, g=0: None)
    e = pre + "embeddings."
    esum = torch.empty((B, H), device=dev, dtype=torch.float32)
    estats = torch.empty((B, 2), device=dev, dtype=torch.float32)
    x32 = torch.empty((B, H), device=dev, dtype=torch.float32)
    xb = ops.bert_embed(ids.to(torch.int64), st.f(e + "word_embeddings.weight"), st.f(e + "position_embeddings.weight"),
                        st.f(e + "token_type_embeddings.weight")[0], st.f(e + "LayerNorm.weight"),
                        st.f(e + "LayerNorm.bias"), eps, S, out_f32=x32, sum_out=esum, stats=estats)
    d_emb = site("demo.emb", ph)
    if d_emb is not None:                                   # BertEmbeddings.dropout (HF:111): same mask on both copies
        T.dropout_apply(x32, d_emb)
        T.dropout_apply(xb, d_emb)
    # attention-probability dropout of a length-1 sequence keeps or drops one whole head: one draw per head_dim columns
    hshift = (H // int(demo_module.bert.config.num_attention_heads)).bit_length() - 1
    saved = {"esum": esum, "estats": estats, "layers": [], "ids": ids.to(torch.int64).contiguous().view(-1), "hshift": hshift}
    for i in range(int(demo_module.bert.config.num_hidden_layers)):
        p = f"{pre}encoder.layer.{i}."
        s = {"xb": xb}
        # one key per sequence: softmax == 1, so attention-probability dropout (HF:205) keeps or drops a whole head of
        # the value projection: one draw per (patient, head) = per 64 columns
        v = ops.gemm_bias_act(xb, st.w(p + "attention.self.value.weight"), st.f(p + "attention.self.value.bias"),
                              drop=site(f"demo.{i}.attn", pa, hshift))
        t1 = ops.gemm_bias_act(v, st.w(p + "attention.output.dense.weight"), st.f(p + "attention.output.dense.bias"),
                               residual=x32, out_dtype=torch.float32, drop=site(f"demo.{i}.h1", ph))
        st1 = torch.empty((B, 2), device=dev, dtype=torch.float32)
        x1b, x1f = ops.layernorm(t1, st.f(p + "attention.output.LayerNorm.weight"),
                                 st.f(p + "attention.output.LayerNorm.bias"), eps, want_f32=True, stats=st1)
        if B <= T.SKINNY_MAX_ROWS and FUSE_GELU:
            pa_, h = T.linear_gelu_small(x1b, st.w(p + "intermediate.dense.weight"), st.f(p + "intermediate.dense.bias"))
        else:
            pa_ = ops.gemm_bias_act(x1b, st.w(p + "intermediate.dense.weight"), st.f(p + "intermediate.dense.bias"))
            h = T.gelu_fwd(pa_)
        t2 = ops.gemm_bias_act(h, st.w(p + "output.dense.weight"), st.f(p + "output.dense.bias"), residual=x1f,
                               out_dtype=torch.float32, drop=site(f"demo.{i}.h2", ph))
        st2 = torch.empty((B, 2), device=dev, dtype=torch.float32)
        xb, x32 = ops.layernorm(t2, st.f(p + "output.LayerNorm.weight"), st.f(p + "output.LayerNorm.bias"), eps,
                                want_f32=True, stats=st2)
        s.update(v=v, t1=t1, st1=st1, x1b=x1b, pre=pa_, h=h, t2=t2, st2=st2)
        saved["layers"].append(s)
    did = [age, gender, eth, ins] if codes is None else list(codes)
    tabs = [st.f(f"{dpre}{n}_embedding.weight") for n in table_names]
    saved["demo_ids"] = [t.to(torch.int64).contiguous() for t in did]
    saved["table_names"] = tuple(table_names)
    if len(tabs) == 4:
        return ops.demo_add(x32, H, saved["demo_ids"], tabs), saved
    return ops.embed_mean_add(x32, H, saved["demo_ids"], tabs), saved


FUSE_BIAS_GRAD = os.environ.get("FAME_FUSE_BIAS_GRAD", "1") != "0"
FUSE_GELU = os.environ.get("FAME_FUSE_GELU", "1") != "0"    # <= 32 rows: GELU forward / backward inside the skinny GEMM


def _lin_bwd(st, wname, bname, dy_bf16, x_bf16, colsum_src=None):
    """Bias and weight gradients of y = x W^T + b into the flat gradient buffer."""
    small = dy_bf16.shape[0] <= T.SKINNY_MAX_ROWS
    if small and FUSE_BIAS_GRAD:
        # <= 32 rows (demographic tower): the weight-gradient kernel also writes the bias gradient from the dY tile it
        # has staged -- one launch instead of two on a chain that is bound by launch latency
        T.linear_wgrad(dy_bf16, x_bf16, st.gr(wname), accumulate=False, dbias=st.gr(bname))
        return
    T.colsum(colsum_src if colsum_src is not None else dy_bf16, st.gr(bname))
    # every weight gradient is produced exactly once per step: the <= 32-row kernel overwrites its (already zeroed)
    # slice with plain stores, the split-K tensor-core product accumulates into it with atomics
    T.linear_wgrad(dy_bf16, x_bf16, st.gr(wname), accumulate=not small)


DEMO_BUCKET_LAYERS = demo_bucket_layers(12)
SHARDED_OPT = os.environ.get("FAME_SHARDED_OPT", "1") != "0"    # data parallel: ZeRO-1 over the large matrices


class _GradReducer:
    """Per bucket, as soon as the backward has completed it (asynchronous, NCCL over NVLink, overlapping the remaining
    backward kernels):
      SHARDED bucket    reduce-scatter (SUM): this rank receives the global gradient of ITS 1/N of the bucket (in place)
      REPLICATED bucket all-reduce (SUM)
      single process    nothing to reduce
    and then the bucket's contribution to the squared gradient norm on a third stream that follows the collective, so
    that after the last bucket only its own collective and 13 us of norm are left before AdamW.  finish() joins
    everything into the current stream; with a sharded optimizer the per-rank partial norms are summed there (one
    8-byte all-reduce; the replicated bucket is counted by rank 0 only)."""

    def __init__(self, st, group):
        self.st, self.group = st, group
        self.plan = st.shard_plan(group)
        st.active_plan = self.plan
        self.post = st.post_stream()
        self.used = False

    def ready(self, key):
        b = self.st.grad_buckets().get(key)
        if b is None:
            return
        lo, hi, kind = b
        if hi <= lo or kind == LOCAL:
            return
        cur = torch.cuda.current_stream()
        g = self.st.g[lo:hi]
        if self.group is not None:
            import torch.distributed as dist
            target = self.st.sumsq
            if self.plan is not None and kind == SHARDED:
                a, e = self.st.my_range(key, self.plan)
                mine = self.st.g[a:e]
                work = dist.reduce_scatter_tensor(mine, g, group=self.group, async_op=True)
                g = mine
            else:
                work = dist.all_reduce(g, group=self.group, async_op=True)
                if self.plan is not None and self.plan[0] != 0:
                    target = self.st.sumsq_scratch          # identical on every rank: only rank 0's copy is counted
            with torch.cuda.stream(self.post):
                work.wait()
                T.grad_sumsq(g, target)
        else:
            self.post.wait_stream(cur)
            with torch.cuda.stream(self.post):
                T.grad_sumsq(g, self.st.sumsq)
        self.used = True

    def finish(self):
        if self.used:
            torch.cuda.current_stream().wait_stream(self.post)
        if self.plan is not None:
            import torch.distributed as dist
            dist.all_reduce(self.st.sumsq, group=self.group)
        self.st.sumsq_valid = True


DEMO_WGRAD_BRANCH = os.environ.get("FAME_DEMO_WGRAD_BRANCH", "1") != "0"


class _WgradBranch:
    """Weight-gradient launches of the demographic tower on their own stream.  Per layer the backward is a chain of
    LayerNorm-backward and data-gradient kernels (each needs the previous one's output); the four weight gradients of a
    layer only CONSUME tensors of that chain (dY and the saved input), nothing on the chain waits for them.  At 32 rows
    every kernel here is latency bound (4-12 us each, 15 per layer), so taking the 48 weight-gradient launches off the
    chain shortens the tower's backward by about 40 %.  join() orders the branch before whatever follows on the current
    stream (a bucket's collective, the end of the backward); the tensors a branch kernel reads stay referenced until
    then (the caching allocator could otherwise hand their memory back to the chain's stream)."""

    def __init__(self, st):
        self.stream = st.wgrad_stream() if (DEMO_WGRAD_BRANCH and st is not None) else None
        self.keep = []

    def run(self, fn, *tensors):
        if self.stream is None:
            fn()
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.extend(tensors)

    def join(self):
        if self.stream is not None and self.keep:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.keep.clear()


def _demo_backward(st, model, saved, ddemo, reducer=None, ds=None, dpre="behrt_demo."):
    wg = _WgradBranch(st)
    pre = dpre + "bert."
    ph = ds.p_demo_hidden if ds is not None else 0.0
    pa = ds.p_demo_attn if ds is not None else 0.0
    site = (lambda n, p, g=0: ds.site(n, p, g)) if ds is not None else (lambda n, p, g=0: None)
    tabs_g = [st.gr(f"{dpre}{n}_embedding.weight") for n in saved.get("table_names", DEMO_TABLES)]
    if len(tabs_g) == 4:
        T.demo_add_bwd(ddemo, saved["demo_ids"], tabs_g)
    else:
        ops.embed_mean_add_bwd(ddemo, saved["demo_ids"], tabs_g)
    dx = ddemo                                                      # f32 [B,768]: gradient of the last hidden state
    cut_layers = demo_bucket_layers(len(saved["layers"]))
    for i in reversed(range(len(saved["layers"]))):
        p = f"{pre}encoder.layer.{i}."
        s = saved["layers"][i]
        # t = residual + dropout(dense(.)): the residual branch takes dt (f32), the dense layer's backward the masked
        # copy (without dropout the two coincide)
        dt2b, dt2f, dt2m = T.layernorm_bwd_drop(s["t2"], dx, s["st2"], st.f(p + "output.LayerNorm.weight"),
                                                st.gr(p + "output.LayerNorm.weight"), st.gr(p + "output.LayerNorm.bias"),
                                                want_bf16=ph <= 0, want_f32=True, drop=site(f"demo.{i}.h2", ph))
        if dt2m is None:
            dt2m = dt2b
            wg.run(lambda: _lin_bwd(st, p + "output.dense.weight", p + "output.dense.bias", dt2m, s["h"], colsum_src=dt2f),
                   dt2m, dt2f)
        else:
            wg.run(lambda: _lin_bwd(st, p + "output.dense.weight", p + "output.dense.bias", dt2m, s["h"]), dt2m)
        if dt2m.shape[0] <= T.SKINNY_MAX_ROWS and FUSE_GELU and st.wt(p + "output.dense.weight") is not None:
            dpre = T.linear_dgrad(dt2m, st.w(p + "output.dense.weight"), wT=st.wt(p + "output.dense.weight"), aux=s["pre"],
                                  aux_mode=T.AUX_GELU_BWD_BF16)
        else:
            dh = T.linear_dgrad(dt2m, st.w(p + "output.dense.weight"), wT=st.wt(p + "output.dense.weight"))
            dpre = T.gelu_bwd(s["pre"], dh)
        wg.run(lambda: _lin_bwd(st, p + "intermediate.dense.weight", p + "intermediate.dense.bias", dpre, s["x1b"]), dpre)
        dx1 = T.linear_dgrad(dpre, st.w(p + "intermediate.dense.weight"), out_dtype=torch.float32, aux=dt2f,
                             aux_mode=T.AUX_ADD_F32, wT=st.wt(p + "intermediate.dense.weight"))
        dt1b, dt1f, dt1m = T.layernorm_bwd_drop(s["t1"], dx1, s["st1"], st.f(p + "attention.output.LayerNorm.weight"),
                                                st.gr(p + "attention.output.LayerNorm.weight"),
                                                st.gr(p + "attention.output.LayerNorm.bias"), want_bf16=ph <= 0,
                                                want_f32=True, drop=site(f"demo.{i}.h1", ph))
        if dt1m is None:
            dt1m = dt1b
            wg.run(lambda: _lin_bwd(st, p + "attention.output.dense.weight", p + "attention.output.dense.bias", dt1m, s["v"],
                                    colsum_src=dt1f), dt1m, dt1f)
        else:
            wg.run(lambda: _lin_bwd(st, p + "attention.output.dense.weight", p + "attention.output.dense.bias", dt1m, s["v"]),
                   dt1m)
        # s["v"] is the context = head-dropped value projection; its gradient passes the same per-head mask
        dv = T.linear_dgrad(dt1m, st.w(p + "attention.output.dense.weight"),
                            wT=st.wt(p + "attention.output.dense.weight"), drop=site(f"demo.{i}.attn", pa, saved.get("hshift", 6)))
        wg.run(lambda: _lin_bwd(st, p + "attention.self.value.weight", p + "attention.self.value.bias", dv, s["xb"]), dv)
        # one key per sequence: softmax == 1, so query / key receive exactly zero gradient (their .grad stays 0 and
        # AdamW still applies weight decay to them, as in the reference)
        dx = T.linear_dgrad(dv, st.w(p + "attention.self.value.weight"), out_dtype=torch.float32, aux=dt1f,
                            aux_mode=T.AUX_ADD_F32, wT=st.wt(p + "attention.self.value.weight"))
        if reducer is not None and i in cut_layers:
            wg.join()                                               # the bucket's weight gradients are complete
            reducer.ready(("demo", i))
    e = pre + "embeddings."
    T.dropout_apply(dx, site("demo.emb", ph))                       # BertEmbeddings.dropout backward (no-op when off)
    _, dsum = T.layernorm_bwd(saved["esum"], dx, saved["estats"], st.f(e + "LayerNorm.weight"),
                              st.gr(e + "LayerNorm.weight"), st.gr(e + "LayerNorm.bias"), want_bf16=False, want_f32=True)
    T.bert_embed_bwd(dsum, saved["ids"], st.gr(e + "word_embeddings.weight"), st.gr(e + "position_embeddings.weight"),
                     st.gr(e + "token_type_embeddings.weight")[0], 1, pad_idx=0)
    wg.join()


# ------------------------------------------------------------------------------------------------ lab tower
def _lab_forward(st, model, lab, ds=None, lab_module=None, pre="behrt_lab."):
    lab_module = model.behrt_lab if lab_module is None else lab_module
    B, L = lab.shape
    H, nh = 768, lab_module.nhead
    lab = lab.float().contiguous()
    x = ops.lab_embed(lab, st.f(pre + "token_embedding.weight")[:, 0].contiguous(), st.f(pre + "token_embedding.bias"),
                      st.f(pre + "pos_embedding"))
    saved = {"lab": lab, "layers": []}
    dev = lab.device
    for i, layer in enumerate(lab_module.transformer_encoder.layers):
        p = f"{pre}transformer_encoder.layers.{i}."
        pr = ds.lab[i] if ds is not None else dict(attn=0.0, d1=0.0, act=0.0, d2=0.0)
        site = (lambda n, q: ds.site(f"lab.{i}.{n}", q)) if ds is not None else (lambda n, q: None)
        s = {"x": x}
        qkv = ops.gemm_bias_act(x, st.w(p + "self_attn.in_proj_weight"), st.f(p + "self_attn.in_proj_bias"))
        lse = torch.empty((B, nh, L), device=dev, dtype=torch.float32)       # saved for the attention backward
        ctx = ops.attn_fwd(qkv, B, L, nh, H // nh, lse=lse, drop=site("attn", pr["attn"]))
        # x = norm1(x + dropout1(out_proj(ctx)));  x = norm2(x + dropout2(linear2(dropout(relu(linear1(x))))))
        # the residual of this K = 768 projection is added by the LayerNorm kernel (and again by its backward), not by
        # the GEMM epilogue, whose scattered residual loads made the launch 3x longer than its main loop (ncu r02:
        # 57 us, tensor pipe 25 % active, against 19 us of math)
        t1 = ops.gemm_bias_act(ctx, st.w(p + "self_attn.out_proj.weight"), st.f(p + "self_attn.out_proj.bias"),
                               drop=site("d1", pr["d1"]))
        st1 = torch.empty((B * L, 2), device=dev, dtype=torch.float32)
        x1 = ops.layernorm(t1, st.f(p + "norm1.weight"), st.f(p + "norm1.bias"), layer.norm1.eps, stats=st1, residual=x)
        h = ops.gemm_bias_act(x1, st.w(p + "linear1.weight"), st.f(p + "linear1.bias"), act=ops.ACT_RELU,
                              drop=site("act", pr["act"]))
        t2 = ops.gemm_bias_act(h, st.w(p + "linear2.weight"), st.f(p + "linear2.bias"), residual=x1,
                               drop=site("d2", pr["d2"]))
        st2 = torch.empty((B * L, 2), device=dev, dtype=torch.float32)
        x = ops.layernorm(t2, st.f(p + "norm2.weight"), st.f(p + "norm2.bias"), layer.norm2.eps, stats=st2)
        s.update(qkv=qkv, ctx=ctx, lse=lse, t1=t1, st1=st1, x1=x1, h=h, t2=t2, st2=st2)
        saved["layers"].append(s)
    return ops.seq_mean(x, B, L), saved


FUSED_ATTN_BWD = os.environ.get("FAME_FUSED_ATTN_BWD", "1") != "0"   # 0: materialise P / dS + three batched GEMMs


def _attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, drop=None, fused=None):
    """dqkv [T, 3*nh*D] bf16 from dctx [T, nh*D].  Fused (default): fame_attn_bwd_fused -- both score products
    (Q K^T and dO V^T), the softmax recomputed from the forward's row log-sum-exp (and, in training with attention
    dropout, the forward's mask from its seed) and the products that consume P / dS run in one kernel per pass (dK/dV
    pass, dQ pass); P and dS stay in TMEM.  Unfused (FAME_FUSED_ATTN_BWD=0): P and dS are written once in bf16 and three
    batched tensor-core products turn them into dV, dK, dQ."""
    dev = qkv.device
    delta = T.attn_delta(dctx, ctx, B, L, nh, D)
    if FUSED_ATTN_BWD if fused is None else fused:
        return T.attn_bwd_fused(qkv, dctx, lse, delta, B, L, nh, D, D ** -0.5, drop=drop)
    W = 3 * nh * D
    HD = nh * D
    ldp = (L + 7) // 8 * 8
    sq = (L * W, D)                 # (b0 = sequence, b1 = head) strides inside the packed qkv tensor
    sc = (L * HD, D)                # same inside ctx / dctx
    ss = (nh * L * ldp, L * ldp)    # inside the [B, nh, L, ldp] probability / score-gradient tensors
    p, ds = T.attn_bwd_pds(qkv, dctx, lse, delta, B, L, nh, D, ldp, D ** -0.5, drop=drop)
    dqkv = torch.empty((B * L, W), device=dev, dtype=torch.bfloat16)
    # dV = P^T dO, dK = dS^T Q  (A MN-major: stored [query rows, key cols]; B MN-major: stored [query rows, d cols])
    T.gemm_ex(p, dctx, dqkv, L, D, L, a_mn=True, b_mn=True, lda=ldp, ldb=HD, ldy=W, nb0=B, nb1=nh, sa=ss, sb=sc, sy=sq,
              y_off=2 * HD, tag="attn_dv")
    T.gemm_ex(ds, qkv, dqkv, L, D, L, a_mn=True, b_mn=True, lda=ldp, ldb=W, ldy=W, nb0=B, nb1=nh, sa=ss, sb=sq, sy=sq,
              y_off=HD, tag="attn_dk")
    # dQ = dS K  (A K-major over keys; B = K stored [key rows, d cols] -> MN-major)
    T.gemm_ex(ds, qkv, dqkv, L, D, L, a_mn=False, b_mn=True, lda=ldp, ldb=W, ldy=W, nb0=B, nb1=nh, sa=ss, sb=sq, sy=sq,
              b_off=HD, tag="attn_dq")
    return dqkv


def _lab_backward(st, model, saved, dlab, ds=None, lab_module=None, pre="behrt_lab.", reducer=None):
    lab_module = model.behrt_lab if lab_module is None else lab_module
    lab = saved["lab"]
    B, L = lab.shape
    H, nh = 768, lab_module.nhead
    dx = T.seq_mean_bwd(dlab, B, L)                                  # bf16 [B*L, 768]
    for i in reversed(range(len(saved["layers"]))):
        p = f"{pre}transformer_encoder.layers.{i}."
        pr = ds.lab[i] if ds is not None else dict(attn=0.0, d1=0.0, act=0.0, d2=0.0)
        site = (lambda n, q: ds.site(f"lab.{i}.{n}", q)) if ds is not None else (lambda n, q: None)
        s = saved["layers"][i]
        # dt2: gradient of t2 = x1 + dropout2(linear2(h)) -> residual branch; dt2m (masked) -> linear2
        dt2, _, dt2m = T.layernorm_bwd_drop(s["t2"], dx, s["st2"], st.f(p + "norm2.weight"), st.gr(p + "norm2.weight"),
                                            st.gr(p + "norm2.bias"), drop=site("d2", pr["d2"]))
        dt2m = dt2 if dt2m is None else dt2m
        _lin_bwd(st, p + "linear2.weight", p + "linear2.bias", dt2m, s["h"])
        # s["h"] = dropout(relu(.)) is zero exactly where ReLU or the dropout zeroed it; the kept entries carry 1/(1-p)
        d_act = site("act", pr["act"])
        dh = T.linear_dgrad(dt2m, st.w(p + "linear2.weight"), aux=s["h"], aux_mode=T.AUX_RELU_MASK_BF16,
                            alpha=DropSites.inv_keep(d_act))
        _lin_bwd(st, p + "linear1.weight", p + "linear1.bias", dh, s["x1"])
        dx1 = T.linear_dgrad(dh, st.w(p + "linear1.weight"), aux=dt2, aux_mode=T.AUX_ADD_BF16)
        dt1, _, dt1m = T.layernorm_bwd_drop(s["t1"], dx1, s["st1"], st.f(p + "norm1.weight"), st.gr(p + "norm1.weight"),
                                            st.gr(p + "norm1.bias"), drop=site("d1", pr["d1"]), residual=s["x"])
        dt1m = dt1 if dt1m is None else dt1m
        _lin_bwd(st, p + "self_attn.out_proj.weight", p + "self_attn.out_proj.bias", dt1m, s["ctx"])
        dctx = T.linear_dgrad(dt1m, st.w(p + "self_attn.out_proj.weight"))
        dqkv = _attn_backward(s["qkv"], dctx, s["ctx"], s["lse"], B, L, nh, H // nh, drop=site("attn", pr["attn"]))
        _lin_bwd(st, p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", dqkv, s["x"])
        dx = T.linear_dgrad(dqkv, st.w(p + "self_attn.in_proj_weight"), aux=dt1, aux_mode=T.AUX_ADD_BF16)
        if reducer is not None and i == 1:
            reducer.ready(("lab", 1))                # the large matrices of every layer >= 1 are complete
    # token embedding: Linear(1, 768).weight has shape [768, 1] -> its gradient is the [768] vector
    T.lab_embed_bwd(dx, lab, st.gr(pre + "pos_embedding"), st.gr(pre + "token_embedding.weight").view(-1),
                    st.gr(pre + "token_embedding.bias"))


# ------------------------------------------------------------------------------------------------ fusion head
_PROJ = ("demo_projector.0.", "lab_projector.0.", "text_projector.0.")


def _fusion_pack(st):
    f = st.f
    return dict(
        wp_t=torch.stack([f(p + "weight").t().contiguous() for p in _PROJ]).contiguous(),
        bp=torch.stack([f(p + "bias") for p in _PROJ]).contiguous(), sig_w=f("sig_weights"),
        w3_t=f("fusion_mlp.0.weight").t().contiguous(), b3=f("fusion_mlp.0.bias"), w4=f("fusion_mlp.3.weight"),
        b4=f("fusion_mlp.3.bias"))


def _fusion_backward(st, fo, embs, dlogits, w_mod, lambda_l1, d_fus=None, w_mod_dev=None, branch=None):
    """Returns (d demo_emb, d lab_emb) f32 [B,768]; writes every head gradient into the flat buffer.  The chain
    dlogits -> dhid -> dgated -> dproj -> d embeddings is what both towers' backward passes wait for, so it is enqueued
    first; the head's own parameter gradients (five small products, five column sums) only consume tensors of that chain
    and run on `branch` (a _WgradBranch; the caller joins it before the tail bucket is reduced)."""
    B = dlogits.shape[0]
    dev = dlogits.device
    f, g = st.f, st.gr
    branch = _WgradBranch(None) if branch is None else branch
    hid = fo["hid_drop"] if d_fus is not None else torch.relu(fo["pre_relu"])  # [B,512] input of fusion_mlp[3]
    dhid = T.fusion_bwd_hidden(dlogits, f("fusion_mlp.3.weight"), fo["pre_relu"])
    T.dropout_apply(dhid, d_fus)                                               # fusion_mlp[2] backward (no-op when off)
    dgated = torch.empty((B, 768), device=dev, dtype=torch.float32)
    T.sgemm(dhid, 512, 1, f("fusion_mlp.0.weight"), 768, 1, dgated, B, 768, 512)   # dgated[B,768] = dhid W3
    dproj = T.fusion_bwd_gate(dgated, fo["proj"], f("sig_weights"), w_mod, lambda_l1, g("sig_weights"),
                              w_mod_dev=w_mod_dev)
    demb = []
    for m, pn in enumerate(_PROJ[:2]):                                          # the text embedding is an input
        dpm = dproj[:, 256 * m:256 * (m + 1)]                                   # [B,256] view, row stride 768
        de = torch.empty((B, 768), device=dev, dtype=torch.float32)
        T.sgemm(dpm, 768, 1, f(pn + "weight"), 768, 1, de, B, 768, 256)         # demb = dpm Wp
        demb.append(de)

    def param_grads():
        # fusion_mlp.3: dW4[3,512] = dlogits^T hid ; db4 = colsum(dlogits)
        T.sgemm(dlogits, 1, 3, hid, 512, 1, g("fusion_mlp.3.weight"), 3, 512, B)
        T.colsum(dlogits, g("fusion_mlp.3.bias"))
        # fusion_mlp.0: dW3[512,768] = dhid^T gated ; db3
        T.sgemm(dhid, 1, 512, fo["gated"], 768, 1, g("fusion_mlp.0.weight"), 512, 768, B)
        T.colsum(dhid, g("fusion_mlp.0.bias"))
        for m, pn in enumerate(_PROJ):
            dpm = dproj[:, 256 * m:256 * (m + 1)]
            # dWp[256,768] = dpm^T emb ; dbp = colsum(dpm)
            T.sgemm(dpm, 1, 768, embs[m], 768, 1, g(pn + "weight"), 256, 768, B)
            T.colsum(dpm, g(pn + "bias"))

    branch.run(param_grads, dlogits, hid, dhid, dproj, fo, embs)
    return demb


def _lab_budget(group=None):
    from . import _lib
    r = RESERVED_SMS if group is None else RESERVED_SMS_DP
    return max(2, _lib.load().fame_sm_count() - r) if r > 0 else 0


def begin_step(st, group=None):
    """Start of a training step, off the critical path on the third stream while the forward runs: zero the gradient
    buffer (55 us) and refresh the transposed bf16 shadows the demographic backward reads (85 us); the backward waits
    for this stream.  With a sharded optimizer the previous step's AdamW refreshed only this rank's 1/N of the large
    matrices' bf16 shadows: the rest is all-gathered now, on NCCL's stream, bucket by bucket in the order the forward
    consumes them (lab tower first: it is the critical path).  Returns {bucket: work}; a tower waits for its own
    buckets only (work.wait() on the stream that runs it).
    (Tried and measured no better, r02 device timelines: the refresh under the fusion head instead of beside the forward
    -- the head's short kernels then queue behind its 20 k CTAs; a grid-stride refresh with two CTAs per SM -- its
    long-lived CTAs hold shared memory the persistent tensor-core kernels need, whichever phase it runs in.)"""
    post = st.post_stream()
    post.wait_stream(torch.cuda.current_stream())
    plan = st.shard_plan(group)
    gathers = st.gather_shadows(group, plan) if plan is not None else {}
    with torch.cuda.stream(post):
        st.zero_grad()
        st.sumsq.zero_()
        st.sumsq_scratch.zero_()
        for k, w_ in gathers.items():
            if isinstance(k, tuple) and k[0] == "demo":
                w_.wait()
        st.refresh_transposed()
    return gathers


# ------------------------------------------------------------------------------------------------ one optimisation step
def forward_backward(model, batch, pos_weight, lambda_edd, lambda_l1, w_mod, group=None, want_outputs=False,
                     debug=None):
    """Forward + loss + backward for one batch; gradients land in the flat buffer.  Returns loss_out f32 [4] (device)
    = (total, bce, leddi, l1) of the GLOBAL batch.  debug: optional dict that receives the per-patient tensors a
    parity test compares (embeddings, dlogits, d loss / d embeddings)."""
    st = get_state(model)
    st.set_w_mod(w_mod)
    (ids, mask, age, gender, eth, ins, lab, text, labels) = batch
    # off the critical path, on the third stream while the forward runs: zero the gradient buffer, refresh the transposed
    # bf16 shadows the demographic backward reads; the backward waits for this stream below
    post = st.post_stream()
    gathers = begin_step(st, group)
    demo_gathers = [w for k, w in gathers.items() if isinstance(k, tuple) and k[0] == "demo"]
    ds = DropSites(model, st.step_dev)
    if not ds.any:
        ds = None                                                             # parity configuration / eval: no site active
    # The two towers are independent until the fusion head.  At 32 patients the demographic tower is ~270 launches of
    # 3-9 us (weight streaming, latency bound) and the lab tower a chain of tensor-core kernels, so they run on two
    # streams (two branches of the captured graph) and the short kernels fill the gaps between the long ones.
    main = torch.cuda.current_stream()
    side = st.side_stream() if TWO_STREAMS else None
    if side is not None:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for w_ in demo_gathers:
                w_.wait()
            demo, sv_d = _demo_forward(st, model, ids, age, gender, eth, ins, ds)
        if ("lab", 1) in gathers:
            gathers[("lab", 1)].wait()
        ops.set_sm_budget(_lab_budget(group))
        labe, sv_l = _lab_forward(st, model, lab, ds)
        ops.set_sm_budget(0)
        main.wait_stream(side)
    else:
        for w_ in gathers.values():
            w_.wait()
        demo, sv_d = _demo_forward(st, model, ids, age, gender, eth, ins, ds)
        labe, sv_l = _lab_forward(st, model, lab, ds)
    text = text.float().contiguous()
    pk = _fusion_pack(st)
    if want_outputs:
        cls = (model.classifier_demo, model.classifier_lab, model.classifier_text)
        pk["wc"] = torch.stack([c.weight.detach().float() for c in cls]).contiguous()
        pk["bc"] = torch.stack([c.bias.detach().float() for c in cls]).contiguous()
    else:
        pk["wc"] = pk["bc"] = pk["b4"]                                        # unused (mod_logits not requested)
    fo = ops.fusion_fwd((demo, labe, text), pk, w_mod, want_mod_logits=want_outputs, want_intermediates=True,
                        w_mod_dev=st.w_mod_dev)
    d_fus = ds.site("fusion.hidden", ds.p_fusion) if ds is not None else None
    if d_fus is not None:
        # fusion_mlp = Linear, ReLU, Dropout, Linear (10_FAME.py:255): the fused kernel's logits skip the dropout, so
        # the last layer is redone on the dropped hidden activations (three small launches, training only)
        hid = T.dropout_apply(torch.relu(fo["pre_relu"]), d_fus)
        logits = pk["b4"].repeat(hid.shape[0], 1)
        T.sgemm(hid, 512, 1, pk["w4"], 1, 512, logits, hid.shape[0], 3, 512, accumulate=True)
        fo["hid_drop"], fo["logits"] = hid, logits
    labels = labels.float().contiguous()
    attrs = [batch[i].to(torch.int64).contiguous() for i in ATTR_IDX]
    stats = ops.loss_stats(fo["logits"], labels, attrs, pos_weight)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(stats, group=group)                                    # 104 int64: global-batch statistics
    st.last_stats = stats                               # stats[103] != 0: a bad attribute code -> NaN loss (train_step raises)
    loss, dlogits = ops.loss_fwd_bwd(fo["logits"], labels, attrs, pos_weight, stats, st.f("sig_weights"), lambda_edd,
                                     lambda_l1)
    if group is not None:
        import torch.distributed as dist
        # every rank adds lambda_l1 * sign(sig_weights) below; the gradient all-reduce is a SUM -> share it out
        lambda_l1 = lambda_l1 / dist.get_world_size(group)
    # gradient SUM over ranks, bucket by bucket as the backward completes them (sig_weights, produced here by the
    # fusion head, lives in the 'rest' region and travels with the last demographic bucket)
    red = _GradReducer(st, group)
    main.wait_stream(post)                              # gradient buffer zeroed, transposed shadows current
    head_branch = _WgradBranch(st)
    ddemo, dlab = _fusion_backward(st, fo, (demo, labe, text), dlogits, w_mod, lambda_l1, d_fus, w_mod_dev=st.w_mod_dev,
                                   branch=head_branch)
    if debug is not None:
        debug.update(demo=demo, lab=labe, dlogits=dlogits, ddemo=ddemo, dlab=dlab, logits=fo["logits"],
                     proj=fo["proj"], pre_relu=fo["pre_relu"])
    # the demographic tower owns 88 % of the gradient bytes: its buckets cross NVLink while the tensor-core-bound lab
    # backward runs (side stream, or simply first when single-stream)
    if side is not None:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _demo_backward(st, model, sv_d, ddemo, red, ds)
        ops.set_sm_budget(_lab_budget(group))
        _lab_backward(st, model, sv_l, dlab, ds, reducer=red)
        ops.set_sm_budget(0)
        main.wait_stream(side)
    else:
        _demo_backward(st, model, sv_d, ddemo, red, ds)
        _lab_backward(st, model, sv_l, dlab, ds, reducer=red)
    head_branch.join()                                  # the head's parameter gradients live in the tail bucket
    red.ready("tail")
    red.finish()
    # tensors that crossed streams (allocated on one, read on the other) stay referenced until both branches have been
    # enqueued and joined: the caching allocator may hand a freed block back to its own stream immediately
    del sv_d, sv_l, ddemo, dlab, demo, labe
    return loss, fo


def _hyper(optimizer):
    g = optimizer.param_groups[0]
    return dict(lr=g["lr"], weight_decay=g.get("weight_decay", 0.01), betas=tuple(g.get("betas", (0.9, 0.999))),
                eps=g.get("eps", 1e-8))


def _pos_weight(criterion, device):
    pw = getattr(criterion, "pos_weight", None)
    if pw is None:
        return torch.ones(3, device=device, dtype=torch.float32)
    return pw.detach().to(device=device, dtype=torch.float32).contiguous()


def train_step(model, dataloader, optimizer, device, criterion, beta=1.0, lambda_edd=0.8, lambda_l1=0.01, target=1.0,
               threshold=0.5, old_eddi_weights=None, group=None):
    """Drop-in for 10_FAME.py:401-449: one pass over ``dataloader``, one optimizer step per batch; returns
    (running_loss, running_bce_loss) as python floats.  ``optimizer`` supplies lr / weight_decay / betas / eps (a
    torch.optim.AdamW built on model.parameters(), as the reference does); the update itself is the fused
    clip + AdamW kernel on the flat buffers.  ``target`` and ``threshold`` are ignored, as in the reference."""
    model.train()
    w_mod = model.modality_weights(old_eddi_weights)
    pw = _pos_weight(criterion, device)
    hp = _hyper(optimizer)
    st = get_state(model)
    acc = torch.zeros(2, device=device, dtype=torch.float32)
    for batch in dataloader:
        batch = [x.to(device, non_blocking=True) for x in batch]
        loss = optimisation_step(model, batch, pw, lambda_edd, lambda_l1, w_mod, hp, group=group)
        acc += loss[:2]                                                        # accumulate on device, sync once
    sync_parameters(model, group)               # sharded optimizer: every rank sees the whole fp32 model again
    running_loss, running_bce = acc.tolist()
    if running_loss != running_loss and st.last_stats is not None and int(st.last_stats[103].item()) != 0:
        raise ValueError("sensitive-attribute code outside 0..7 in a training batch (age / ethnicity / insurance ids): "
                         "the LEDDI subgroup statistics cannot represent it")
    return running_loss, running_bce


USE_CUDA_GRAPH = True
MAX_STEP_GRAPHS = 4


def optimisation_step(model, batch, pw, lambda_edd, lambda_l1, w_mod, hp, group=None, use_graph=None):
    """forward + loss + backward + clip + AdamW for one batch.  The ~350 kernel launches of a step are captured into
    a CUDA graph per batch shape (first call of a shape runs eagerly, the second captures, later ones replay): at
    32 patients per GPU the step is otherwise bound by host launch overhead, not by the GPU."""
    st = get_state(model)
    st.sync_external_updates()                 # e.g. load_state_dict() between two steps
    use_graph = USE_CUDA_GRAPH if use_graph is None else use_graph
    if group is not None and not _GRAPH_WITH_COLLECTIVES:
        use_graph = False
    st.set_hyper(hp["lr"], hp["weight_decay"])

    def eager(b):
        loss, _ = forward_backward(model, b, pw, lambda_edd, lambda_l1, w_mod, group=group)
        st.clip_and_step(hp["lr"], hp["weight_decay"], hp["betas"], hp["eps"], max_norm=1.0)
        return loss

    if not use_graph:
        return eager(batch)
    st.set_w_mod(w_mod)                         # device buffer: NOT part of the key (it changes every epoch)
    key = (tuple((tuple(x.shape), x.dtype) for x in batch), lambda_edd, lambda_l1, hp["betas"], hp["eps"],
           pw.data_ptr(), group is not None, DropSites(model, st.step_dev).key())
    entry = st.graphs.get(key)
    if entry is None:
        # bounded cache, least recently used first out: each captured step owns a private pool with all saved
        # activations (0.5 - 1 GB at 32 patients); a run sees two shapes per loader (full and ragged last batch)
        while len(st.graphs) >= MAX_STEP_GRAPHS:
            st.graphs.pop(next(iter(st.graphs)))
        st.graphs[key] = {"graph": None}
        return eager(batch)
    st.graphs[key] = st.graphs.pop(key)         # most recently used last
    if entry["graph"] is None:
        static = [torch.empty_like(x) for x in batch]
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            entry["loss"] = eager(static)
        entry["graph"], entry["static"] = g, static
        st.step -= 1                                          # the capture itself executed nothing
    for s, x in zip(entry["static"], batch):
        s.copy_(x, non_blocking=True)
    entry["graph"].replay()
    st.step += 1
    return entry["loss"]


_GRAPH_WITH_COLLECTIVES = True      # NCCL collectives are captured into the step graph (thread-local capture mode)


def fame_forward_train(model, batch8, w_mod, return_modality_logits, return_gated_vector, return_intermediate):
    """Forward of the model in training mode called directly (outside train_step): outputs only, no autograd graph
    (gradients are produced by train.forward_backward, not by torch.autograd)."""
    st = get_state(model)
    ids, mask, age, gender, eth, ins, lab, text = batch8
    with torch.no_grad():
        demo, _ = _demo_forward(st, model, ids, age, gender, eth, ins)
        labe, _ = _lab_forward(st, model, lab)
        pk = _fusion_pack(st)
        cls = (model.classifier_demo, model.classifier_lab, model.classifier_text)
        pk["wc"] = torch.stack([c.weight.detach().float() for c in cls]).contiguous()
        pk["bc"] = torch.stack([c.bias.detach().float() for c in cls]).contiguous()
        o = ops.fusion_fwd((demo, labe, text.float().contiguous()), pk, w_mod, want_mod_logits=return_modality_logits,
                           want_intermediates=return_gated_vector or return_intermediate)
    out = {"fused_logits": o["logits"], "dynamic_weights": {"demo": w_mod[0], "lab": w_mod[1], "text": w_mod[2]},
           "sigmoid_weights": o["sig"]}
    if return_modality_logits:
        out["modality_logits"] = {"demo": o["mod_logits"][0], "lab": o["mod_logits"][1], "text": o["mod_logits"][2]}
    if return_gated_vector:
        out["gated_vector"] = o["gated"]
    if return_intermediate:
        out["fusion_pre_relu"] = o["pre_relu"]
    return out


def demo_forward_train(module, *a):
    raise NotImplementedError("call the parent MultimodalTransformer_EDDI_Sigmoid (or train.train_step) in training mode")


def lab_forward_train(module, *a):
    raise NotImplementedError("call the parent MultimodalTransformer_EDDI_Sigmoid (or train.train_step) in training mode")

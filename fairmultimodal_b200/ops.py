"""Thin torch-tensor wrappers over the C ABI (``include/fame_b200.h``).

PyTorch is used for device memory and streams only: each wrapper checks dtype / contiguity, extracts raw
device pointers and launches on the current CUDA stream.  No op has a CPU or eager fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_RELU, DT_BF16, DT_F32  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# Launch accounting for bench.py: number of kernels launched, and (when tracing) CUDA-event pairs around each
# launch on the launching stream, tagged so that per-kernel durations and algorithmic work can be summed.
LAUNCHES = 0
_TRACE = None


def start_trace():
    global _TRACE
    _TRACE = []


def stop_trace():
    """Returns [(op name, tag, start_event, end_event, work)], work = algorithmic FLOPs or bytes of the launch."""
    global _TRACE
    t, _TRACE = _TRACE, None
    return t


def _call(name, args, work=0.0, tag="", ws=None):
    global LAUNCHES
    LAUNCHES += 1
    wp, wb = (ws.data_ptr(), ws.numel() * ws.element_size()) if ws is not None else (0, 0)
    if _TRACE is None:
        _lib.call(name, args, _stream(), wp, wb)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, args, _stream(), wp, wb)
    e1.record()
    _TRACE.append((name, tag, e0, e1, work))


def _call_flat(name, work, *args):
    global LAUNCHES
    LAUNCHES += 1
    if _TRACE is None:
        _lib.call_flat(name, _stream(), *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call_flat(name, _stream(), *args)
    e1.record()
    _TRACE.append((name, "", e0, e1, work))


def set_sm_budget(sms: int) -> None:
    """SMs that persistent tensor-core kernels launched from now on may occupy (0 = all)."""
    _lib.check(_lib.load().fame_set_sm_budget(int(sms)), "fame_set_sm_budget")


def _cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.FameError(f"{name} must be a CUDA tensor (fairmultimodal_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.FameError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _rowmajor(t: torch.Tensor, name: str) -> int:
    """Return the leading dimension of a 2-D tensor whose last dim is contiguous."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise _lib.FameError(f"{name} must be 2-D with a contiguous last dimension")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _set_drop(dst, drop):
    if drop is not None:
        dst.step, dst.seed, dst.thresh16, dst.group_shift = drop.step, drop.seed, drop.thresh16, drop.group_shift


def gemm_bias_act(x, w, bias=None, residual=None, act=ACT_NONE, out=None, out_dtype=torch.bfloat16, drop=None):
    """y = dropout(act(x @ w.T + bias)) (+ residual).  x [M,K] bf16, w [N,K] bf16, bias [N] f32, residual [M,N] bf16
    or f32; drop: a _lib.DropoutCfg (training) or None."""
    _cuda(x, "x", torch.bfloat16)
    _cuda(w, "w", torch.bfloat16)
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise _lib.FameError(f"gemm: x is [*,{K}] but w is [*,{w.shape[1]}]")
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype)
    a = _lib.GemmArgs()
    a.x, a.ldx = x.data_ptr(), _rowmajor(x, "x")
    a.w, a.ldw = w.data_ptr(), _rowmajor(w, "w")
    a.bias = _cuda(bias, "bias", torch.float32).data_ptr() if bias is not None else None
    if residual is not None:
        _cuda(residual, "residual")
        if residual.dtype not in (torch.bfloat16, torch.float32):
            raise _lib.FameError("residual must be bf16 or f32")
        a.residual, a.ldr = residual.data_ptr(), _rowmajor(residual, "residual")
        a.residual_dtype = DT_F32 if residual.dtype == torch.float32 else DT_BF16
    else:
        a.residual, a.ldr = None, 0
    a.y, a.ldy = out.data_ptr(), _rowmajor(out, "out")
    a.y_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    a.M, a.N, a.K, a.act = M, N, K, act
    _set_drop(a.drop, drop)
    _call("fame_gemm_bias_act", a, 2.0 * M * N * K, f"{'sk' if M <= 32 and K % 32 == 0 else 'tc'}:{M}x{N}x{K}")
    return out


def layernorm(x, gamma, beta, eps, out=None, out_f32=None, want_f32=False, stats=None, residual=None):
    """x bf16 or f32 [rows, cols] -> bf16 `out` (default) and / or f32 `out_f32`; optional {mean, rstd} per row.
    residual (bf16, optional): LN(x + residual).  Returns out (bf16) unless want_f32, in which case (out_bf16, out_f32)."""
    _cuda(x, "x")
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.FameError("layernorm: x must be bf16 or f32")
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16)
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((rows, cols), device=x.device, dtype=torch.float32)
    a = _lib.LayerNormArgs()
    a.x, a.ldx = x.data_ptr(), _rowmajor(x, "x")
    a.x_dtype = DT_F32 if x.dtype == torch.float32 else DT_BF16
    a.gamma = _cuda(gamma, "gamma", torch.float32).data_ptr()
    a.beta = _cuda(beta, "beta", torch.float32).data_ptr()
    a.y, a.ldy = out.data_ptr(), _rowmajor(out, "out")
    if out_f32 is not None:
        if _rowmajor(out_f32, "out_f32") != a.ldy:
            raise _lib.FameError("layernorm: bf16 and f32 outputs must share the leading dimension")
        a.y_f32 = out_f32.data_ptr()
    if stats is not None:
        a.stats = _cuda(stats, "stats", torch.float32).data_ptr()
    a.rows, a.cols, a.eps = rows, cols, eps
    if residual is not None:
        _cuda(residual, "residual", torch.bfloat16)
        a.residual, a.ldr = residual.data_ptr(), _rowmajor(residual, "residual")
    _call("fame_layernorm", a, (6.0 if residual is not None else 4.0) * rows * cols)
    return (out, out_f32) if want_f32 else out


def bert_embed(ids, word, pos, type_emb, gamma, beta, eps, seq_len, err_flag=None, out_f32=None, sum_out=None,
               stats=None):
    """ids int64 [tokens] (flattened [chunks, seq_len]) -> bf16 [tokens, hidden] (+ optional f32 copy, pre-LN sum
    and {mean, rstd} for the backward pass)."""
    _cuda(ids, "ids", torch.int64)
    ids = ids.contiguous().view(-1)
    hidden = word.shape[1]
    out = torch.empty((ids.numel(), hidden), device=ids.device, dtype=torch.bfloat16)
    a = _lib.BertEmbedArgs()
    if out_f32 is not None:
        a.y_f32 = _cuda(out_f32, "out_f32", torch.float32).data_ptr()
    if sum_out is not None:
        a.sum_out = _cuda(sum_out, "sum_out", torch.float32).data_ptr()
    if stats is not None:
        a.stats = _cuda(stats, "stats", torch.float32).data_ptr()
    a.ids = ids.data_ptr()
    a.word = _cuda(word, "word", torch.float32).data_ptr()
    a.pos = _cuda(pos, "pos", torch.float32).data_ptr()
    a.type0 = _cuda(type_emb, "type_emb", torch.float32).data_ptr()
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    a.y = out.data_ptr()
    a.err_flag = err_flag.data_ptr() if err_flag is not None else None
    a.tokens, a.seq_len, a.hidden, a.vocab, a.eps = ids.numel(), seq_len, hidden, word.shape[0], eps
    _call("fame_bert_embed", a, ids.numel() * (8.0 + 10.0 * hidden))
    return out


def mask_kv_len(key_mask):
    """uint8 [batch, seq] key mask -> int32 [batch]: |value| = 1 + index of the last attended key (0 if none); the value
    is >= 0 for a prefix mask (ones then zeros: a padded batch) and negative when the attended keys have holes (the
    attention kernel then reads the mask bytes instead of deriving key validity from the length)."""
    _cuda(key_mask, "key_mask", torch.uint8)
    if not key_mask.is_contiguous() or key_mask.dim() != 2:
        raise _lib.FameError("key_mask must be contiguous uint8 [batch, seq]")
    batch, seq = key_mask.shape
    out = torch.empty(batch, device=key_mask.device, dtype=torch.int32)
    _call_flat("fame_mask_kv_len", 1.0 * batch * seq, key_mask.data_ptr(), batch, seq, out.data_ptr())
    return out


def attn_fwd(qkv, batch, seq, heads, head_dim, key_mask=None, scale=None, out=None, algo=0, lse=None, kv_len=None,
             drop=None):
    """qkv bf16 [batch*seq, 3*heads*head_dim] -> ctx bf16 [batch*seq, heads*head_dim].  lse (optional f32
    [batch, heads, seq]) receives the row log-sum-exp the backward pass needs.  kv_len (optional int32 [batch], from
    mask_kv_len(key_mask)) lets the kernel skip key blocks that hold only masked keys; the result does not change."""
    _cuda(qkv, "qkv", torch.bfloat16)
    if out is None:
        out = torch.empty((batch * seq, heads * head_dim), device=qkv.device, dtype=torch.bfloat16)
    a = _lib.AttnFwdArgs()
    a.qkv, a.ld_qkv = qkv.data_ptr(), _rowmajor(qkv, "qkv")
    if key_mask is not None:
        _cuda(key_mask, "key_mask", torch.uint8)
        if not key_mask.is_contiguous() or key_mask.numel() != batch * seq:
            raise _lib.FameError("key_mask must be contiguous uint8 [batch, seq]")
        a.key_mask = key_mask.data_ptr()
    else:
        a.key_mask = None
    a.ctx, a.ld_ctx = out.data_ptr(), _rowmajor(out, "ctx")
    a.batch, a.seq, a.heads, a.head_dim = batch, seq, heads, head_dim
    a.scale = float(scale) if scale is not None else head_dim ** -0.5
    a.algo = algo
    a.lse = _cuda(lse, "lse", torch.float32).data_ptr() if lse is not None else None
    if kv_len is not None:
        if key_mask is None or kv_len.numel() != batch:
            raise _lib.FameError("kv_len needs key_mask and must be int32 [batch]")
        a.kv_len = _cuda(kv_len, "kv_len", torch.int32).data_ptr()
    else:
        a.kv_len = None
    _set_drop(a.drop, drop)
    _call("fame_attn_fwd", a, 4.0 * batch * heads * seq * seq * head_dim)
    return out


def attn_cls(q, kv, batch, seq, heads, head_dim, k_col0, v_col0, key_mask=None, scale=None):
    """Attention of ONE query per sequence: q bf16 [batch, heads*head_dim] (the query projection of the CLS rows), kv
    bf16 [batch*seq, *] holding K at columns k_col0.. and V at v_col0.. -> ctx bf16 [batch, heads*head_dim]."""
    _cuda(q, "q", torch.bfloat16)
    _cuda(kv, "kv", torch.bfloat16)
    out = torch.empty((batch, heads * head_dim), device=q.device, dtype=torch.bfloat16)
    mp = None
    if key_mask is not None:
        _cuda(key_mask, "key_mask", torch.uint8)
        if not key_mask.is_contiguous() or key_mask.numel() != batch * seq:
            raise _lib.FameError("key_mask must be contiguous uint8 [batch, seq]")
        mp = key_mask.data_ptr()
    _call_flat("fame_attn_cls", 4.0 * batch * seq * heads * head_dim, q.data_ptr(), _rowmajor(q, "q"), kv.data_ptr(),
               _rowmajor(kv, "kv"), k_col0, v_col0, mp, out.data_ptr(), _rowmajor(out, "ctx"), batch, seq, heads,
               head_dim, float(scale) if scale is not None else head_dim ** -0.5)
    return out


def segment_mean(x, offsets, cols=None, ldx=None, mode="mean"):
    """out[p] = mean of rows offsets[p]:offsets[p+1] of x (row stride ldx elements); zeros for empty segments."""
    _cuda(x, "x")
    _cuda(offsets, "offsets", torch.int32)
    if x.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.FameError("segment_mean: x must be bf16 or f32")
    if ldx is None:
        ldx = _rowmajor(x, "x")
    if cols is None:
        cols = x.shape[-1]
    patients = offsets.numel() - 1
    out = torch.empty((patients, cols), device=x.device, dtype=torch.float32)
    a = _lib.SegmentMeanArgs()
    a.x, a.ldx = x.data_ptr(), ldx
    a.x_dtype = DT_BF16 if x.dtype == torch.bfloat16 else DT_F32
    a.offsets, a.out = offsets.data_ptr(), out.data_ptr()
    a.patients, a.cols = patients, cols
    a.mode = {"mean": 0, "max": 1}[mode]
    # algorithmic bytes: rows read (not known without a sync: caller may refine) + offsets + output
    _call("fame_segment_mean", a, 4.0 * (patients + 1) + 4.0 * patients * cols)
    return out


def lab_embed(lab, w_tok, b_tok, pos):
    """lab f32 [B, L] -> bf16 [B*L, hidden]:  lab * w_tok + b_tok + pos[l]."""
    _cuda(lab, "lab", torch.float32)
    B, L = lab.shape
    hidden = pos.shape[1]
    out = torch.empty((B * L, hidden), device=lab.device, dtype=torch.bfloat16)
    a = _lib.LabEmbedArgs()
    a.lab = lab.contiguous().data_ptr()
    a.w_tok = _cuda(w_tok, "w_tok", torch.float32).data_ptr()
    a.b_tok = _cuda(b_tok, "b_tok", torch.float32).data_ptr()
    a.pos = _cuda(pos, "pos", torch.float32).data_ptr()
    a.y = out.data_ptr()
    a.batch, a.L, a.hidden = B, L, hidden
    _call("fame_lab_embed", a, B * L * (4.0 + 2.0 * hidden))
    return out


def seq_mean(x, batch, L):
    """x bf16 [batch*L, cols] -> f32 [batch, cols] (mean over the L rows of each sequence)."""
    _cuda(x, "x", torch.bfloat16)
    cols = x.shape[1]
    out = torch.empty((batch, cols), device=x.device, dtype=torch.float32)
    a = _lib.SeqMeanArgs()
    a.x, a.out = x.data_ptr(), out.data_ptr()
    a.batch, a.L, a.cols = batch, L, cols
    _call("fame_seq_mean", a, 2.0 * batch * L * cols + 4.0 * batch * cols)
    return out


def demo_add(cls, ld_cls, ids, tables):
    """out[b] = cls_row[b] + mean of the 4 demographic embedding rows (ids clamped to the table range)."""
    _cuda(cls, "cls")
    if cls.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.FameError("demo_add: cls must be bf16 or f32")
    B = ids[0].numel()
    hidden = tables[0].shape[1]
    out = torch.empty((B, hidden), device=cls.device, dtype=torch.float32)
    a = _lib.DemoAddArgs()
    a.cls, a.ld_cls = cls.data_ptr(), ld_cls
    a.cls_dtype = DT_F32 if cls.dtype == torch.float32 else DT_BF16
    keep = []
    for k in range(4):
        i = _cuda(ids[k], "ids", torch.int64).contiguous()
        t = _cuda(tables[k], "table", torch.float32).contiguous()
        keep += [i, t]
        a.ids[k], a.table[k], a.n_rows[k] = i.data_ptr(), t.data_ptr(), t.shape[0]
    a.out, a.batch, a.hidden = out.data_ptr(), B, hidden
    _call("fame_demo_add", a, B * hidden * (2.0 + 16.0 + 4.0))
    return out


def embed_mean_add(cls, ld_cls, ids, tables):
    """out[b] = cls_row[b] + mean over the n = len(tables) <= 8 code tables of table_k[clamp(ids_k[b])] (f32 [B, hidden])."""
    _cuda(cls, "cls")
    if cls.dtype not in (torch.bfloat16, torch.float32):
        raise _lib.FameError("embed_mean_add: cls must be bf16 or f32")
    n = len(tables)
    if not 1 <= n <= 8 or len(ids) != n:
        raise _lib.FameError("embed_mean_add: 1..8 tables with one id vector each")
    B, hidden = ids[0].numel(), tables[0].shape[1]
    out = torch.empty((B, hidden), device=cls.device, dtype=torch.float32)
    a = _lib.EmbedMeanArgs()
    a.cls, a.ld_cls, a.cls_dtype, a.n_tables = cls.data_ptr(), ld_cls, DT_F32 if cls.dtype == torch.float32 else DT_BF16, n
    keep = []
    for k in range(n):
        i = _cuda(ids[k], "ids", torch.int64).contiguous()
        t = _cuda(tables[k], "table", torch.float32).contiguous()
        keep += [i, t]
        a.ids[k], a.table[k], a.n_rows[k] = i.data_ptr(), t.data_ptr(), t.shape[0]
    a.out, a.batch, a.hidden = out.data_ptr(), B, hidden
    _call("fame_embed_mean_add", a, B * hidden * (4.0 + 4.0 * n + 4.0))
    return out


def embed_mean_add_bwd(dout, ids, dtables):
    """dtable_k[clamp(ids_k[b])] += dout[b] / n into the (zeroed) f32 gradient tables."""
    n = len(dtables)
    B, hidden = dout.shape
    a = _lib.EmbedMeanArgs()
    a.n_tables = n
    keep = []
    for k in range(n):
        i = _cuda(ids[k], "ids", torch.int64).contiguous()
        keep.append(i)
        a.ids[k], a.dtable[k], a.n_rows[k] = i.data_ptr(), _cuda(dtables[k], "dtable", torch.float32).data_ptr(), dtables[k].shape[0]
    a.dout = _cuda(dout, "dout", torch.float32).contiguous().data_ptr()
    a.batch, a.hidden = B, hidden
    _call("fame_embed_mean_add_bwd", a, B * hidden * 4.0 * (1 + 2 * n))


def fusion_fwd(emb, packed, w_mod, want_mod_logits=False, want_intermediates=False, w_mod_dev=None):
    """emb = (demo, lab, text) f32 [B,768]; packed = dict of fp32 fusion weights (see modules._pack_fusion).
    Returns dict(logits, sig, [mod_logits], [proj, gated, pre_relu])."""
    B = emb[0].shape[0]
    dev = emb[0].device
    a = _lib.FusionFwdArgs()
    keep = []
    for m in range(3):
        e = _cuda(emb[m], "emb", torch.float32).contiguous()
        keep.append(e)
        a.emb[m] = e.data_ptr()
        a.w_mod[m] = float(w_mod[m])
    for k in ("wp_t", "bp", "sig_w", "w3_t", "b3", "w4", "b4", "wc", "bc"):
        setattr(a, k, packed[k].data_ptr())
    out = {"logits": torch.empty((B, 3), device=dev, dtype=torch.float32),
           "sig": torch.empty(768, device=dev, dtype=torch.float32)}
    a.logits, a.sig_out = out["logits"].data_ptr(), out["sig"].data_ptr()
    if want_mod_logits:
        out["mod_logits"] = torch.empty((3, B, 3), device=dev, dtype=torch.float32)
        a.mod_logits = out["mod_logits"].data_ptr()
    if want_intermediates:
        out["proj"] = torch.empty((B, 768), device=dev, dtype=torch.float32)
        out["gated"] = torch.empty((B, 768), device=dev, dtype=torch.float32)
        out["pre_relu"] = torch.empty((B, 512), device=dev, dtype=torch.float32)
        a.proj, a.gated, a.pre_relu = out["proj"].data_ptr(), out["gated"].data_ptr(), out["pre_relu"].data_ptr()
    a.B = B
    if w_mod_dev is not None:        # f32 [3] on the device: read by the kernels at run time (overrides w_mod)
        a.w_mod_dev = _cuda(w_mod_dev, "w_mod_dev", torch.float32).data_ptr()
    ws_bytes = 0 if want_intermediates else _lib.load().fame_fusion_fwd_workspace_bytes(B)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8) if ws_bytes else None
    _call("fame_fusion_fwd", a, B * (3 * 768 * 4.0 + 3 * 4.0) + 4.0 * (3 * 768 * 256 + 768 * 512), ws=ws)
    return out


def _attr_ptrs(a, attrs, keep):
    for k in range(3):
        t = _cuda(attrs[k], "attr", torch.int64).contiguous()
        keep.append(t)
        a.attr[k] = t.data_ptr()


def loss_stats(logits, labels, attrs, pos_weight, stats=None):
    """Per-rank LEDDI / BCE statistics (int64 [104]); accumulates into `stats` when given."""
    _cuda(logits, "logits", torch.float32)
    B = logits.shape[0]
    if stats is None:
        stats = torch.zeros(_lib.LOSS_STATS_LEN, device=logits.device, dtype=torch.int64)
    a = _lib.LossStatsArgs()
    keep = [logits.contiguous(), _cuda(labels, "labels", torch.float32).contiguous(),
            _cuda(pos_weight, "pos_weight", torch.float32).contiguous()]
    a.logits, a.labels, a.pos_weight = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
    _attr_ptrs(a, attrs, keep)
    a.stats, a.B = stats.data_ptr(), B
    _call("fame_loss_stats", a, 48.0 * B)
    return stats


def loss_fwd_bwd(logits, labels, attrs, pos_weight, stats, sig_w, lambda_edd, lambda_l1, want_grad=True):
    """Global statistics -> (loss_out f32 [4] = total, bce, leddi, l1;  dlogits f32 [B,3] or None)."""
    B = logits.shape[0]
    dev = logits.device
    a = _lib.LossFwdBwdArgs()
    keep = [logits.contiguous(), labels.contiguous(), pos_weight.contiguous()]
    a.logits, a.labels, a.pos_weight = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr()
    _attr_ptrs(a, attrs, keep)
    a.stats = _cuda(stats, "stats", torch.int64).data_ptr()
    if sig_w is not None:
        a.sig_w, a.n_sig = _cuda(sig_w, "sig_w", torch.float32).data_ptr(), sig_w.numel()
    a.lambda_edd, a.lambda_l1 = float(lambda_edd), float(lambda_l1)
    loss = torch.empty(4, device=dev, dtype=torch.float32)
    dlogits = torch.empty((B, 3), device=dev, dtype=torch.float32) if want_grad else None
    a.loss_out = loss.data_ptr()
    a.dlogits = dlogits.data_ptr() if want_grad else None
    a.B = B
    _call("fame_loss_fwd_bwd", a, 60.0 * B)
    return loss, dlogits


def eval_counts(logits, labels, attrs, thr, sweep=None, out=None, logits_are_probs=False):
    """Integer confusion counts per (outcome, attr, code) + totals (+ F1-sweep histogram).  uint64-as-int64 [914]."""
    _cuda(logits, "logits", torch.float32)
    N = logits.shape[0]
    if out is None:
        out = torch.zeros(_lib.EVAL_COUNTS_LEN, device=logits.device, dtype=torch.int64)
    a = _lib.EvalCountsArgs()
    keep = [_cuda(labels, "labels", torch.float32).contiguous()]
    a.logits, a.ld, a.labels = logits.data_ptr(), _rowmajor(logits, "logits"), keep[0].data_ptr()
    _attr_ptrs(a, attrs, keep)
    for k in range(3):
        a.thr[k] = float(thr[k])
    if sweep is not None:
        keep.append(_cuda(sweep, "sweep", torch.float64).contiguous())
        a.sweep = keep[-1].data_ptr()
    a.out, a.N, a.logits_are_probs = out.data_ptr(), N, int(logits_are_probs)
    _call("fame_eval_counts", a, 48.0 * N)
    return out


def sigmoid_probs(logits, labels=None):
    """float32 probabilities [3, N] (outcome-major) and uint8 labels [3, N]."""
    N = logits.shape[0]
    probs = torch.empty((3, N), device=logits.device, dtype=torch.float32)
    y8 = torch.empty((3, N), device=logits.device, dtype=torch.uint8) if labels is not None else None
    a = _lib.SigmoidProbsArgs()
    a.logits, a.ld = _cuda(logits, "logits", torch.float32).data_ptr(), _rowmajor(logits, "logits")
    if labels is not None:
        labels = _cuda(labels, "labels", torch.float32).contiguous()
        a.labels, a.y8 = labels.data_ptr(), y8.data_ptr()
    a.probs, a.N = probs.data_ptr(), N
    _call("fame_sigmoid_probs", a, 28.0 * N)
    return probs, y8


def rank_counts(scores, y8, i0=0, i1=None, acc=None):
    """Tie-aware rank statistics of one outcome.  Returns acc = dict(auroc2 int64[1], ap f64[1], pn int64[2])."""
    N = scores.numel()
    i1 = N if i1 is None else i1
    dev = scores.device
    if acc is None:
        acc = {"auroc2": torch.zeros(1, device=dev, dtype=torch.int64), "ap": torch.zeros(1, device=dev, dtype=torch.float64),
               "pn": torch.zeros(2, device=dev, dtype=torch.int64)}
    if i1 <= i0:
        return acc
    lib = _lib.load()
    ws_bytes = lib.fame_rank_counts_workspace_bytes(i1 - i0)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    a = _lib.RankCountsArgs()
    a.scores, a.y = _cuda(scores, "scores", torch.float32).data_ptr(), _cuda(y8, "y8", torch.uint8).data_ptr()
    a.N, a.i0, a.i1 = N, i0, i1
    a.auroc2, a.ap_sum, a.npos_nneg = acc["auroc2"].data_ptr(), acc["ap"].data_ptr(), acc["pn"].data_ptr()
    global LAUNCHES
    LAUNCHES += 2
    _lib.call("fame_rank_counts", a, _stream(), ws.data_ptr(), ws_bytes)
    return acc

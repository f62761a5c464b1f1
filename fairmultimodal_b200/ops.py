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


def _call(name, args, work=0.0, tag=""):
    global LAUNCHES
    LAUNCHES += 1
    if _TRACE is None:
        _lib.call(name, args, _stream())
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, args, _stream())
    e1.record()
    _TRACE.append((name, tag, e0, e1, work))


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


def gemm_bias_act(x, w, bias=None, residual=None, act=ACT_NONE, out=None, out_dtype=torch.bfloat16):
    """y = act(x @ w.T + bias) (+ residual).  x [M,K] bf16, w [N,K] bf16, bias [N] f32, residual [M,N] bf16."""
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
        _cuda(residual, "residual", torch.bfloat16)
        a.residual, a.ldr = residual.data_ptr(), _rowmajor(residual, "residual")
    else:
        a.residual, a.ldr = None, 0
    a.y, a.ldy = out.data_ptr(), _rowmajor(out, "out")
    a.y_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    a.M, a.N, a.K, a.act = M, N, K, act
    _call("fame_gemm_bias_act", a, 2.0 * M * N * K, f"{N}x{K}")
    return out


def layernorm(x, gamma, beta, eps, out=None):
    _cuda(x, "x", torch.bfloat16)
    rows, cols = x.shape
    if out is None:
        out = torch.empty_like(x)
    a = _lib.LayerNormArgs()
    a.x, a.ldx = x.data_ptr(), _rowmajor(x, "x")
    a.gamma = _cuda(gamma, "gamma", torch.float32).data_ptr()
    a.beta = _cuda(beta, "beta", torch.float32).data_ptr()
    a.y, a.ldy = out.data_ptr(), _rowmajor(out, "out")
    a.rows, a.cols, a.eps = rows, cols, eps
    _call("fame_layernorm", a, 4.0 * rows * cols)
    return out


def bert_embed(ids, word, pos, type_emb, gamma, beta, eps, seq_len, err_flag=None):
    """ids int64 [tokens] (flattened [chunks, seq_len]) -> bf16 [tokens, hidden]."""
    _cuda(ids, "ids", torch.int64)
    ids = ids.contiguous().view(-1)
    hidden = word.shape[1]
    out = torch.empty((ids.numel(), hidden), device=ids.device, dtype=torch.bfloat16)
    a = _lib.BertEmbedArgs()
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


def attn_fwd(qkv, batch, seq, heads, head_dim, key_mask=None, scale=None, out=None, algo=0):
    """qkv bf16 [batch*seq, 3*heads*head_dim] -> ctx bf16 [batch*seq, heads*head_dim]."""
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
    _call("fame_attn_fwd", a, 4.0 * batch * heads * seq * seq * head_dim)
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

"""Torch-tensor wrappers over the backward / optimizer entry points of the C ABI (see include/fame_b200.h)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from ._lib import AUX_ADD_BF16, AUX_ADD_F32, AUX_GELU_BWD_BF16, AUX_NONE, AUX_RELU_MASK_BF16, DT_BF16, DT_F32  # noqa: F401


def _dt(t):
    return DT_F32 if t.dtype == torch.float32 else DT_BF16


def _flat(name, *args):
    ops.LAUNCHES += 1
    if ops._TRACE is None:
        _lib.call_flat(name, ops._stream(), *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call_flat(name, ops._stream(), *args)
    e1.record()
    ops._TRACE.append((name, "", e0, e1, 0.0))


def _p(t):
    return None if t is None else t.data_ptr()


def gemm_ex(a, b, y, M, N, K, a_mn=False, b_mn=False, lda=None, ldb=None, ldy=None, bias=None, aux=None,
            aux_mode=AUX_NONE, ld_aux=0, act=0, alpha=1.0, nb0=1, nb1=1, sa=(0, 0), sb=(0, 0), sy=(0, 0), saux=(0, 0),
            n_valid=0, a_off=0, b_off=0, y_off=0, aux_off=0, tag="", split_k=0, drop=None, pre_act=None):
    """General tcgen05 GEMM (fame_gemm_ex).  a / b / y / aux are tensors (any shape); geometry is explicit:
    leading dimensions, element offsets and batch strides (b0, b1) in elements."""
    g = _lib.GemmExArgs()
    g.a.ptr = a.data_ptr() + 2 * a_off
    g.a.ld, g.a.stride_b0, g.a.stride_b1, g.a.mn_major = lda, sa[0], sa[1], int(a_mn)
    g.b.ptr = b.data_ptr() + 2 * b_off
    g.b.ld, g.b.stride_b0, g.b.stride_b1, g.b.mn_major = ldb, sb[0], sb[1], int(b_mn)
    g.bias = _p(bias)
    if aux is not None:
        g.aux = aux.data_ptr() + aux.element_size() * aux_off
        g.ld_aux, g.aux_stride_b0, g.aux_stride_b1, g.aux_mode = ld_aux, saux[0], saux[1], aux_mode
    g.y = y.data_ptr() + y.element_size() * y_off
    g.ldy, g.y_stride_b0, g.y_stride_b1, g.y_dtype = ldy, sy[0], sy[1], _dt(y)
    g.M, g.N, g.K, g.nb0, g.nb1 = M, N, K, nb0, nb1
    g.act, g.alpha, g.n_valid, g.split_k = act, alpha, n_valid, split_k
    ops._set_drop(g.drop, drop)
    if pre_act is not None:
        g.pre_act, g.ld_pre = pre_act.data_ptr(), pre_act.stride(0)
    skinny = M <= SKINNY_MAX_ROWS and not a_mn and not b_mn and nb0 == 1 and nb1 == 1 and K % 32 == 0 and split_k == 0 and n_valid == 0
    ops._call("fame_gemm_ex", g, 2.0 * M * N * K * nb0 * nb1, f"{'sk' if skinny else 'tc'}:{tag}:{M}x{N}x{K}")
    return y


SKINNY_MAX_ROWS = 32


def linear_dgrad(dy, w, out=None, out_dtype=torch.bfloat16, aux=None, aux_mode=AUX_NONE, wT=None, alpha=1.0, drop=None):
    """dX[T, K] = dY[T, N] @ W[N, K]  (+ aux residual gradient, or ReLU-masked by aux).  With at most 32 rows and a
    transposed shadow wT [K, N] available the product runs as a K-major x K-major skinny GEMM (weight streaming)."""
    T, N = dy.shape
    K = w.shape[1]
    if out is None:
        out = torch.empty((T, K), device=dy.device, dtype=out_dtype)
    if wT is not None and T <= SKINNY_MAX_ROWS and N % 32 == 0:
        return gemm_ex(dy, wT, out, T, K, N, lda=dy.stride(0), ldb=wT.stride(0), ldy=out.stride(0), aux=aux,
                       aux_mode=aux_mode, ld_aux=aux.stride(0) if aux is not None else 0, tag="dgrad", alpha=alpha,
                       drop=drop)
    return gemm_ex(dy, w, out, T, K, N, a_mn=False, b_mn=True, lda=dy.stride(0), ldb=w.stride(0), ldy=out.stride(0),
                   aux=aux, aux_mode=aux_mode, ld_aux=aux.stride(0) if aux is not None else 0, tag="dgrad", alpha=alpha,
                   drop=drop)


def linear_wgrad(dy, x, out, accumulate=True, dbias=None):
    """dW[N, K] (+)= dY[T, N]^T @ X[T, K] -> f32 `out` (a view into the flat gradient buffer).  With accumulate (the
    training step: the buffer was zeroed by zero_grad) the token contraction is split over the SMs (split-K) and the
    slices are added with float4 atomics; otherwise `out` is overwritten by a single-pass product."""
    T, N = dy.shape
    K = x.shape[1]
    if T <= SKINNY_MAX_ROWS:
        _flat("fame_wgrad_small", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0),
              T, N, K, int(accumulate), _p(dbias))
        return out
    if dbias is not None:
        raise ValueError("the fused bias gradient exists on the <= 32-row path only")
    return gemm_ex(dy, x, out, N, K, T, a_mn=True, b_mn=True, lda=dy.stride(0), ldb=x.stride(0), ldy=out.stride(0),
                   tag="wgrad", split_k=-1 if accumulate else 0)


def layernorm_bwd(x, dy, stats, gamma, dgamma, dbeta, want_bf16=True, want_f32=False):
    return layernorm_bwd_drop(x, dy, stats, gamma, dgamma, dbeta, want_bf16, want_f32, None)[:2]


def layernorm_bwd_drop(x, dy, stats, gamma, dgamma, dbeta, want_bf16=True, want_f32=False, drop=None, residual=None):
    """Returns (dx bf16 | None, dx f32 | None, dx_drop bf16 | None): with `drop` (a _lib.DropoutCfg) dx_drop is dx
    with that dropout mask re-applied -- the gradient of the layer under the dropout.  residual (bf16, optional): the
    forward LayerNorm ran on x + residual (ops.layernorm(..., residual=))."""
    rows, cols = x.shape
    dxb = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    dxf = torch.empty((rows, cols), device=x.device, dtype=torch.float32) if want_f32 else None
    on = drop is not None and drop.thresh16 > 0
    dxd = torch.empty((rows, cols), device=x.device, dtype=torch.bfloat16) if on else None
    _flat("fame_layernorm_bwd", x.data_ptr(), _dt(x), dy.data_ptr(), _dt(dy), stats.data_ptr(), gamma.data_ptr(),
          _p(dxb), _p(dxf), _p(dgamma), _p(dbeta), rows, cols, _p(dxd), ctypes.addressof(drop) if on else None,
          _p(residual))
    return dxb, dxf, dxd


def dropout_apply(x, drop):
    """In-place dropout of a 2-D bf16 / f32 tensor with the mask of site `drop` (no-op when thresh16 == 0)."""
    if drop is None or drop.thresh16 == 0:
        return x
    rows, cols = x.shape
    _flat("fame_dropout_apply", x.data_ptr(), _dt(x), x.stride(0), rows, cols, ctypes.addressof(drop))
    return x


def gelu_fwd(pre):
    h = torch.empty_like(pre)
    _flat("fame_gelu_fwd", pre.data_ptr(), h.data_ptr(), pre.numel())
    return h


def gelu_bwd(pre, dh):
    d = torch.empty_like(pre)
    _flat("fame_gelu_bwd", pre.data_ptr(), dh.data_ptr(), d.data_ptr(), pre.numel())
    return d


def colsum(x, out):
    rows, cols = x.shape
    _flat("fame_colsum", x.data_ptr(), _dt(x), x.stride(0), rows, cols, out.data_ptr())


def seq_mean_bwd(dout, batch, L):
    cols = dout.shape[1]
    dx = torch.empty((batch * L, cols), device=dout.device, dtype=torch.bfloat16)
    _flat("fame_seq_mean_bwd", dout.data_ptr(), dx.data_ptr(), batch, L, cols)
    return dx


def lab_embed_bwd(dx, lab, dpos, dw, dbias):
    B, L = lab.shape
    _flat("fame_lab_embed_bwd", dx.data_ptr(), lab.data_ptr(), dpos.data_ptr(), dw.data_ptr(), dbias.data_ptr(), B, L,
          dpos.shape[1])


def attn_bwd_softmax(s, dp, rows, seq, ld, scale):
    p = torch.empty((rows, ld), device=s.device, dtype=torch.bfloat16)
    ds = torch.empty((rows, ld), device=s.device, dtype=torch.bfloat16)
    _flat("fame_attn_bwd_softmax", s.data_ptr(), dp.data_ptr(), p.data_ptr(), ds.data_ptr(), rows, seq, ld, float(scale))
    return p, ds


def attn_delta(dctx, ctx, batch, seq, heads, head_dim):
    """delta f32 [batch, heads, seq] = rowsum(dO * O) per (token, head)."""
    if dctx.stride(0) != ctx.stride(0):
        raise ValueError("attn_delta: dO and O must share one row stride")
    delta = torch.empty((batch, heads, seq), device=ctx.device, dtype=torch.float32)
    _flat("fame_attn_delta", dctx.data_ptr(), ctx.data_ptr(), ctx.stride(0), delta.data_ptr(), batch, seq, heads, head_dim)
    return delta


def attn_bwd_pds(qkv, dctx, lse, delta, batch, seq, heads, head_dim, ldp, scale, drop=None):
    """P and dS (bf16 [batch*heads*seq, ldp]) from the packed qkv, dO, the forward's lse and delta; scores stay in TMEM."""
    rows = batch * heads * seq
    p = torch.empty((rows, ldp), device=qkv.device, dtype=torch.bfloat16)
    ds = torch.empty((rows, ldp), device=qkv.device, dtype=torch.bfloat16)
    _flat("fame_attn_bwd_pds", qkv.data_ptr(), qkv.stride(0), dctx.data_ptr(), dctx.stride(0), lse.data_ptr(),
          delta.data_ptr(), p.data_ptr(), ds.data_ptr(), ldp, batch, seq, heads, head_dim, float(scale),
          ctypes.addressof(drop) if drop is not None and drop.thresh16 > 0 else None)
    return p, ds


def attn_bwd_fused(qkv, dctx, lse, delta, batch, seq, heads, head_dim, scale, drop=None):
    """dqkv (bf16 [batch*seq, 3*heads*head_dim]: dQ | dK | dV) from the packed qkv, dO, the forward's lse and delta; the
    probabilities and score gradients stay in TMEM (two launches: dK/dV pass, dQ pass)."""
    dqkv = torch.empty((batch * seq, 3 * heads * head_dim), device=qkv.device, dtype=torch.bfloat16)
    _flat("fame_attn_bwd_fused", qkv.data_ptr(), qkv.stride(0), dctx.data_ptr(), dctx.stride(0), lse.data_ptr(),
          delta.data_ptr(), dqkv.data_ptr(), dqkv.stride(0), batch, seq, heads, head_dim, float(scale),
          ctypes.addressof(drop) if drop is not None and drop.thresh16 > 0 else None)
    return dqkv


def bert_embed_bwd(d_sum, ids, dword, dpos, dtype0, seq, pad_idx=0):
    tokens, hidden = d_sum.shape
    _flat("fame_bert_embed_bwd", d_sum.data_ptr(), ids.data_ptr(), dword.data_ptr(), dpos.data_ptr(), dtype0.data_ptr(),
          tokens, seq, hidden, dword.shape[0], pad_idx)


def demo_add_bwd(dout, ids, dtables):
    B, hidden = dout.shape
    _flat("fame_demo_add_bwd", dout.data_ptr(), ids[0].data_ptr(), ids[1].data_ptr(), ids[2].data_ptr(),
          ids[3].data_ptr(), dtables[0].data_ptr(), dtables[1].data_ptr(), dtables[2].data_ptr(), dtables[3].data_ptr(),
          dtables[0].shape[0], dtables[1].shape[0], dtables[2].shape[0], dtables[3].shape[0], B, hidden)


def sgemm(a, sam, sak, b, sbk, sbn, c, M, N, K, alpha=1.0, accumulate=False):
    """C[M,N] = alpha * sum_k A(m,k) B(k,n) (+C) in fp32 with explicit element strides."""
    _flat("fame_sgemm_small", a.data_ptr(), sam, sak, b.data_ptr(), sbk, sbn, c.data_ptr(), c.stride(0), M, N, K,
          float(alpha), int(accumulate))
    return c


def fusion_bwd_hidden(dlogits, w4, pre):
    B = dlogits.shape[0]
    dhid = torch.empty((B, 512), device=dlogits.device, dtype=torch.float32)
    _flat("fame_fusion_bwd_hidden", dlogits.data_ptr(), w4.data_ptr(), pre.data_ptr(), dhid.data_ptr(), B)
    return dhid


def fusion_bwd_gate(dgated, proj, sig_w, w_mod, lambda_l1, dsig, w_mod_dev=None):
    B = dgated.shape[0]
    dproj = torch.empty((B, 768), device=dgated.device, dtype=torch.float32)
    _flat("fame_fusion_bwd_gate", dgated.data_ptr(), proj.data_ptr(), sig_w.data_ptr(), float(w_mod[0]), float(w_mod[1]),
          float(w_mod[2]), float(lambda_l1), dproj.data_ptr(), dsig.data_ptr(), B, _p(w_mod_dev))
    return dproj


def grad_sumsq(g, out):
    _flat("fame_grad_sumsq", g.data_ptr(), g.numel(), out.data_ptr())


def clip_adamw(p, g, m, v, sumsq, max_norm, lr, beta1, beta2, eps, weight_decay, step, grad_norm_out=None,
               step_dev=None, hyper_dev=None, p_bf16=None):
    _flat("fame_clip_adamw", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), sumsq.data_ptr(),
          float(max_norm), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
          _p(grad_norm_out), _p(step_dev), _p(hyper_dev), _p(p_bf16))


def decay_only(p, lr, weight_decay, hyper_dev=None, p_bf16=None):
    """AdamW step of a range whose gradient and moments are identically zero: p <- p (1 - lr wd)."""
    _flat("fame_decay_only", p.data_ptr(), p.numel(), float(lr), float(weight_decay), _p(hyper_dev), _p(p_bf16))


def transpose_bf16_table(table, n_entries, total_tiles):
    _flat("fame_transpose_bf16_table", table.data_ptr(), n_entries, total_tiles)


def cast_bf16(x, y):
    _flat("fame_cast_bf16", x.data_ptr(), y.data_ptr(), x.numel())


def focal_loss_fwd_bwd(logits, labels, pos_weight, gamma=2.0, alpha=1.0, want_grad=True, batch_total=None):
    """sum_i mean_b FocalLoss(gamma, alpha, pos_weight_i)(logits[:, i], labels[:, i]) -> (loss f64 [1], dlogits | None).
    batch_total (device int64 [1], optional): global batch size of a data-parallel step (the mean runs over it)."""
    B = logits.shape[0]
    loss = torch.zeros(1, device=logits.device, dtype=torch.float64)
    dl = torch.empty((B, 3), device=logits.device, dtype=torch.float32) if want_grad else None
    keep = (logits.float().contiguous(), labels.float().contiguous(), pos_weight.float().contiguous())
    _flat("fame_focal_loss_fwd_bwd", keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), float(gamma), float(alpha),
          B, loss.data_ptr(), _p(dl), _p(batch_total))
    return loss, dl


def relu_(x):
    _flat("fame_relu_fwd", x.data_ptr(), x.numel())
    return x


def relu_bwd_(dh, pre):
    _flat("fame_relu_bwd", dh.data_ptr(), pre.data_ptr(), dh.numel())
    return dh


def linear_gelu_small(x, w, bias):
    """<= 32 rows: h = gelu(x W^T + b) and the bf16 pre-activation from ONE weight-streaming launch.  Returns (pre, h)."""
    M, K = x.shape
    N = w.shape[0]
    pre = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    h = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    gemm_ex(x, w, h, M, N, K, lda=x.stride(0), ldb=w.stride(0), ldy=N, bias=bias, act=_lib.ACT_GELU_ERF, pre_act=pre,
            tag="fwd")
    return pre, h

"""Unit parity of each CUDA kernel, called through the C ABI, against a plain PyTorch fp32 statement of the op."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _close(got, ref, rel):
    got, ref = got.float(), ref.float()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= rel * (ref.abs().max().item() + 1e-12)


@pytest.mark.parametrize("M,N,K,bias,res,act,f32", [
    (128, 256, 64, False, False, 0, False), (256, 512, 768, True, False, 0, False),
    (200, 264, 72, True, True, 0, False), (1000, 768, 3072, True, True, 0, False),
    (777, 3072, 768, True, False, 1, False), (512, 2048, 768, True, False, 2, False),
    (33, 256, 768, True, False, 0, True), (1, 8, 8, False, False, 0, True),
    (148 * 128 * 2 + 5, 768, 768, True, True, 0, False),
])
def test_gemm(M, N, K, bias, res, act, f32):
    from fairmultimodal_b200 import ops
    torch.manual_seed(M + N + K)
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda") if bias else None
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    y = ops.gemm_bias_act(x, w, b, r, act, out_dtype=torch.float32 if f32 else torch.bfloat16)
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b
    ref = torch.nn.functional.gelu(ref) if act == 1 else torch.relu(ref) if act == 2 else ref
    if res:
        ref = ref + r.float()
    _close(y, ref, 1e-5 if f32 else 1e-2)        # bf16 output rounding: 2^-8 relative


@pytest.mark.parametrize("M,N,K,bias,res,act,f32", [
    (32, 768, 768, True, "f32", 0, True), (32, 3072, 768, True, None, 1, False), (32, 768, 3072, True, "f32", 0, True),
    (32, 768, 768, True, None, 0, False), (17, 256, 96, False, "bf16", 2, False), (1, 8, 32, True, None, 0, True),
    (8, 2304, 768, True, None, 0, False), (32, 3072, 768, False, "mask", 0, False),
])
def test_skinny_gemm(M, N, K, bias, res, act, f32):
    """<= 32 rows: the weight-streaming mma.sync path (skinny_gemm.cuh), same epilogues as the tcgen05 kernel."""
    from fairmultimodal_b200 import ops
    from fairmultimodal_b200 import ops_train as T
    torch.manual_seed(M * 7 + N + K)
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda") if bias else None
    out_dtype = torch.float32 if f32 else torch.bfloat16
    ref = x.float() @ w.float().t()
    if bias:
        ref = ref + b
    ref = torch.nn.functional.gelu(ref) if act == 1 else torch.relu(ref) if act == 2 else ref
    if res == "mask":
        aux = torch.randn(M, N, device="cuda").bfloat16()
        y = torch.empty(M, N, device="cuda", dtype=out_dtype)
        T.gemm_ex(x, w, y, M, N, K, lda=K, ldb=K, ldy=N, bias=b, aux=aux, aux_mode=T.AUX_RELU_MASK_BF16, ld_aux=N, act=act)
        ref = ref * (aux.float() > 0)
    else:
        r = None if res is None else (torch.randn(M, N, device="cuda") if res == "f32" else torch.randn(M, N, device="cuda").bfloat16())
        y = ops.gemm_bias_act(x, w, b, r, act, out_dtype=out_dtype)
        if r is not None:
            ref = ref + r.float()
    _close(y, ref, 1e-5 if f32 else 1e-2)


def test_skinny_dgrad_through_transposed_shadow():
    from fairmultimodal_b200 import ops_train as T
    torch.manual_seed(3)
    dy = (torch.randn(32, 3072, device="cuda") * 0.1).bfloat16()
    w = (torch.randn(3072, 768, device="cuda") * 0.05).bfloat16()
    ref = dy.float() @ w.float()
    a = T.linear_dgrad(dy, w, out_dtype=torch.float32)                       # tcgen05, MN-major W
    b = T.linear_dgrad(dy, w, out_dtype=torch.float32, wT=w.t().contiguous())  # skinny, K-major W^T
    _close(a, ref, 1e-5)
    _close(b, ref, 1e-5)


@pytest.mark.parametrize("M,N,K", [(32, 768, 3072), (32, 3072, 768), (5, 72, 200), (1, 8, 8), (32, 100, 36)])
def test_wgrad_small(M, N, K):
    from fairmultimodal_b200 import ops_train as T
    torch.manual_seed(M + N + K)
    dy = (torch.randn(M, N, device="cuda") * 0.1).bfloat16()
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    ref = dy.float().t() @ x.float()
    out = torch.full((N, K), 7.0, device="cuda")
    T.linear_wgrad(dy, x, out, accumulate=False)
    _close(out, ref, 1e-5)
    T.linear_wgrad(dy, x, out, accumulate=True)
    _close(out, 2 * ref, 1e-5)
    # fused bias gradient (column sums of dY) from the same launch: overwrite, then accumulate
    db = torch.full((N,), 3.0, device="cuda")
    T.linear_wgrad(dy, x, out, accumulate=False, dbias=db)
    _close(out, ref, 1e-5)
    _close(db, dy.float().sum(0), 1e-5)
    T.linear_wgrad(dy, x, out, accumulate=True, dbias=db)
    _close(db, 2 * dy.float().sum(0), 1e-5)


def test_transpose_table():
    import numpy as np
    from fairmultimodal_b200 import ops_train as T
    torch.manual_seed(4)
    mats = [torch.randn(r, c, device="cuda").bfloat16() for r, c in ((768, 768), (3072, 768), (100, 72), (64, 200))]
    outs = [torch.zeros(m.shape[1], m.shape[0], device="cuda", dtype=torch.bfloat16) for m in mats]
    rec = np.zeros(len(mats), dtype=np.dtype([("src", "<u8"), ("dst", "<u8"), ("rows", "<i4"), ("cols", "<i4"),
                                              ("tile0", "<i4"), ("tiles_x", "<i4")]))
    tiles = 0
    for i, (m, o) in enumerate(zip(mats, outs)):
        tx, ty = (m.shape[1] + 63) // 64, (m.shape[0] + 63) // 64
        rec[i] = (m.data_ptr(), o.data_ptr(), m.shape[0], m.shape[1], tiles, tx)
        tiles += tx * ty
    table = torch.from_numpy(rec.view(np.uint8).copy()).cuda()
    T.transpose_bf16_table(table, len(mats), tiles)
    for m, o in zip(mats, outs):
        assert torch.equal(o, m.t().contiguous())


def test_gemm_rejects_bad_shapes():
    from fairmultimodal_b200 import _lib, ops
    x = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.FameError):
        ops.gemm_bias_act(x, w)                    # K % 8 != 0


@pytest.mark.parametrize("H,D,B,S,masked", [
    # BERT heads (head_dim 64): the note encoder's hot shape (12 x 64, 512 tokens, padded chunks) and odd lengths
    (12, 64, 1, 128, False), (12, 64, 2, 512, False), (12, 64, 3, 512, True), (12, 64, 2, 300, True),
    (12, 64, 2, 77, False), (12, 64, 2, 1, False), (12, 64, 2, 1000, True), (12, 64, 4, 256, True),
    (12, 64, 40, 512, True),                       # more work items than one wave of 148 persistent CTAs
    # lab-encoder heads (head_dim 96; L = 542 is the real cohort's token count: 5 key blocks with a ragged tail)
    (8, 96, 3, 542, False), (8, 96, 2, 542, True), (8, 96, 2, 12, False), (8, 96, 1, 1300, False),
    (8, 96, 5, 128, False), (8, 96, 40, 542, False),
])
def test_attention(H, D, B, S, masked):
    """attn_fwd_pair_kernel (the only forward attention kernel) against fp32 torch softmax(QK^T / sqrt(d)) V, with
    the saved log-sum-exp checked as well."""
    from fairmultimodal_b200 import ops
    torch.manual_seed(S + B)
    qkv = torch.randn(B * S, 3 * H * D, device="cuda").bfloat16()
    mask = kv_len = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), device="cuda")
        mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).to(torch.uint8).contiguous()
        kv_len = ops.mask_kv_len(mask)                  # as bert.encode passes it on the hot path
    lse = torch.empty(B, H, S, device="cuda")
    ctx = ops.attn_fwd(qkv, B, S, H, D, key_mask=mask, kv_len=kv_len, lse=lse)
    q, k, v = qkv.float().view(B, S, 3, H, D).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * D ** -0.5
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H * D)
    _close(ctx, ref, 2e-2)
    assert (lse - torch.logsumexp(s, -1) * 1.4426950408889634).abs().max().item() <= 2e-2


def test_attention_rejects_removed_algos():
    from fairmultimodal_b200 import _lib, ops
    qkv = torch.zeros(128, 3 * 12 * 64, device="cuda", dtype=torch.bfloat16)
    for algo in (1, 2):
        with pytest.raises(_lib.FameError):
            ops.attn_fwd(qkv, 1, 128, 12, 64, algo=algo)


@pytest.mark.parametrize("H,D,B,S", [(12, 64, 40, 512), (8, 96, 7, 542), (12, 64, 3, 700), (12, 64, 200, 512)])
def test_attention_kv_len_skips_only_masked_blocks(H, D, B, S):
    """Key blocks beyond the last attended key are skipped: the context is bit-identical to the unskipped kernel
    (masked keys have probability exactly 0 either way), for prefix masks, masks with holes and all-masked rows."""
    from fairmultimodal_b200 import ops
    torch.manual_seed(S + B)
    qkv = torch.randn(B * S, 3 * H * D, device="cuda").bfloat16()
    lens = torch.randint(1, S + 1, (B,), device="cuda")
    lens[0] = S
    lens[1] = 1
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None])
    mask[2, : S // 2 : 3] = False                       # holes inside the attended prefix
    if B > 4:
        mask[3] = False                                  # nothing attended at all
    mask = mask.to(torch.uint8).contiguous()
    kv_len = ops.mask_kv_len(mask)
    last = torch.where(mask.bool(), torch.arange(1, S + 1, device="cuda")[None, :], 0).max(dim=1).values
    assert torch.equal(kv_len.long().abs(), last)        # integer indexing: bit-exact
    prefix = mask.long().sum(dim=1) == last               # sign: prefix mask (validity from the length) or a mask with holes
    assert torch.equal(kv_len >= 0, prefix) and not bool(prefix[2]) and bool(prefix[0])
    lse0 = torch.empty(B, H, S, device="cuda")
    lse1 = torch.empty(B, H, S, device="cuda")
    full = ops.attn_fwd(qkv, B, S, H, D, key_mask=mask, lse=lse0)
    skip = ops.attn_fwd(qkv, B, S, H, D, key_mask=mask, kv_len=kv_len, lse=lse1)
    assert torch.equal(full, skip)
    assert torch.equal(lse0, lse1)


@pytest.mark.parametrize("rows,cols", [(1003, 768), (1, 768), (16, 8), (130, 1024), (257, 520)])
def test_layernorm_bf16_fast_path(rows, cols):
    """bf16 -> bf16 LayerNorm (two rows per warp): odd row counts, narrow and 1024-wide rows, in place, statistics."""
    from fairmultimodal_b200 import ops
    torch.manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device="cuda") * 2 + 0.5).bfloat16()
    g, b = torch.randn(cols, device="cuda"), torch.randn(cols, device="cuda")
    ref = torch.nn.functional.layer_norm(x.float(), (cols,), g, b, 1e-12)
    stats = torch.empty(rows, 2, device="cuda")
    _close(ops.layernorm(x, g, b, 1e-12, stats=stats), ref, 1e-2)
    xf = x.float()
    assert torch.allclose(stats[:, 0], xf.mean(1), atol=1e-5, rtol=1e-5)
    assert torch.allclose(stats[:, 1], (xf.var(1, unbiased=False) + 1e-12).rsqrt(), rtol=1e-4)
    y = x.clone()
    ops.layernorm(y, g, b, 1e-12, out=y)                 # in place, as the note encoder calls it
    _close(y, ref, 1e-2)
    # fused residual add: LN(bf16(x + r)), in place over x
    r = torch.randn(rows, cols, device="cuda").bfloat16()
    ref_r = torch.nn.functional.layer_norm((x.float() + r.float()).bfloat16().float(), (cols,), g, b, 1e-12)
    y = x.clone()
    ops.layernorm(y, g, b, 1e-12, out=y, residual=r)
    _close(y, ref_r, 1e-2)


def test_layernorm_and_embed():
    from fairmultimodal_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(1003, 768, device="cuda").bfloat16()
    g, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
    for eps in (1e-12, 1e-5):
        _close(ops.layernorm(x, g, b, eps), torch.nn.functional.layer_norm(x.float(), (768,), g, b, eps), 1e-2)
    V, S = 1000, 64
    word, pos, typ = (torch.randn(n, 768, device="cuda") * 0.02 for n in (V, 512, 2))
    ids = torch.randint(0, V, (5, S), device="cuda")
    y = ops.bert_embed(ids, word, pos, typ[0].contiguous(), g, b, 1e-12, S)
    ref = torch.nn.functional.layer_norm(word[ids] + typ[0] + pos[None, :S], (768,), g, b, 1e-12).view(-1, 768)
    _close(y, ref, 1e-2)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.bert_embed(torch.full((1, S), V + 3, device="cuda"), word, pos, typ[0].contiguous(), g, b, 1e-12, S, err_flag=flag)
    assert flag.item() == 1                       # out-of-range ids are reported, not silently read


@pytest.mark.parametrize("mode", ["mean", "max"])
def test_segment_reduce_bit_exact(mode):
    """Chunk->patient indexing and the f32 mean are bit-exact against numpy (empty, single, ragged segments)."""
    from fairmultimodal_b200 import ops
    rng = np.random.default_rng(0)
    counts = np.array([4, 0, 1, 16, 3, 7, 0, 2, 33], dtype=np.int32)
    offs = np.zeros(len(counts) + 1, dtype=np.int32)
    offs[1:] = counts.cumsum()
    x = rng.standard_normal((int(offs[-1]), 768)).astype(np.float32)
    out = ops.segment_mean(torch.from_numpy(x).cuda(), torch.from_numpy(offs).cuda(), mode=mode).cpu().numpy()
    f = np.mean if mode == "mean" else np.max
    ref = np.stack([f(x[offs[i]:offs[i + 1]], axis=0) if counts[i] else np.zeros(768, np.float32) for i in range(len(counts))])
    np.testing.assert_array_equal(out, ref)
    # CLS gather folded into the row stride (bf16 hidden state, seq 16)
    hs = torch.randn(int(offs[-1]) * 16, 768, device="cuda").bfloat16()
    out = ops.segment_mean(hs, torch.from_numpy(offs).cuda(), cols=768, ldx=16 * 768, mode=mode).cpu().numpy()
    cls = hs.view(-1, 16, 768)[:, 0].float().cpu().numpy()
    ref = np.stack([f(cls[offs[i]:offs[i + 1]], axis=0) if counts[i] else np.zeros(768, np.float32) for i in range(len(counts))])
    np.testing.assert_array_equal(out, ref)


@pytest.mark.parametrize("B,S,masked", [(3, 512, True), (2, 512, False), (5, 77, True), (2, 1, False), (260, 128, True)])
def test_attention_cls_query(B, S, masked):
    """fame_attn_cls (one query per sequence, K / V of all tokens; the note encoder's last layer) against fp32 torch,
    including an all-masked sequence (zero row) and strided q (CLS rows of a [B*S, H] tensor)."""
    from fairmultimodal_b200 import ops
    H, D = 12, 64
    torch.manual_seed(S + B)
    x = torch.randn(B * S, H * D, device="cuda").bfloat16()              # queries are rows 0, S, 2S, ... of this
    kv = torch.randn(B * S, 2 * H * D, device="cuda").bfloat16()
    mask = None
    if masked:
        lens = torch.randint(1, S + 1, (B,), device="cuda")
        mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None])
        if B > 2:
            mask[1] = False                                                # nothing attended: zero row
        mask = mask.to(torch.uint8).contiguous()
    q = x.view(B, S, H * D)[:, 0, :]
    ctx = ops.attn_cls(q, kv, B, S, H, D, 0, H * D, key_mask=mask)
    qf = q.float().view(B, H, 1, D)
    k = kv[:, :H * D].float().view(B, S, H, D).permute(0, 2, 1, 3)
    v = kv[:, H * D:].float().view(B, S, H, D).permute(0, 2, 1, 3)
    s = (qf @ k.transpose(-1, -2)) * D ** -0.5
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    pr = torch.softmax(s, -1)
    pr = torch.where(torch.isnan(pr), torch.zeros_like(pr), pr)
    ref = (pr @ v).reshape(B, H * D)
    _close(ctx, ref, 1e-2)
    if masked and B > 2:
        assert ctx[1].abs().max().item() == 0.0

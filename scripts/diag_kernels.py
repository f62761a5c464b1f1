"""GPU bring-up diagnostics: run one kernel family per process and print error patterns, not just pass/fail.

Usage (on a B200):  python scripts/diag_kernels.py <gemm|attn|rowwise|all>
Each family runs in its own subprocess under a timeout so a trap in one does not poison the others.
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _err_report(name, got, ref, tol):
    import torch

    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    mx = err.max().item()
    bad = err > tol * scale
    print(f"[{name}] max_abs_err={mx:.4e} ref_max={scale:.4e} rel={mx / scale:.3e} bad={int(bad.sum())}/{bad.numel()}",
          flush=True)
    ok = mx <= tol * scale and torch.isfinite(got).all().item()
    if not ok:
        rows = bad.any(dim=1).nonzero().flatten().tolist()
        cols = bad.any(dim=0).nonzero().flatten().tolist()
        print(f"   bad rows ({len(rows)}): {rows[:24]}{'...' if len(rows) > 24 else ''}")
        print(f"   bad cols ({len(cols)}): {cols[:24]}{'...' if len(cols) > 24 else ''}")
        print("   got[0:4,0:8] =", got[0:4, 0:8].tolist())
        print("   ref[0:4,0:8] =", ref[0:4, 0:8].tolist())
        print("   nan/inf in got:", int((~torch.isfinite(got)).sum()))
    return ok


def run_gemm():
    import torch
    from fairmultimodal_b200 import ops

    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    cases = [
        # M, N, K, bias, res, act, f32out
        (128, 256, 64, False, False, 0, False),
        (128, 256, 128, False, False, 0, False),
        (128, 256, 768, True, False, 0, False),
        (256, 512, 768, True, False, 0, False),
        (200, 264, 72, True, True, 0, False),
        (1000, 768, 3072, True, True, 0, False),
        (777, 3072, 768, True, False, 1, False),
        (512, 2048, 768, True, False, 2, False),
        (33, 256, 768, True, False, 0, True),
        (4096, 2304, 768, True, False, 0, False),
        (148 * 128 * 3 + 5, 768, 768, True, True, 0, False),
    ]
    for (M, N, K, hb, hr, act, f32) in cases:
        x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev) if hb else None
        r = torch.randn(M, N, device=dev).bfloat16() if hr else None
        y = ops.gemm_bias_act(x, w, b, r, act, out_dtype=torch.float32 if f32 else torch.bfloat16)
        torch.cuda.synchronize()
        ref = x.float() @ w.float().t()
        if hb:
            ref = ref + b
        if act == 1:
            ref = torch.nn.functional.gelu(ref)
        elif act == 2:
            ref = torch.relu(ref)
        if hr:
            ref = ref + r.float()
        ok &= _err_report(f"gemm M{M} N{N} K{K} b{int(hb)} r{int(hr)} act{act} f32{int(f32)}", y, ref, 1e-2)
    # timing of the note-encoder shapes
    for (M, N, K, act) in [(131072, 2304, 768, 0), (131072, 768, 768, 0), (131072, 3072, 768, 1), (131072, 768, 3072, 0)]:
        x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm_bias_act(x, w, b, None, act, out=y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.gemm_bias_act(x, w, b, None, act, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[gemm-time] M{M} N{N} K{K} act{act}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(x, w)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[cublas-time] M{M} N{N} K{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def run_attn():
    import torch
    from fairmultimodal_b200 import ops

    torch.manual_seed(1)
    dev = "cuda"
    ok = True
    for (H, D, B, S, masked, algo, mag) in [
        (12, 64, 1, 128, False, 3, 1.0), (12, 64, 2, 512, False, 3, 1.0), (12, 64, 3, 512, True, 3, 1.0),
        (12, 64, 2, 300, True, 3, 1.0), (12, 64, 2, 77, False, 3, 1.0), (12, 64, 2, 1000, True, 3, 1.0),
        (8, 96, 2, 128, False, 3, 1.0), (8, 96, 3, 542, False, 3, 1.0), (8, 96, 2, 542, True, 3, 1.0),
        (8, 96, 2, 12, False, 3, 1.0), (8, 96, 1, 1300, False, 3, 1.0),
        # more work items than SMs (several items per persistent CTA), and large scores (row maxima grow by > 2^8
        # between key blocks: exercises the in-TMEM rescale of the accumulator)
        (12, 64, 40, 512, True, 3, 1.0), (8, 96, 40, 542, False, 3, 1.0), (12, 64, 4, 512, False, 3, 3.0),
        (8, 96, 3, 542, True, 3, 3.0), (12, 64, 40, 640, False, 3, 2.5),
    ]:
        qkv = (torch.randn(B * S, 3 * H * D, device=dev) * mag).bfloat16()
        mask = None
        if masked:
            lens = torch.randint(1, S + 1, (B,), device=dev)
            mask = (torch.arange(S, device=dev)[None, :] < lens[:, None]).to(torch.uint8).contiguous()
        ctx = ops.attn_fwd(qkv, B, S, H, D, key_mask=mask, algo=algo)
        torch.cuda.synchronize()
        q, k, v = qkv.float().view(B, S, 3, H, D).permute(2, 0, 3, 1, 4)
        s = (q @ k.transpose(-1, -2)) * D ** -0.5
        if mask is not None:
            s = s.masked_fill(mask[:, None, None, :] == 0, float("-inf"))
        ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H * D)
        ok &= _err_report(f"attn algo{algo} H{H} D{D} B{B} S{S} masked{int(masked)}", ctx, ref, 2e-2)
    for (H, D, B, S, algo) in [(12, 64, 256, 512, 3), (8, 96, 256, 542, 3), (8, 96, 32, 542, 3)]:
        qkv = torch.randn(B * S, 3 * H * D, device=dev).bfloat16()
        out = torch.empty(B * S, H * D, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.attn_fwd(qkv, B, S, H, D, out=out, algo=algo)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.attn_fwd(qkv, B, S, H, D, out=out, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[attn-time] algo{algo} H{H} D{D} B{B} S{S}: {ms:.3f} ms  {4.0 * B * H * S * S * D / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def run_rowwise():
    import torch
    from fairmultimodal_b200 import ops

    torch.manual_seed(2)
    dev = "cuda"
    ok = True
    x = torch.randn(1003, 768, device=dev).bfloat16()
    g, b = torch.randn(768, device=dev), torch.randn(768, device=dev)
    for eps in (1e-12, 1e-5):
        y = ops.layernorm(x, g, b, eps)
        ref = torch.nn.functional.layer_norm(x.float(), (768,), g, b, eps)
        ok &= _err_report(f"layernorm eps{eps}", y, ref, 1e-2)
    V, S, Cn = 28996, 512, 5
    word, pos, typ = torch.randn(V, 768, device=dev) * 0.02, torch.randn(512, 768, device=dev) * 0.02, torch.randn(2, 768, device=dev) * 0.02
    ids = torch.randint(0, V, (Cn, S), device=dev)
    y = ops.bert_embed(ids, word, pos, typ[0].contiguous(), g, b, 1e-12, S)
    e = word[ids] + typ[0] + pos[None, :S]
    ref = torch.nn.functional.layer_norm(e, (768,), g, b, 1e-12).view(-1, 768)
    ok &= _err_report("bert_embed", y, ref, 1e-2)
    # segment mean: f32 bit-exactness vs sequential sum / n, plus CLS-strided bf16
    counts = torch.tensor([4, 0, 1, 16, 3, 7, 0, 2], dtype=torch.int32)
    offs = torch.zeros(len(counts) + 1, dtype=torch.int32)
    offs[1:] = counts.cumsum(0)
    Ctot = int(offs[-1])
    xs = torch.randn(Ctot, 768, device=dev)
    out = ops.segment_mean(xs, offs.to(dev))
    import numpy as np
    xn = xs.cpu().numpy()
    refn = np.stack([xn[offs[i]:offs[i + 1]].mean(axis=0) if counts[i] > 0 else np.zeros(768, np.float32) for i in range(len(counts))])
    exact = np.array_equal(out.cpu().numpy(), refn.astype(np.float32))
    print(f"[segment_mean f32] bit-exact vs numpy: {exact}  maxdiff={np.abs(out.cpu().numpy() - refn).max():.3e}")
    ok &= bool(np.abs(out.cpu().numpy() - refn).max() < 1e-6)
    hs = torch.randn(Ctot * 16, 768, device=dev).bfloat16()  # pretend seq_len 16: CLS rows are every 16th row
    out = ops.segment_mean(hs, offs.to(dev), cols=768, ldx=16 * 768)
    cls = hs.view(Ctot, 16, 768)[:, 0].float().cpu().numpy()
    refn = np.stack([cls[offs[i]:offs[i + 1]].mean(axis=0) if counts[i] > 0 else np.zeros(768, np.float32) for i in range(len(counts))])
    ok &= bool(np.abs(out.cpu().numpy() - refn).max() < 1e-5)
    print(f"[segment_mean bf16 strided] maxdiff={np.abs(out.cpu().numpy() - refn).max():.3e}")
    return ok


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        rc = 0
        for fam in ("rowwise", "gemm", "attn"):
            print(f"===== {fam} =====", flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), fam], timeout=420)
                print(f"===== {fam} exit {r.returncode} =====", flush=True)
                rc |= r.returncode
            except subprocess.TimeoutExpired:
                print(f"===== {fam} TIMEOUT =====", flush=True)
                rc |= 1
        sys.exit(rc)
    fn = {"gemm": run_gemm, "attn": run_attn, "rowwise": run_rowwise}[what]
    sys.exit(0 if fn() else 1)

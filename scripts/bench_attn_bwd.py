"""Attention backward of one lab-tower layer (8 heads x 96, L = 542): fused two-pass kernel vs the P / dS kernel + three batched
GEMMs, with and without attention dropout, at 32 and 1024 patients.  CUDA events, L2 flushed between iterations."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.build()
from fairmultimodal_b200 import _lib, ops, train
from fairmultimodal_b200 import ops_train as T

def drop_cfg(p, step):
    c = _lib.DropoutCfg(); c.step = step.data_ptr(); c.seed = 1234; c.thresh16 = int(round(p * 65536)); c.group_shift = 0
    return c

out = {}
step = torch.zeros(1, device="cuda", dtype=torch.int32)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for B in (32, 1024):
    L, nh, D = 542, 8, 96
    qkv = (torch.randn(B * L, 3 * nh * D, device="cuda") * 0.7).bfloat16()
    dctx = (torch.randn(B * L, nh * D, device="cuda") * 0.1).bfloat16()
    for p in (0.0, 0.1):
        c = drop_cfg(p, step) if p > 0 else None
        lse = torch.empty(B, nh, L, device="cuda")
        ctx = ops.attn_fwd(qkv, B, L, nh, D, lse=lse, drop=c)
        for fused in (True, False):
            for _ in range(2):
                train._attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, drop=c, fused=fused)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); train._attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, drop=c, fused=fused); e1.record()
                torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            flop = 5 * 2.0 * B * nh * L * L * D            # the five products of the mathematical backward
            out[f"B{B}_p{p}_{'fused' if fused else 'unfused'}"] = {"ms": ms, "tflops_5prod": flop / ms / 1e9}
            print(B, p, fused, f"{ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s (5-product count)", flush=True)
        # forward for reference
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.attn_fwd(qkv, B, L, nh, D, lse=lse, drop=c); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[f"B{B}_p{p}_forward"] = {"ms": ms, "tflops": 4.0 * B * nh * L * L * D / ms / 1e9}
        print(B, p, "fwd", f"{ms:.3f} ms", flush=True)
json.dump(out, open("gpurun_out/attn_bwd_bench.json", "w"), indent=1)

# ---- forward, note-encoder shape (256 chunks x 12 heads x 512 tokens x 64): algo 0 / 5 = 64-key blocks, two CTAs per SM;
# 3 = 128-key blocks, one CTA per SM; 4 = 3 + ping-pong of the softmax warpgroups
B, L, nh, D = 256, 512, 12, 64
qkv = (torch.randn(B * L, 3 * nh * D, device="cuda") * 0.7).bfloat16()
for algo in (0, 3, 4, 5):
    for _ in range(3):
        ops.attn_fwd(qkv, B, L, nh, D, algo=algo)
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attn_fwd(qkv, B, L, nh, D, algo=algo); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    out[f"note_fwd_algo{algo}"] = {"ms": ms, "tflops": 4.0 * B * nh * L * L * D / ms / 1e9}
    print("note fwd algo", algo, f"{ms:.3f} ms  {4.0 * B * nh * L * L * D / ms / 1e9:.0f} TFLOP/s", flush=True)
json.dump(out, open("gpurun_out/attn_bwd_bench.json", "w"), indent=1)

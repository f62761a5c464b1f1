"""(Needs the instrumented build: FAME_NVCC_EXTRA=-DFAME_ATTN_INSTRUMENT python __graft_entry__.py --force.)
Where does the fused attention backward spend its time?  FAME_ATTN_DEBUG switches (results are wrong, timing only):
0 baseline, 1 no exponential / FMA math, 3 no math and no TMEM score loads, 4 no accumulating MMAs, 7 all three,
8 half of the streamed bytes, 15 everything."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fairmultimodal_b200 import ops, train
    B, L, nh, D = 512, 542, 8, 96
    qkv = (torch.randn(B * L, 3 * nh * D, device="cuda") * 0.7).bfloat16()
    dctx = (torch.randn(B * L, nh * D, device="cuda") * 0.1).bfloat16()
    lse = torch.empty(B, nh, L, device="cuda")
    ctx = ops.attn_fwd(qkv, B, L, nh, D, lse=lse)
    for _ in range(2):
        train._attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, fused=True)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); train._attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, fused=True); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"debug={os.environ.get('FAME_ATTN_DEBUG', '0')}: {sorted(ts)[2]:.3f} ms (B={B}, both passes + delta kernel)")
else:
    for d in ("0", "7", "8", "15"):
        env = dict(os.environ, FAME_ATTN_DEBUG=d)
        subprocess.run([sys.executable, __file__, "run"], env=env)

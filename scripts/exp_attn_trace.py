"""(Needs the instrumented build: FAME_NVCC_EXTRA=-DFAME_ATTN_INSTRUMENT python __graft_entry__.py --force.)
Event timeline (clock64) of CTA 0 of the fused attention backward: MMA issuer and softmax warp 0.  FAME_ATTN_DEBUG=16 (+7)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import _lib, ops, train
B, L, nh, D = 128, 542, 8, 96
qkv = (torch.randn(B * L, 3 * nh * D, device="cuda") * 0.7).bfloat16()
dctx = (torch.randn(B * L, nh * D, device="cuda") * 0.1).bfloat16()
lse = torch.empty(B, nh, L, device="cuda")
ctx = ops.attn_fwd(qkv, B, L, nh, D, lse=lse)
for _ in range(2):
    train._attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, fused=True)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 8192)()
lib.fame_debug_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int32]
rc = lib.fame_debug_attn_trace(buf, 8192)
a = np.frombuffer(buf, dtype=np.int64).reshape(2, 2048, 2)
for who, name in ((0, "mma"), (1, "softmax warp 0")):
    ev = a[who]
    n = int((ev[:, 1] != 0).sum())
    ev = ev[:n]
    t0 = ev[0, 1] if n else 0
    print(f"== {name}: {n} events (last launch = dQ pass); first 140:")
    prev = t0
    for e, t in ev[:140]:
        print(f"  ev {int(e):4d}  t {int(t - t0):8d}  +{int(t - prev):6d}")
        prev = t

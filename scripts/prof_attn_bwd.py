"""One fused attention backward (both passes) at 256 patients x 8 heads x 542 tokens x 96 for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import ops, train
B, L, nh, D = int(os.environ.get("PB", "256")), 542, 8, 96
qkv = (torch.randn(B * L, 3 * nh * D, device="cuda") * 0.7).bfloat16()
dctx = (torch.randn(B * L, nh * D, device="cuda") * 0.1).bfloat16()
lse = torch.empty(B, nh, L, device="cuda")
ctx = ops.attn_fwd(qkv, B, L, nh, D, lse=lse)
for _ in range(2):
    train._attn_backward(qkv, dctx, ctx, lse, B, L, nh, D, fused=True)
torch.cuda.synchronize()

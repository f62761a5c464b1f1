"""The lab tower's memory-bound kernels alone at the training-step size (32 patients x 542 tokens = 17 344 rows x 768):
time per launch with inputs L2-warm (as inside the step, where the producer has just written them) and L2-cold, against
the bytes each one has to move.  Usage: python scripts/bench_lab_small_kernels.py [out.json]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import _lib, ops
from fairmultimodal_b200 import ops_train as T

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
B, L, H = 32, 542, 768
rows = B * L


def timeit(fn, cold, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {}


def report(name, fn, nbytes):
    w, c = timeit(fn, False), timeit(fn, True)
    out[name] = dict(us_warm=1e3 * w, us_cold=1e3 * c, bytes=nbytes, frac_cold=nbytes / c / 1e6 / peak)
    print(f"{name:28s} warm {1e3 * w:6.1f} us   cold {1e3 * c:6.1f} us   {nbytes / 1e6:6.1f} MB   cold = {100 * nbytes / c / 1e6 / peak:4.1f} % of copy bw", flush=True)


step = torch.zeros(1, device="cuda", dtype=torch.int32)
c = _lib.DropoutCfg(); c.step = step.data_ptr(); c.seed = 99; c.thresh16 = 6554; c.group_shift = 0
x = torch.randn(rows, H, device="cuda").bfloat16()
r = torch.randn(rows, H, device="cuda").bfloat16()
dy = torch.randn(rows, H, device="cuda").bfloat16()
g, b = torch.randn(H, device="cuda"), torch.randn(H, device="cuda")
stats = torch.empty(rows, 2, device="cuda")
nb = rows * H * 2
report("layernorm fwd", lambda: ops.layernorm(x, g, b, 1e-5, stats=stats), 2 * nb)
report("layernorm fwd + residual", lambda: ops.layernorm(x, g, b, 1e-5, stats=stats, residual=r), 3 * nb)
dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
report("ln_bwd", lambda: T.layernorm_bwd_drop(x, dy, stats, g, dg, db), 3 * nb)
report("ln_bwd+drop", lambda: T.layernorm_bwd_drop(x, dy, stats, g, dg, db, drop=c), 4 * nb)
report("ln_bwd+drop+residual", lambda: T.layernorm_bwd_drop(x, dy, stats, g, dg, db, drop=c, residual=r), 5 * nb)
for cols in (768, 2048, 2304):
    y = torch.randn(rows, cols, device="cuda").bfloat16()
    o = torch.zeros(cols, device="cuda")
    report(f"colsum {cols}", lambda: T.colsum(y, o), rows * cols * 2)
report("attn_delta", lambda: T.attn_delta(dy, x, B, L, 8, 96), 2 * nb + rows * 8 * 4)
lab = torch.randn(B, L, device="cuda")
dpos, dw, dbias = torch.zeros(L, H, device="cuda"), torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
report("lab_embed_bwd", lambda: T.lab_embed_bwd(dy, lab, dpos, dw, dbias), nb + L * H * 4)
wt, bt, pos = torch.randn(H, device="cuda"), torch.randn(H, device="cuda"), torch.randn(L, H, device="cuda")
report("lab_embed", lambda: ops.lab_embed(lab, wt, bt, pos), nb + L * H * 4)
report("seq_mean", lambda: ops.seq_mean(x, B, L), nb)
dm = torch.randn(B, H, device="cuda")
report("seq_mean_bwd", lambda: T.seq_mean_bwd(dm, B, L), nb)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

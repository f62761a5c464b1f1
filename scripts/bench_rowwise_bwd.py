"""LayerNorm backward and bias column sums alone at the config-3 size (555 008 rows x 768 / 2048 / 2304 bf16): achieved
HBM bandwidth against the measured copy bandwidth (CUDA events, L2 flushed between iterations)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import _lib, ops
from fairmultimodal_b200 import ops_train as T

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
rows = 1024 * 542


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {}
step = torch.zeros(1, device="cuda", dtype=torch.int32)
c = _lib.DropoutCfg(); c.step = step.data_ptr(); c.seed = 99; c.thresh16 = 6554; c.group_shift = 0
x = torch.randn(rows, 768, device="cuda").bfloat16()
r = torch.randn(rows, 768, device="cuda").bfloat16()
dy = torch.randn(rows, 768, device="cuda").bfloat16()
g, b = torch.randn(768, device="cuda"), torch.randn(768, device="cuda")
stats = torch.empty(rows, 2, device="cuda")
ops.layernorm(x, g, b, 1e-5, stats=stats, residual=r)
dg, db = torch.zeros(768, device="cuda"), torch.zeros(768, device="cuda")
nb = rows * 768 * 2
for name, kw, traffic in (("ln_bwd", dict(), 3 * nb), ("ln_bwd+drop", dict(drop=c), 4 * nb), ("ln_bwd+drop+residual", dict(drop=c, residual=r), 5 * nb)):
    ms = timeit(lambda: T.layernorm_bwd_drop(x, dy, stats, g, dg, db, **kw))
    out[name] = dict(ms=ms, gbs=traffic / ms / 1e6, frac=traffic / ms / 1e6 / peak)
    print(f"{name:24s} {ms:7.3f} ms  {traffic / ms / 1e6:7.0f} GB/s  {100 * traffic / ms / 1e6 / peak:5.1f} % of the copy bandwidth", flush=True)
del x, r, dy
for cols in (768, 2048, 2304):
    y = torch.randn(rows, cols, device="cuda").bfloat16()
    o = torch.zeros(cols, device="cuda")
    ms = timeit(lambda: T.colsum(y, o))
    traffic = rows * cols * 2
    out[f"colsum_{cols}"] = dict(ms=ms, gbs=traffic / ms / 1e6, frac=traffic / ms / 1e6 / peak)
    print(f"colsum {cols:5d} cols        {ms:7.3f} ms  {traffic / ms / 1e6:7.0f} GB/s  {100 * traffic / ms / 1e6 / peak:5.1f} % of the copy bandwidth", flush=True)
    del y
json.dump(out, open("gpurun_out/rowwise_bwd_bench.json", "w"), indent=1)

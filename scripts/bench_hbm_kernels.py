"""Achieved HBM bandwidth of the memory-bound kernels of the FAME hot path (north_star item 2/3: pool, fusion,
EDDI / loss statistics, evaluation counts, AdamW), at the BASELINE config size (latency) and at a scaled size whose
algorithmic traffic is >= 1 GB (bandwidth).  Algorithmic bytes are the formulas of SURVEY.md 8(d) / DESIGN.md.

    python scripts/bench_hbm_kernels.py [out.json]

Timing: CUDA events on the launching stream, 5 warm-up + 20 timed launches, inputs rotated through > 126 MB (L2) of
distinct buffers when the working set is smaller than L2; peak = MEASURED_PEAKS.json hbm_gbs (copy bandwidth).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fairmultimodal_b200 import ops, synth, train  # noqa: E402
from fairmultimodal_b200 import ops_train as T  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = 6553.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
L2_BYTES = 126e6
flush_buf = torch.empty(int(256e6), device=dev, dtype=torch.uint8)


def timeit(fn, n_variants, warm=5, iters=20, flush=False):
    """fn(i) launches the kernel on input variant i % n_variants.  With flush=True L2 is overwritten between
    launches (outside the timed spans: one event pair per launch)."""
    for i in range(warm):
        fn(i % n_variants)
    torch.cuda.synchronize()
    if not flush:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i % n_variants)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    tot = 0.0
    for i in range(iters):
        flush_buf.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i % n_variants)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


rows = []


def report(kernel, case, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6
    rows.append(dict(kernel=kernel, case=case, algorithmic_bytes=int(nbytes), ms=round(ms, 5), gbs=round(gbs, 1),
                     frac_of_measured_hbm_peak=round(gbs / PEAK, 4), note=note))
    print(f"{kernel:22s} {case:44s} {nbytes / 1e6:10.2f} MB  {ms * 1e3:9.1f} us  {gbs:8.1f} GB/s  "
          f"{gbs / PEAK * 100:5.1f}% of measured HBM peak {note}", flush=True)


rng = np.random.default_rng(0)

# ---------------------------------------------------------------- K5 chunk -> patient mean pool (10_FAME.py:153-172)
for P, lo, hi, case in ((1000, 4, 4, "config 1: 1k patients x 4 chunks"),
                        (46000, 1, 16, "config 5: 46k patients x U{1..16} chunks"),
                        (400000, 1, 16, "scaled: 400k patients x U{1..16} chunks")):
    n = rng.integers(lo, hi + 1, size=P)
    off = np.zeros(P + 1, dtype=np.int32)
    off[1:] = np.cumsum(n)
    C = int(off[-1])
    offs = torch.from_numpy(off).to(dev)
    nbytes = C * 768 * 2 + (P + 1) * 4 + P * 768 * 4
    nv = max(1, int(2 * L2_BYTES // nbytes) + 1) if nbytes < 4 * L2_BYTES else 1
    xs = [torch.randn(C, 768, device=dev, dtype=torch.bfloat16) for _ in range(min(nv, 8))]
    ms = timeit(lambda i: ops.segment_mean(xs[i % len(xs)], offs), len(xs), flush=nbytes < 4 * L2_BYTES)
    report("segment_mean", case, nbytes, ms, "bf16 CLS rows in, f32 out")
    del xs

# ---------------------------------------------------------------- lab sequence mean (10_FAME.py:223)
for B, L, case in ((32, 542, "config 4: 32 x 542 tokens"), (1024, 542, "config 3: 1024 x 542 tokens")):
    x = torch.randn(B * L, 768, device=dev, dtype=torch.bfloat16)
    nbytes = B * L * 768 * 2 + B * 768 * 4
    ms = timeit(lambda i: ops.seq_mean(x, B, L), 1, flush=nbytes < 4 * L2_BYTES)
    report("seq_mean", case, nbytes, ms)
    del x

# ---------------------------------------------------------------- K7 fusion head forward (10_FAME.py:261-313)
shapes = {k: v for k, v in synth.fame_shapes(lab_tokens=4).items() if not k.startswith("behrt_")}
sd = {k: torch.from_numpy(v).to(dev) for k, v in synth.synth_state_dict(shapes, 2).items()}
proj = ("demo_projector.0.", "lab_projector.0.", "text_projector.0.")
cls = ("classifier_demo.", "classifier_lab.", "classifier_text.")
pk = dict(wp_t=torch.stack([sd[p + "weight"].t().contiguous() for p in proj]).contiguous(),
          bp=torch.stack([sd[p + "bias"] for p in proj]).contiguous(), sig_w=sd["sig_weights"],
          w3_t=sd["fusion_mlp.0.weight"].t().contiguous(), b3=sd["fusion_mlp.0.bias"], w4=sd["fusion_mlp.3.weight"],
          b4=sd["fusion_mlp.3.bias"], wc=torch.stack([sd[c + "weight"] for c in cls]).contiguous(),
          bc=torch.stack([sd[c + "bias"] for c in cls]).contiguous())
for B, case in ((32, "config 4: 32 patients"), (46000, "config 5: 46k patients"), (400000, "scaled: 400k patients")):
    emb = [torch.randn(B, 768, device=dev) for _ in range(3)]
    nbytes = B * (3 * 768 * 4 + 3 * 4) + 4 * (3 * 768 * 256 + 768 * 512)
    ms = timeit(lambda i: ops.fusion_fwd(emb, pk, (0.33, 0.33, 0.33)), 1, flush=nbytes < 4 * L2_BYTES)
    report("fusion_fwd", case, nbytes, ms, "3 x f32 [B,768] in, logits out; 2.2 MFLOP/patient fp32 on CUDA cores")
    del emb

# ---------------------------------------------------------------- K8 loss statistics, K9 evaluation counts
for N, case in ((32, "config 4: 32 patients"), (46000, "config 5: 46k patients"), (32_000_000, "scaled: 32M patients")):
    z = torch.randn(N, 3, device=dev)
    y = (torch.rand(N, 3, device=dev) < 0.3).float()
    attrs = [torch.randint(0, 5, (N,), device=dev, dtype=torch.int64) for _ in range(3)]
    pw = torch.tensor([4.9, 1.3, 0.55], device=dev)
    nbytes = N * 48
    st = torch.zeros(104, device=dev, dtype=torch.int64)
    ms = timeit(lambda i: ops.loss_stats(z, y, attrs, pw, stats=st), 1, flush=nbytes < 4 * L2_BYTES)
    report("loss_stats", case, nbytes, ms, "48 B/patient (logits, labels f32 x3, 3 int64 codes)")
    stats = ops.loss_stats(z, y, attrs, pw)
    ms = timeit(lambda i: ops.loss_fwd_bwd(z, y, attrs, pw, stats, None, 0.8, 0.0), 1, flush=nbytes < 4 * L2_BYTES)
    report("loss_fwd_bwd", case, N * 60, ms, "48 B/patient in + 12 B/patient dlogits out")
    out = torch.zeros(914, device=dev, dtype=torch.int64)
    sweep = torch.linspace(0, 1, 101, dtype=torch.float64, device=dev)
    ms = timeit(lambda i: ops.eval_counts(z, y, attrs, (0.5, 0.4, 0.6), out=out), 1, flush=nbytes < 4 * L2_BYTES)
    report("eval_counts", case, nbytes, ms, "EDDI / EO confusion counts, 3 outcomes x 3 attrs")
    ms = timeit(lambda i: ops.eval_counts(z, y, attrs, (0.5, 0.4, 0.6), sweep=sweep, out=out), 1,
                flush=nbytes < 4 * L2_BYTES)
    report("eval_counts+f1sweep", case, nbytes, ms, "+ 101-threshold F1 histogram (calibrate_thresholds)")
    del z, y, attrs

# ---------------------------------------------------------------- clip + AdamW over the flat buffers (10_FAME.py:446-447)
n = 97_910_792
p = torch.randn(n, device=dev) * 0.02
g = torch.randn(n, device=dev) * 1e-3
m = torch.zeros(n, device=dev)
v = torch.zeros(n, device=dev)
pb = torch.empty(n, device=dev, dtype=torch.bfloat16)
sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
gn = torch.zeros(1, device=dev)
step_dev = torch.ones(1, device=dev, dtype=torch.int32)
hyper = torch.tensor([1e-5, 0.01], device=dev)
ms = timeit(lambda i: T.grad_sumsq(g, sumsq), 1)
report("grad_sumsq", "97.9M parameters", n * 4, ms)
ms = timeit(lambda i: T.clip_adamw(p, g, m, v, sumsq, 1.0, 1e-5, 0.9, 0.999, 1e-8, 0.01, 0, gn, step_dev=step_dev,
                                   hyper_dev=hyper, p_bf16=pb), 1)
report("clip_adamw", "97.9M parameters", n * 30, ms, "r: p,g,m,v  w: p,m,v + bf16 shadow = 30 B/param")

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "hbm_kernels.json")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(dict(peak_gbs=PEAK, peak_source="MEASURED_PEAKS.json hbm_gbs (copy)", rows=rows), open(out_path, "w"), indent=1)

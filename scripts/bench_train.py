"""Per-kernel breakdown of one FAME training step (BASELINE config 4a: B = 32 per GPU, L = 542, text embeddings
precomputed) and of the config-3 forward+backward shape.  Usage: python scripts/bench_train.py [B] [L] [steps]"""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, ops, synth, train  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 542
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10

demo = modules.BEHRTModel_Demo(5, 2, 5, 5)
lab = modules.BEHRTModel_Lab(L)
model = modules.MultimodalTransformer_EDDI_Sigmoid(768, demo, lab, "cuda").cuda()
co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=1)
co["text"] = np.random.default_rng(0).standard_normal((B, 768)).astype(np.float32)
keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features",
        "text", "labels")
batch = [torch.from_numpy(co[k]).cuda() for k in keys]
pw = torch.from_numpy(synth.pos_weight(co["labels"])).cuda()
if os.environ.get("FAME_DROPOUT", "1") == "0":
    modules.set_dropout(model, 0.0)
print("dropout:", "off (parity configuration)" if os.environ.get("FAME_DROPOUT", "1") == "0" else "0.1 (reference train() mode)")
model.train()
st = train.get_state(model)
w = (0.33, 0.33, 0.33)


hp = dict(lr=1e-5, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
graph = os.environ.get("FAME_NO_GRAPH") is None


def one(g=None):
    train.optimisation_step(model, batch, pw, 0.8, 0.01, w, hp, use_graph=graph if g is None else g)


for _ in range(3):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    one()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"train step B={B} L={L}: {ms:.3f} ms/step  {B / ms * 1e3:.1f} patients/s  launches/step={0}")
l0 = ops.LAUNCHES
ops.start_trace()
one(False)
torch.cuda.synchronize()
tr = ops.stop_trace()
agg = collections.OrderedDict()
for name, tag, a, b, work in tr:
    k = name + (":" + tag.split("x")[0] if tag in ("dgrad", "wgrad") or tag.startswith("attn") else "")
    d = agg.setdefault(k, [0.0, 0, 0.0])
    d[0] += a.elapsed_time(b)
    d[1] += 1
    d[2] += work
tot = sum(v[0] for v in agg.values())
print(f"traced step: sum of kernel spans {tot:.3f} ms over {ops.LAUNCHES - l0} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    tf = f"{v[2] / v[0] / 1e9:8.1f} TF/s" if "gemm" in k and v[0] > 0 else ""
    print(f"  {k:38s} {v[0]:8.3f} ms  x{v[1]:4d}  {tf}")

"""Diagnostic: is the difference between the CUDA-graph trajectory and the eager trajectory of train_step larger than
the run-to-run difference of two eager trajectories (float-atomic ordering amplified by Adam)?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, synth, train  # noqa: E402

KEYS9 = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids",
         "lab_features", "text", "labels")
L, B = 24, 8
co = synth.make_cohort(B * 5, lab_tokens=L, chunks=0, with_tokens=False, seed=31)
co["text"] = (np.random.default_rng(2).standard_normal((B * 5, 768)) * 0.5).astype(np.float32)
batches = [[torch.from_numpy(co[k][i * B:(i + 1) * B]) for k in KEYS9] for i in range(5)]
pw = torch.tensor([3.0, 1.2, 0.6])
w0 = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.fame_shapes(lab_tokens=L), 12).items()}


def run(graph):
    train.USE_CUDA_GRAPH = graph
    demo = modules.BEHRTModel_Demo(5, 2, 5, 5)
    lab = modules.BEHRTModel_Lab(L)
    model = modules.MultimodalTransformer_EDDI_Sigmoid(768, demo, lab, "cuda")
    model.load_state_dict(w0)
    modules.set_dropout(model, 0.0)
    model = model.cuda()
    crit = torch.nn.BCEWithLogitsLoss(pos_weight=pw.cuda())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    st = train.get_state(model)
    p0 = st.p.clone()
    losses = train.train_step(model, batches, opt, "cuda", crit, lambda_edd=0.8, lambda_l1=0.01)
    return losses, (st.p - p0).clone(), st


la, da, st = run(False)
lb, db, _ = run(False)
lg, dg, _ = run(True)
print("losses eager/eager/graph:", la, lb, lg)
n = da.norm()
print(f"eager-eager rel diff {((da - db).norm() / n).item():.4f}   graph-eager rel diff {((dg - da).norm() / n).item():.4f}"
      f"   graph-eager2 {((dg - db).norm() / n).item():.4f}")
rows = []
for name, off in st.offsets.items():
    k = st.views[name].numel()
    a, b, g = da[off:off + k], db[off:off + k], dg[off:off + k]
    rows.append((((g - a).norm() ** 2).item(), ((a - b).norm() ** 2).item(), (a.norm() ** 2).item(), name))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("top contributors to |graph - eager|^2  (share, same for eager-eager, |update|^2):")
for r in rows[:14]:
    print(f"  {r[0] / tot:6.3f}  ee={r[1]:.3e}  ge={r[0]:.3e}  upd={r[2]:.3e}  {r[3]}")

"""Diagnostic: per-step determinism of the training step (two eager runs, lr = 0)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, synth, train
KEYS9 = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features", "text", "labels")
L, B = 24, 8
co = synth.make_cohort(B * 5, lab_tokens=L, chunks=0, with_tokens=False, seed=31)
co["text"] = (np.random.default_rng(2).standard_normal((B * 5, 768)) * 0.5).astype(np.float32)
batches = [[torch.from_numpy(co[k][i * B:(i + 1) * B]).cuda() for k in KEYS9] for i in range(5)]
pw = torch.tensor([3.0, 1.2, 0.6]).cuda()
w0 = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.fame_shapes(lab_tokens=L), 12).items()}
hp = dict(lr=0.0, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8)
def run():
    m = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), "cuda")
    m.load_state_dict(w0); modules.set_dropout(m, 0.0); m = m.cuda().train()
    st = train.get_state(m)
    out = []
    for b in batches:
        train.optimisation_step(m, b, pw, 0.8, 0.01, (0.33, 0.33, 0.33), hp, use_graph=False)
        torch.cuda.synchronize()
        out.append((st.g.clone(), st.m.clone(), st.grad_norm.item(), st.sumsq.item()))
    return out, st
a, st = run(); b, _ = run()
for i in range(5):
    dg = ((a[i][0]-b[i][0]).norm()/a[i][0].norm()).item(); dm = ((a[i][1]-b[i][1]).norm()/a[i][1].norm()).item()
    print(f"step {i}: g rel diff {dg:.2e}  m rel diff {dm:.2e}  grad_norm {a[i][2]:.6f} vs {b[i][2]:.6f}  sumsq {a[i][3]:.9e} vs {b[i][3]:.9e}")
i = 0
rows = []
for name, off in st.offsets.items():
    n = st.views[name].numel()
    rows.append(((a[i][1][off:off+n]-b[i][1][off:off+n]).norm().item(), a[i][1][off:off+n].norm().item(), (a[i][0][off:off+n]-b[i][0][off:off+n]).norm().item(), name))
rows.sort(reverse=True)
for r in rows[:8]: print(f"step0 m diff={r[0]:.3e} m norm={r[1]:.3e} g diff={r[2]:.3e} {r[3]}")

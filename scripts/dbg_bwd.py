"""Diagnostic: repeatability of forward_backward on identical inputs, per parameter."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, synth, train
KEYS9 = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features", "text", "labels")
L, B = 24, 8
co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=31)
co["text"] = (np.random.default_rng(2).standard_normal((B, 768)) * 0.5).astype(np.float32)
batch = [torch.from_numpy(co[k]).cuda() for k in KEYS9]
pw = torch.tensor([3.0, 1.2, 0.6]).cuda()
w0 = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.fame_shapes(lab_tokens=L), 12).items()}
m = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), "cuda")
m.load_state_dict(w0); modules.set_dropout(m, 0.0); m = m.cuda().train()
st = train.get_state(m)
gs = []
for rep in range(6):
    junk = torch.randn((rep + 1) * 100003, device="cuda")
    train.forward_backward(m, batch, pw, 0.8, 0.01, (0.33, 0.33, 0.33))
    torch.cuda.synchronize()
    gs.append(st.g.clone())
    del junk
for rep in range(1, 6):
    print("rep", rep, "rel diff vs rep0:", ((gs[rep] - gs[0]).norm() / gs[0].norm()).item())
worst = max(range(1, 6), key=lambda r: (gs[r] - gs[0]).norm().item())
names = list(st.offsets.items())
print("per-parameter diffs (rep", worst, "vs 0), demo layers from 11 down:")
for name, off in reversed(names):
    n = st.views[name].numel()
    d = (gs[worst][off:off+n] - gs[0][off:off+n]).norm().item(); r = gs[0][off:off+n].norm().item()
    if ("layer.11." in name or "layer.10." in name or "embedding" in name or "lab" in name or "fusion" in name or "projector" in name) and "bias" not in name:
        print(f"  {d / (r + 1e-30):.3e}  {name}")

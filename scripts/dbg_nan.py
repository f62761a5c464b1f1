"""Diagnostic: poison every torch.empty with NaN and look for NaNs in the gradients (an uninitialised read)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_empty = torch.empty
def poisoned(*a, **k):
    t = _empty(*a, **k)
    if t.is_cuda:
        if t.dtype.is_floating_point: t.fill_(float("nan"))
        elif t.dtype == torch.uint8: t.fill_(0xAB)
    return t
torch.empty = poisoned
from fairmultimodal_b200 import modules, synth, train
KEYS9 = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features", "text", "labels")
for (L, B) in ((24, 8), (542, 32), (40, 5)):
    co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=31)
    co["text"] = (np.random.default_rng(2).standard_normal((B, 768)) * 0.5).astype(np.float32)
    batch = [torch.from_numpy(co[k]).cuda() for k in KEYS9]
    pw = torch.tensor([3.0, 1.2, 0.6]).cuda()
    w0 = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(synth.fame_shapes(lab_tokens=L), 12).items()}
    m = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), "cuda")
    m.load_state_dict(w0); modules.set_dropout(m, 0.0); m = m.cuda().train()
    st = train.get_state(m)
    gs = []
    for rep in range(4):
        loss, _ = train.forward_backward(m, batch, pw, 0.8, 0.01, (0.33, 0.33, 0.33))
        torch.cuda.synchronize()
        gs.append(st.g.clone())
    bad = [n for n, off in st.offsets.items() if not torch.isfinite(st.g[off:off + st.views[n].numel()]).all()]
    print(f"L={L} B={B}: loss {loss.tolist()} non-finite grads in {len(bad)} tensors: {bad[:6]}")
    print("   rep diffs:", [f"{((g - gs[0]).norm() / gs[0].norm()).item():.2e}" for g in gs[1:]])
    del m, st

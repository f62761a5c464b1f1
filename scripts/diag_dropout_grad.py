"""Diagnostic: cosine between the dropout-free gradient and gradients drawn with dropout 0.1 (several step counters), and of
their mean -- the mean over draws must approach the dropout-free direction if forward and backward masks agree."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fairmultimodal_b200 import modules, synth, train
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
KEYS9 = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features", "text", "labels")
L, B = 40, 16
co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=3)
co["text"] = (np.random.default_rng(1).standard_normal((B, 768)) * 0.5).astype(np.float32)
batch = [torch.from_numpy(co[k]).cuda() for k in KEYS9]
pw = torch.tensor([3.0, 1.2, 0.6], device="cuda")
w = (0.33, 0.33, 0.33)
shapes = synth.fame_shapes(lab_tokens=L)
model = modules.MultimodalTransformer_EDDI_Sigmoid(768, modules.BEHRTModel_Demo(5, 2, 5, 5), modules.BEHRTModel_Lab(L), "cuda")
model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 4).items()}, strict=True)
model = model.cuda().train()
modules.set_dropout(model, 0.0)
for lam in (0.8, 0.0):
    modules.set_dropout(model, 0.0)
    loss0, _ = train.forward_backward(model, batch, pw, lam, 0.01, w)
    st = train.get_state(model)
    g0 = st.g.clone()
    regs = {k: v for k, v in st.region.items()}
    modules.set_dropout(model, 0.1)
    st = train.get_state(model)
    acc = torch.zeros_like(g0)
    cs = []
    for k in range(24):
        st.step_dev.fill_(k)
        loss1, _ = train.forward_backward(model, batch, pw, lam, 0.01, w)
        g1 = st.g.clone()
        acc += g1
        cs.append(torch.nn.functional.cosine_similarity(g0, g1, dim=0).item())
    cm = torch.nn.functional.cosine_similarity(g0, acc, dim=0).item()
    print(f"lambda_edd={lam}: loss0 {loss0.tolist()} cos per draw min {min(cs):.3f} median {np.median(cs):.3f} max {max(cs):.3f}; cos(mean of 24) {cm:.3f}; |g0| {g0.norm():.3e} |mean| {(acc/24).norm():.3e}")
    idx = {}
    for n, o in st.offsets.items():
        t = "demo" if n.startswith("behrt_demo.") else "lab" if n.startswith("behrt_lab.") else "head"
        idx.setdefault(t, []).append((o, st.gr(n).numel()))
    for t, spans in idx.items():
        a = torch.cat([g0[o:o + k] for o, k in spans]); b = torch.cat([acc[o:o + k] for o, k in spans])
        print(f"   {t}: |g0| {a.norm():.3e} cos(mean) {torch.nn.functional.cosine_similarity(a, b, dim=0).item():.3f}")

"""Generate tests/golden/behrt_combined.npz by running the UNMODIFIED reference 01_BEHRT.py (structured-only baseline,
SURVEY.md 8 f-2) on seeded synthetic inputs.  Build container only (needs /root/reference):

    python oracle/make_golden_behrt.py

Weights are regenerated from fairmultimodal_b200.synth (behrt_combined_shapes, seed), not stored.  TEST INFRASTRUCTURE.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fairmultimodal_b200 import synth  # noqa: E402

REF = os.path.join(os.environ.get("FAME_REFERENCE_ROOT", "/root/reference"), "FinalCode", "New", "Final", "01_BEHRT.py")
OUT = os.path.join(ROOT, "tests", "golden", "behrt_combined.npz")
L, B, WSEED = 40, 12, 9


def load_ref():
    for name in ("iterstrat", "iterstrat.ml_stratifiers", "matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["iterstrat.ml_stratifiers"].MultilabelStratifiedShuffleSplit = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.lines"].Line2D = object
    spec = importlib.util.spec_from_file_location("behrt_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    torch.manual_seed(0)
    model = ref.BEHRTModel_Combined(L, hidden_size=768)
    shapes = synth.behrt_combined_shapes(lab_tokens=L)
    sd_ref = model.state_dict()
    assert list(sd_ref.keys()) == list(shapes.keys()), (list(sd_ref.keys())[:5], list(shapes.keys())[:5])
    assert all(tuple(sd_ref[k].shape) == tuple(shapes[k]) for k in shapes)
    w = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, WSEED).items()}
    model.load_state_dict(w, strict=True)
    for m in model.modules():                                   # parity configuration: dropout off
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=21)
    lab, labels = torch.from_numpy(co["lab_features"]), torch.from_numpy(co["labels"])
    out = {"lab": co["lab_features"], "labels": co["labels"]}
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = torch.cat(model(lab), dim=1).numpy()
    # one iteration of the training loop body, 01_BEHRT.py:217-232
    pw = np.array([3.0, 1.2, 0.6], dtype=np.float32)
    fns = [torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor(float(p))) for p in pw]
    opt = ref.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    model.train()
    opt.zero_grad()
    lm, ll, lc = model(lab)
    loss = fns[0](lm.squeeze(), labels[:, 0]) + fns[1](ll.squeeze(), labels[:, 1]) + fns[2](lc.squeeze(), labels[:, 2])
    loss.backward()
    out["pos_weight"], out["loss"] = pw, np.float32(loss.item())
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    out["grad_norm"] = np.float32(torch.sqrt(sum((g ** 2).sum() for g in grads.values())).item())
    keep = ("fusion_fc.weight", "fusion_fc.bias", "classifier_mort.weight", "classifier_los.bias",
            "lab_model.token_embedding.weight", "lab_model.transformer_encoder.layers.1.linear2.bias",
            "lab_model.transformer_encoder.layers.0.norm1.weight")
    for k in keep:
        g = grads[k].numpy()
        out["grad." + k] = g[:24] if g.ndim == 2 and g.shape[0] > 24 else g       # first rows only: small fixture
    out["gnorm_by_param_names"] = np.array(list(grads.keys()))
    out["gnorm_by_param"] = np.array([g.norm().item() for g in grads.values()], dtype=np.float32)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    for k in ("fusion_fc.bias", "classifier_mech.weight", "lab_model.transformer_encoder.layers.0.linear1.bias"):
        out["updated." + k] = model.state_dict()[k].numpy().copy()
    # metric variants on controlled predictions
    rng = np.random.default_rng(4)
    N = 800
    y = (rng.random(N) < 0.3).astype(int)
    score = np.round(np.clip(0.5 + (y - 0.3) * 0.4 + rng.standard_normal(N) * 0.3, 0, 1) * 64) / 64
    code = rng.choice(5, N, p=(.03, .10, .04, .13, .70))
    ed, sub = ref.compute_eddi(code, y, score, threshold=0.5)
    pred = (score > 0.5).astype(int)
    tpr, fpr = {}, {}
    for gval in np.unique(code):
        tpr[gval], fpr[gval] = ref.calculate_tpr_and_fpr(y[code == gval], pred[code == gval])
    eo = ref.calculate_equalized_odds_difference(tpr, fpr)
    out.update(m_y=y.astype(np.float32), m_score=score.astype(np.float32), m_code=code.astype(np.int64),
               m_eddi=np.float64(ed), m_eddi_sub=np.array([sub[g] for g in sorted(sub)], dtype=np.float64),
               m_eo=np.array([eo["EOTPR"], eo["EOFPR"], eo["EO"]], dtype=np.float64),
               m_attr_eddi=np.float64(ref.compute_attribute_eddi(0.12, 0.05, 0.2)))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if not k.startswith("grad.")})


if __name__ == "__main__":
    main()

"""Generate tests/golden/eddi_fusion.npz by running the UNMODIFIED reference 08_multimodal_eddi_fusion.py (per-batch
EDDI-weighted logit fusion) on seeded synthetic inputs.  Build container only.  TEST INFRASTRUCTURE.
    python oracle/make_golden_eddifusion.py
The oracle restatement is pinned against this fixture now; the B200 implementation of this ablation is a next-round row.
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fairmultimodal_b200 import synth  # noqa: E402

REF = os.path.join(os.environ.get("FAME_REFERENCE_ROOT", "/root/reference"), "FinalCode", "New", "Final",
                   "08_multimodal_eddi_fusion.py")
OUT = os.path.join(ROOT, "tests", "golden", "eddi_fusion.npz")
L, B, WSEED = 24, 16, 23


def load_ref():
    for name in ("iterstrat", "iterstrat.ml_stratifiers", "matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["iterstrat.ml_stratifiers"].MultilabelStratifiedShuffleSplit = object
    spec = importlib.util.spec_from_file_location("eddi_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    torch.manual_seed(0)
    model = ref.MultimodalTransformer(768, ref.BEHRTModel_Demo(5, 2, 5, 5), ref.BEHRTModel_Lab(L), "cpu", beta=0.3)
    shapes = synth.eddi_fusion_shapes(lab_tokens=L)
    sd_ref = model.state_dict()
    assert list(sd_ref.keys()) == list(shapes.keys()), [a for a, b in zip(sd_ref.keys(), shapes.keys()) if a != b][:5]
    assert all(tuple(sd_ref[k].shape) == tuple(shapes[k]) for k in shapes)
    w = synth.synth_state_dict(shapes, WSEED)
    for k in w:                                                   # larger scalar heads: predictions on both sides of 0.5
        if k.startswith("classifier_") and k.endswith("weight"):
            w[k] = w[k] * 6.0
    model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=41)
    text = (np.random.default_rng(6).standard_normal((B, 768)) * 0.8).astype(np.float32)
    t = lambda k: torch.from_numpy(co[k])
    batch8 = (t("demo_dummy_ids"), t("demo_attn_mask"), t("age_ids"), t("gender_ids"), t("ethnicity_ids"),
              t("insurance_ids"), t("lab_features"), torch.from_numpy(text))
    labels = t("labels")
    out = {"text": text, "cohort_seed": np.int64(41), "labels": co["labels"], "head_scale": np.float32(6.0)}
    yd = {"mortality": co["labels"][:, 0], "los": co["labels"][:, 1], "mechanical_ventilation": co["labels"][:, 2]}
    sdict = {k: co["gender_ids"] for k in yd}
    old = {"mortality": (0.4, 0.3, 0.3), "los": (0.33, 0.33, 0.33), "mechanical_ventilation": (0.2, 0.5, 0.3)}
    model.eval()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        a, b, c, det = model(*batch8, beta=0.3, y_true_dict=yd, sensitive_labels_dict=sdict, old_eddi_weights=old)
        a0, b0, c0, _ = model(*batch8)                              # no labels: EDDI = 0, weights 0.33
    out["logits"] = torch.cat([a, b, c], dim=1).numpy()
    out["logits_plain"] = torch.cat([a0, b0, c0], dim=1).numpy()
    names = ("mortality", "los", "mechanical_ventilation")
    out["weights"] = np.array([det[n]["weights"] for n in names], dtype=np.float64)
    out["eddi"] = np.array([det[n]["eddi"][:3] for n in names], dtype=np.float64)
    out["old_weights"] = np.array([old[n] for n in names], dtype=np.float64)
    # one train_step (its criteria are module globals set by the script's main)
    pw = np.array([3.0, 1.2, 0.6], dtype=np.float32)
    ref.criterion_mortality, ref.criterion_los, ref.criterion_mech = (
        ref.FocalLoss(gamma=1, pos_weight=torch.tensor(float(p)), reduction="mean") for p in pw)
    model.train()
    with contextlib.redirect_stdout(io.StringIO()):
        a, b, c, _ = model(*batch8, beta=0.3, y_true_dict=yd, sensitive_labels_dict=sdict, old_eddi_weights=old)
    loss = (ref.criterion_mortality(a, labels[:, 0:1]) + ref.criterion_los(b, labels[:, 1:2]) +
            ref.criterion_mech(c, labels[:, 2:3]) + 1.0 * ((a - 1.0) ** 2).mean())
    model.zero_grad()
    loss.backward()
    out["pos_weight"], out["loss"] = pw, np.float64(loss.item())
    gn, nm = [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            nm.append(k)
            gn.append(p.grad.norm().item())
    out["gnorm_names"], out["gnorm"] = np.array(nm), np.array(gn, dtype=np.float32)
    for k in ("classifier_demo_mort.weight", "classifier_text_mv.bias", "lab_projector.0.bias"):
        out["grad." + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    ds = ref.TensorDataset(*batch8, labels[:, 0], labels[:, 1], labels[:, 2])
    opt = ref.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    with contextlib.redirect_stdout(io.StringIO()):
        out["epoch_loss"] = np.float64(ref.train_step(model, ref.DataLoader(ds, batch_size=B), opt, "cpu", beta=0.3,
                                                      loss_gamma=1.0, target=1.0, old_eddi_weights=old))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, float(out["loss"]), float(out["epoch_loss"]), out["eddi"].round(4).tolist(), out["weights"].round(4).tolist())


if __name__ == "__main__":
    main()

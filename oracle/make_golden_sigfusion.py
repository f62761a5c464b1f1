"""Generate tests/golden/sigmoid_fusion.npz by running the UNMODIFIED reference 09_multimodal_sigmoid_fusion.py
(per-modality sigmoid-gate ablation) on seeded synthetic inputs.  Build container only.
    python oracle/make_golden_sigfusion.py                                            TEST INFRASTRUCTURE.
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

REF = os.path.join(os.environ.get("FAME_REFERENCE_ROOT", "/root/reference"), "FinalCode", "New", "Final",
                   "09_multimodal_sigmoid_fusion.py")
OUT = os.path.join(ROOT, "tests", "golden", "sigmoid_fusion.npz")
L, B, WSEED = 24, 10, 17


def load_ref():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("sig_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    torch.manual_seed(0)
    model = ref.MultimodalTransformer(768, ref.BEHRTModel_Demo(5, 2, 5, 5), ref.BEHRTModel_Lab(L), "cpu")
    shapes = synth.sigmoid_fusion_shapes(lab_tokens=L)
    sd_ref = model.state_dict()
    assert list(sd_ref.keys()) == list(shapes.keys()), [a for a, b in zip(sd_ref.keys(), shapes.keys()) if a != b][:5]
    assert all(tuple(sd_ref[k].shape) == tuple(shapes[k]) for k in shapes)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, WSEED).items()}, strict=True)
    for m in model.modules():                                   # parity configuration: dropout off
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    co = synth.make_cohort(B, lab_tokens=L, chunks=0, with_tokens=False, seed=31)
    text = (np.random.default_rng(5).standard_normal((B, 768)) * 0.5).astype(np.float32)
    t = lambda k: torch.from_numpy(co[k])
    batch8 = (t("demo_dummy_ids"), t("demo_attn_mask"), t("age_ids"), t("gender_ids"), t("ethnicity_ids"),
              t("insurance_ids"), t("lab_features"), torch.from_numpy(text))
    labels = t("labels")
    out = {"text": text, "cohort_seed": np.int64(31), "labels": co["labels"]}
    model.eval()
    with torch.no_grad():
        lm, ll, lc, agg = model(*batch8)
    out["logits_eval"], out["agg_eval"] = torch.cat([lm, ll, lc], dim=1).numpy(), agg.numpy()
    pw = np.array([3.0, 1.2, 0.6], dtype=np.float32)
    crit = [ref.FocalLoss(gamma=1, pos_weight=torch.tensor(float(p)), reduction="mean") for p in pw]
    ds = ref.TensorDataset(*batch8, labels[:, 0], labels[:, 1], labels[:, 2])
    loader = ref.DataLoader(ds, batch_size=B, shuffle=False)
    opt = ref.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    # gradients of the batch before any update
    model.train()
    opt.zero_grad()
    lm, ll, lc, _ = model(*batch8)
    loss = crit[0](lm, labels[:, 0:1]) + crit[1](ll, labels[:, 1:2]) + crit[2](lc, labels[:, 2:3])
    loss.backward()
    out["pos_weight"], out["loss"] = pw, np.float64(loss.item())
    names, norms = [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            names.append(k)
            norms.append(p.grad.norm().item())
    out["gnorm_names"], out["gnorm"] = np.array(names), np.array(norms, dtype=np.float32)
    out["none_grad"] = np.array([k for k, p in model.named_parameters() if p.grad is None])
    for k in ("sig_weights_demo", "sig_weights_lab", "sig_weights_text", "classifier.3.weight", "classifier.0.bias",
              "aggregate_projector.0.bias", "text_projector.0.bias", "behrt_lab.transformer_encoder.layers.1.norm2.bias",
              "BEHRT.age_embedding.weight"):
        out["grad." + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    opt.zero_grad()
    out["epoch_loss"] = np.float64(ref.train_step(model, loader, opt, "cpu", crit[0], crit[1], crit[2]))
    for k in ("sig_weights_text", "classifier.3.bias", "aggregate_projector.0.bias"):
        out["after." + k] = model.state_dict()[k].numpy().copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, float(out["loss"]), float(out["epoch_loss"]), len(names), list(out["none_grad"]))


if __name__ == "__main__":
    main()

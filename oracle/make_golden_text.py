"""Generate tests/golden/text_classifier.npz by running the UNMODIFIED reference 02_BioClinicalBERT.py (text-only
baseline, SURVEY.md 8 f-1: UnstructuredClassifier + FocalLoss + train_model) on seeded synthetic inputs.
Build container only (needs /root/reference):   python oracle/make_golden_text.py        TEST INFRASTRUCTURE.
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
                   "02_BioClinicalBERT.py")
OUT = os.path.join(ROOT, "tests", "golden", "text_classifier.npz")
SHAPES = {"classifier.0.weight": (256, 768), "classifier.0.bias": (256,), "classifier.3.weight": (3, 256),
          "classifier.3.bias": (3,)}
B, WSEED = 24, 13


def load_ref():
    for name in ("iterstrat", "iterstrat.ml_stratifiers", "matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn",
                 "skmultilearn", "skmultilearn.model_selection"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["iterstrat.ml_stratifiers"].MultilabelStratifiedShuffleSplit = object
    sys.modules["skmultilearn.model_selection"].iterative_train_test_split = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.lines"].Line2D = object
    spec = importlib.util.spec_from_file_location("text_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    model = ref.UnstructuredClassifier(768, 256)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == SHAPES
    model.load_state_dict({k: torch.from_numpy(synth.synth_tensor(k, shp, WSEED) * (3.0 if "weight" in k else 1.0))
                           for k, shp in SHAPES.items()})
    model.classifier[2].p = 0.0                                   # parity configuration: dropout off
    rng = np.random.default_rng(8)
    emb = (rng.standard_normal((B, 768)) * 0.6).astype(np.float32)
    labels = (rng.random((B, 3)) < np.array([0.2, 0.4, 0.8])).astype(np.float32)
    pw = np.array([4.0, 1.5, 0.25], dtype=np.float32)
    out = {"emb": emb, "labels": labels, "pos_weight": pw}
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = model(torch.from_numpy(emb)).numpy()
    crit = [ref.FocalLoss(gamma=2, pos_weight=torch.tensor(float(p)), reduction="mean") for p in pw]
    # FocalLoss alone on controlled logits (including large magnitudes)
    z = torch.tensor(np.round(rng.standard_normal((40, 1)) * 3, 2).astype(np.float32))
    y = torch.tensor((rng.random((40, 1)) < 0.5).astype(np.float32))
    out["fl_z"], out["fl_y"], out["fl_value"] = z.numpy(), y.numpy(), np.float64(crit[0](z, y).item())
    # one epoch of train_model over 3 batches of 8 (02_BioClinicalBERT.py:137-152)
    ds = ref.UnstructuredDataset(emb, labels[:, 0], labels[:, 1], labels[:, 2])
    loader = ref.DataLoader(ds, batch_size=8, shuffle=False)
    opt = ref.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    # gradients of the first batch (before any update), for a direct check of the backward
    model.train()
    opt.zero_grad()
    e0, l0 = torch.from_numpy(emb[:8]), torch.from_numpy(labels[:8])
    lg = model(e0)
    loss0 = sum(crit[i](lg[:, i].unsqueeze(1), l0[:, i:i + 1]) for i in range(3))
    loss0.backward()
    out["loss_batch0"] = np.float64(loss0.item())
    for k, p in model.named_parameters():
        out["grad." + k] = p.grad.numpy()[:16].copy() if k == "classifier.0.weight" else p.grad.numpy().copy()
    opt.zero_grad()
    out["epoch_loss"] = np.float64(ref.train_model(model, loader, opt, "cpu", crit[0], crit[1], crit[2]))
    for k, v in model.state_dict().items():
        out["after." + k] = v.numpy()[:16].copy() if k == "classifier.0.weight" else v.numpy().copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, out["loss_batch0"], out["epoch_loss"], out["fl_value"])


if __name__ == "__main__":
    main()

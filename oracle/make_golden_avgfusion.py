"""Generate tests/golden/average_fusion.npz by running the UNMODIFIED reference 07_multimodal_average_fusion.py (concat
fusion over a seven-table demographic encoder) on seeded synthetic inputs.  Build container only.
    python oracle/make_golden_avgfusion.py                                              TEST INFRASTRUCTURE.
The oracle restatement is pinned against this fixture now; the B200 implementation of this ablation is a next-round row.
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
                   "07_multimodal_average_fusion.py")
OUT = os.path.join(ROOT, "tests", "golden", "average_fusion.npz")
B, WSEED = 14, 29
SIZES = dict(num_diseases=10, num_ages=5, num_segments=2, num_adm=4, num_disch=6, num_genders=2, num_eth=5, num_ins=5)


def load_ref():
    for name in ("iterstrat", "iterstrat.ml_stratifiers", "matplotlib", "matplotlib.pyplot", "matplotlib.lines", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["iterstrat.ml_stratifiers"].MultilabelStratifiedShuffleSplit = object
    spec = importlib.util.spec_from_file_location("avg_ref", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_ref()
    torch.manual_seed(0)
    behrt = ref.BEHRTModel(SIZES["num_diseases"], SIZES["num_ages"], SIZES["num_segments"], SIZES["num_adm"],
                           SIZES["num_disch"], SIZES["num_genders"], SIZES["num_eth"], SIZES["num_ins"])
    model = ref.MultimodalTransformer(768, behrt, "cpu")
    shapes = synth.average_fusion_shapes(**SIZES)
    sd_ref = model.state_dict()
    assert list(sd_ref.keys()) == list(shapes.keys()), [(a, b) for a, b in zip(sd_ref.keys(), shapes.keys()) if a != b][:5]
    assert all(tuple(sd_ref[k].shape) == tuple(shapes[k]) for k in shapes)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, WSEED).items()}, strict=True)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    rng = np.random.default_rng(12)
    codes = [rng.integers(0, n + 1, B).astype(np.int64) for n in (5, 2, 4, 6, 2, 5, 5)]     # n + 1: exercises the clamp
    text = (rng.standard_normal((B, 768)) * 0.6).astype(np.float32)
    labels = (rng.random((B, 3)) < np.array([0.2, 0.4, 0.8])).astype(np.float32)
    ids, mask = torch.zeros((B, 1), dtype=torch.long), torch.ones((B, 1), dtype=torch.long)
    args = (ids, mask, *[torch.from_numpy(c) for c in codes], torch.from_numpy(text))
    out = {"codes": np.stack(codes), "text": text, "labels": labels}
    model.eval()
    with torch.no_grad():
        a, b, c, pre = model(*args)
    out["logits_eval"], out["pre_relu_eval"] = torch.cat([a, b, c], dim=1).numpy(), pre.numpy()
    pw = np.array([3.0, 1.2, 0.6], dtype=np.float32)
    crit = [ref.FocalLoss(gamma=1, pos_weight=torch.tensor(float(p)), reduction="mean") for p in pw]
    model.train()
    model.zero_grad()
    a, b, c, _ = model(*args)
    lab = torch.from_numpy(labels)
    loss = crit[0](a, lab[:, 0:1]) + crit[1](b, lab[:, 1:2]) + crit[2](c, lab[:, 2:3])
    loss.backward()
    out["pos_weight"], out["loss"] = pw, np.float64(loss.item())
    nm, gn = [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            nm.append(k)
            gn.append(p.grad.norm().item())
    out["gnorm_names"], out["gnorm"] = np.array(nm), np.array(gn, dtype=np.float32)
    for k in ("ts_linear.bias", "text_linear.bias", "classifier.3.weight", "BEHRT.segment_embedding.weight",
              "BEHRT.discharge_loc_embedding.weight"):
        out["grad." + k] = dict(model.named_parameters())[k].grad.numpy().copy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, float(out["loss"]), len(nm))


if __name__ == "__main__":
    main()

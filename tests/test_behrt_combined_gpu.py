"""Structured-only baseline (01_BEHRT.py: BEHRTModel_Combined, summed-BCE training step, EO / EDDI variants) on the
B200 kernels against golden vectors of the unmodified reference and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
WSEED = 9


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _model(L):
    from fairmultimodal_b200 import behrt_combined as BC, modules, synth
    m = BC.BEHRTModel_Combined(L)
    shapes = synth.behrt_combined_shapes(lab_tokens=L)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())                      # same keys, same order as the reference
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    w = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, WSEED).items()}
    m.load_state_dict(w, strict=True)
    modules.set_dropout(m, 0.0)
    return m.cuda(), w


def test_forward_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "behrt_combined.npz"))
    model, _ = _model(g["lab"].shape[1])
    model.eval()
    lm, ll, lc = model(torch.from_numpy(g["lab"]).cuda())
    assert lm.shape == (g["lab"].shape[0], 1)
    got = torch.cat([lm, ll, lc], dim=1).cpu().numpy()
    ref = g["logits_eval"]
    assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max()         # bf16 encoder: logits rel 1e-2


def test_training_step_matches_reference_golden(golden_dir):
    """One iteration of the reference loop (01_BEHRT.py:217-232): loss, gradients, clipped AdamW update."""
    from fairmultimodal_b200 import behrt_combined as BC
    g = np.load(os.path.join(golden_dir, "behrt_combined.npz"))
    model, _ = _model(g["lab"].shape[1])
    model.train()
    lab, labels = torch.from_numpy(g["lab"]).cuda(), torch.from_numpy(g["labels"]).cuda()
    pw = torch.from_numpy(g["pos_weight"]).cuda()
    loss, logits = BC.forward_backward(model, lab, labels, pw)
    assert abs(loss.item() - float(g["loss"])) < 5e-3                  # summed BCE on bf16-encoder logits
    st = BC.get_state(model)
    names = [str(n) for n in g["gnorm_by_param_names"]]
    ref_norms = g["gnorm_by_param"]
    got = np.array([st.gr(n).norm().item() for n in names])
    rel = np.abs(got - ref_norms) / (ref_norms + 1e-12)
    assert np.median(rel) < 0.02 and rel.max() < 0.15, sorted(zip(rel, names))[-5:]
    gn = float(np.sqrt((got ** 2).sum()))
    assert abs(gn - float(g["grad_norm"])) < 0.03 * float(g["grad_norm"])
    for k in g.files:
        if k.startswith("grad."):
            ref = g[k]
            mine = st.gr(k[5:]).cpu().numpy()
            mine = mine[:ref.shape[0]] if ref.ndim == 2 else mine
            assert np.linalg.norm(mine - ref) <= 0.06 * np.linalg.norm(ref) + 1e-7, k
    st.clip_and_step(1e-3, 0.01, max_norm=1.0)
    for k in g.files:
        if k.startswith("updated."):
            ref = g[k]
            mine = model.state_dict()[k[8:]].cpu().numpy()
            # AdamW's first step moves every element by ~lr * sign(g) (lr = 1e-3): an element whose gradient is below
            # the bf16 noise of the encoder may move the other way, all others must land on the reference's value
            d = np.abs(mine - ref)
            assert (d < 3e-4).mean() > 0.95 and d.max() < 2.5e-3, (k, (d < 3e-4).mean(), d.max())


def test_train_epoch_and_dropout_and_metrics(golden_dir):
    """train_epoch over a DataLoader of the reference's 6-tuples with dropout 0.1 (finite, loss of the right size,
    parameters move); EO (sum / n^2) and EDDI (np.unique groups) variants from the integer count kernel."""
    from fairmultimodal_b200 import behrt_combined as BC, modules, synth
    g = np.load(os.path.join(golden_dir, "behrt_combined.npz"))
    L = g["lab"].shape[1]
    model, _ = _model(L)
    modules.set_dropout(model, 0.1)
    co = synth.make_cohort(48, lab_tokens=L, chunks=0, with_tokens=False, seed=2)
    t = lambda k: torch.from_numpy(co[k])
    ds = torch.utils.data.TensorDataset(t("lab_features"), t("age_ids"), t("gender_ids"), t("ethnicity_ids"),
                                        t("insurance_ids"), t("labels"))
    loader = torch.utils.data.DataLoader(ds, batch_size=16, shuffle=False)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    before = model.fusion_fc.weight.detach().clone()
    l1 = BC.train_epoch(model, loader, opt, "cuda", [3.0, 1.2, 0.6])
    l2 = BC.train_epoch(model, loader, opt, "cuda", [3.0, 1.2, 0.6])
    assert np.isfinite(l1) and np.isfinite(l2) and 0.5 < l1 < 10.0
    assert not torch.equal(before, model.fusion_fc.weight.detach())
    model.eval()
    outs = model(t("lab_features").cuda())
    assert all(torch.isfinite(o).all() for o in outs)
    # metric variants vs the reference's numbers
    code, y, score = g["m_code"], g["m_y"], g["m_score"]
    ed, sub = BC.compute_eddi(code, y, score, 0.5)
    assert abs(ed - float(g["m_eddi"])) < 1e-12
    np.testing.assert_allclose([sub[k] for k in sorted(sub)], g["m_eddi_sub"], atol=1e-12)
    tpr, fpr = BC.group_tpr_fpr(code, y, (score > 0.5).astype(np.float32))
    eo = BC.calculate_equalized_odds_difference(tpr, fpr)
    np.testing.assert_allclose([eo["EOTPR"], eo["EOFPR"], eo["EO"]], g["m_eo"], atol=1e-12)
    assert abs(BC.compute_attribute_eddi(0.12, 0.05, 0.2) - float(g["m_attr_eddi"])) < 1e-12

"""Per-batch EDDI-weighted logit fusion (08_multimodal_eddi_fusion.py: six-layer demographic BERT, nine scalar heads,
in-forward compute_eddi over gender, FocalLoss gamma 1 + (mortality logit - target)^2, clip 1.0 + AdamW) on the B200
kernels against golden vectors of the unmodified reference."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

L, WSEED = 24, 23
NAMES = ("mortality", "los", "mechanical_ventilation")


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _model(g):
    from fairmultimodal_b200 import eddi_fusion as EF, modules, synth
    m = EF.MultimodalTransformer(768, EF.BEHRTModel_Demo(5, 2, 5, 5), EF.BEHRTModel_Lab(L), "cuda", beta=0.3)
    shapes = synth.eddi_fusion_shapes(lab_tokens=L)
    sd = m.state_dict()
    assert list(sd.keys()) == list(shapes.keys())                      # same keys, same order as the reference
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    w = synth.synth_state_dict(shapes, WSEED)
    for k in w:
        if k.startswith("classifier_") and k.endswith("weight"):
            w[k] = w[k] * float(g["head_scale"])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    modules.set_dropout(m, 0.0)
    return m.cuda()


def _inputs(g):
    from fairmultimodal_b200 import synth
    co = synth.make_cohort(g["labels"].shape[0], lab_tokens=L, chunks=0, with_tokens=False, seed=int(g["cohort_seed"]))
    t = lambda k: torch.from_numpy(co[k]).cuda()
    return [t("demo_dummy_ids"), t("demo_attn_mask"), t("age_ids"), t("gender_ids"), t("ethnicity_ids"), t("insurance_ids"),
            t("lab_features"), torch.from_numpy(g["text"]).cuda()], t("labels"), co


def _old(g):
    return {n: tuple(float(x) for x in g["old_weights"][i]) for i, n in enumerate(NAMES)}


def test_counts_formula_equals_numpy_compute_eddi():
    """eddi_from_counts on integer cells == the reference's numpy formula (restated in the oracle) on the raw arrays."""
    from fairmultimodal_b200 import eddi_fusion as EF
    from oracle import fame_oracle as O
    rng = np.random.default_rng(5)
    for n, groups in ((16, 2), (257, 5), (40, 1)):
        y = rng.integers(0, 2, n).astype(np.float32)
        p = rng.random(n).astype(np.float32)
        s = rng.integers(0, groups, n)
        if n == 40:
            p = np.where(y > 0, 0.9, 0.1).astype(np.float32)             # error rate 0: denominator 1.0
        pred = p > 0.5
        cells = np.zeros((8, 4), dtype=np.int64)
        for c in range(8):
            m = s == c
            cells[c] = [(pred & (y > 0) & m).sum(), (~pred & (y > 0) & m).sum(), (pred & (y == 0) & m).sum(),
                        (~pred & (y == 0) & m).sum()]
        got, _ = EF.eddi_from_counts(cells)
        assert got == O.eddi_08(y, p, s)


@pytest.mark.gpu
def test_forward_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "eddi_fusion.npz"))
    model = _model(g).eval()
    batch8, labels, co = _inputs(g)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        a0, b0, c0, det0 = model(*batch8)                                 # no labels: EDDI 0, weights 0.33
        yd = {n: co["labels"][:, i] for i, n in enumerate(NAMES)}          # numpy, as the reference passes them
        sd = {n: co["gender_ids"] for n in NAMES}
        a, b, c, det = model(*batch8, beta=0.3, y_true_dict=yd, sensitive_labels_dict=sd, old_eddi_weights=_old(g))
    plain = torch.cat([a0, b0, c0], dim=1).cpu().numpy()
    assert np.abs(plain - g["logits_plain"]).max() <= 1e-2 * np.abs(g["logits_plain"]).max()     # bf16 towers: rel 1e-2
    assert all(det0[n]["weights"] == (0.33, 0.33, 0.33) and det0[n]["eddi"] == (0.0, 0.0, 0.0, 0.0) for n in NAMES)
    # predictions sit on the same side of 0.5 as the reference's whenever the raw logit is not within bf16 noise of 0:
    # the golden's heads are scaled so that they are not; then counts, EDDI and weights are exact
    eddi = np.array([det[n]["eddi"][:3] for n in NAMES])
    wts = np.array([det[n]["weights"] for n in NAMES])
    np.testing.assert_allclose(eddi, g["eddi"], atol=1e-15)
    np.testing.assert_allclose(wts, g["weights"], atol=1e-15)
    got = torch.cat([a, b, c], dim=1).cpu().numpy()
    assert np.abs(got - g["logits"]).max() <= 1e-2 * np.abs(g["logits"]).max()
    text = out.getvalue()
    assert text.count("No y_true or sensitive_labels provided, setting EDDI values to 0.") == 3
    e, w = g["eddi"][0], g["weights"][0]
    assert f"Computed EDDI - Demo: {e[0]:.4f}, Lab: {e[1]:.4f}, Text: {e[2]:.4f}" in text
    assert f"Modality weights - Demo: {w[0]:.4f}, Lab: {w[1]:.4f}, Text: {w[2]:.4f}" in text
    for n in NAMES:
        assert set(det[n]) == {"eddi", "weights", "probs", "subgroups"} and len(det[n]["probs"]) == 3
        assert det[n]["probs"][0].shape == (labels.shape[0], 1) and len(det[n]["subgroups"]) == 3


@pytest.mark.gpu
def test_training_step_matches_reference_golden(golden_dir):
    from fairmultimodal_b200 import eddi_fusion as EF, unstructured as U
    g = np.load(os.path.join(golden_dir, "eddi_fusion.npz"))
    model = _model(g).train()
    batch8, labels, _ = _inputs(g)
    pw = torch.from_numpy(g["pos_weight"]).cuda()
    with contextlib.redirect_stdout(io.StringIO()):
        loss, logits, det = EF.forward_backward(model, batch8, labels, pw, beta=0.3, loss_gamma=1.0, target=1.0,
                                                old_eddi_weights=_old(g))
    np.testing.assert_allclose(np.array([det[n]["weights"] for n in NAMES]), g["weights"], atol=1e-15)
    assert abs(loss.item() - float(g["loss"])) < 2e-2 * max(1.0, abs(float(g["loss"])))      # bf16-tower logits, scaled heads
    st = EF.get_state(model)
    names = [str(n) for n in g["gnorm_names"]]
    got = np.array([st.gr(n).norm().item() if n in st.offsets else 0.0 for n in names])
    ref = g["gnorm"]
    live = ref > 1e-5 * ref.max()
    rel = np.abs(got[live] - ref[live]) / ref[live]
    assert np.median(rel) < 0.03 and rel.max() < 0.3, sorted(zip(rel, np.array(names)[live]))[-5:]
    assert np.all(got[~live] <= 1e-5 * ref.max())
    assert not any(n.startswith("behrt_demo.bert.pooler.") for n in st.offsets)
    errs = {}
    for k in g.files:
        if k.startswith("grad."):
            mine = st.gr(k[5:]).cpu().numpy()
            errs[k] = float(np.linalg.norm(mine - g[k]) / (np.linalg.norm(g[k]) + 1e-12))
    assert max(errs.values()) < 0.25 and np.median(list(errs.values())) < 0.08, errs
    # the drop-in epoch: one batch, clip 1.0, AdamW(lr 1e-3); then validate_step on the updated model runs
    model2 = _model(g)
    ds = torch.utils.data.TensorDataset(*[b.cpu() for b in batch8], labels[:, 0].cpu(), labels[:, 1].cpu(), labels[:, 2].cpu())
    loader = torch.utils.data.DataLoader(ds, batch_size=labels.shape[0], shuffle=False)
    opt = torch.optim.AdamW(model2.parameters(), lr=1e-3, weight_decay=0.01)
    crit = dict(zip(("criterion_mortality", "criterion_los", "criterion_mech"),
                    [U.FocalLoss(gamma=1, pos_weight=torch.tensor(float(p)), reduction="mean") for p in g["pos_weight"]]))
    with contextlib.redirect_stdout(io.StringIO()):
        ep = EF.train_step(model2, loader, opt, "cuda", beta=0.3, loss_gamma=1.0, target=1.0, old_eddi_weights=_old(g), **crit)
        vl, last = EF.validate_step(model2, loader, "cuda", beta=0.3, old_eddi_weights=_old(g), **crit)
    assert abs(ep - float(g["epoch_loss"])) < 2e-2 * max(1.0, abs(float(g["epoch_loss"])))
    assert np.isfinite(vl) and set(last) == set(NAMES)
    # validate_step on the un-stepped weights = the same objective in eval mode (dropout is 0 in the golden)
    with contextlib.redirect_stdout(io.StringIO()):
        vl0, _ = EF.validate_step(_model(g), loader, "cuda", beta=0.3, old_eddi_weights=_old(g), **crit)
    assert abs(vl0 - float(g["loss"])) < 2e-2 * max(1.0, abs(float(g["loss"])))

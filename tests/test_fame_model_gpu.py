"""FAME model forward, loss and metric kernels on the B200 against the CPU oracle and the golden vectors produced
by the unmodified reference (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ("mortality", "los", "mechanical_ventilation")


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _build_model(L, seed):
    from fairmultimodal_b200 import modules, synth
    demo = modules.BEHRTModel_Demo(5, 2, 5, 5, hidden_size=768)
    lab = modules.BEHRTModel_Lab(lab_token_count=L, hidden_size=768, nhead=8, num_layers=2)
    model = modules.MultimodalTransformer_EDDI_Sigmoid(768, demo, lab, "cuda", fusion_hidden=512, beta=1.0)
    shapes = synth.fame_shapes(lab_tokens=L)
    sd = model.state_dict()
    assert list(sd.keys()) == list(shapes.keys())                      # same keys, same order as the reference
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    w = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, seed).items()}
    model.load_state_dict(w, strict=True)
    modules.set_dropout(model, 0.0)          # parity configuration (the golden vectors were made with dropout off)
    return model.cuda(), w


def _batch(g, dev="cuda"):
    keys = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids",
            "lab_features", "text", "labels")
    return [torch.from_numpy(g[k]).to(dev) for k in keys]


def _rel(got, ref):
    ref = torch.as_tensor(ref).float()
    return ((got.float().cpu() - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


def test_model_forward_matches_reference_golden(golden_dir):
    """logits rel 1e-2 under bf16 (north_star); encoders within 3e-2 of the largest reference magnitude."""
    g = _golden(golden_dir, "model_step.npz")
    model, _ = _build_model(24, int(g["wseed"]))
    model.eval()
    b = _batch(g)
    w = {"mortality": dict(zip(("demo", "lab", "text"), g["weights"].tolist()))}
    with torch.no_grad():
        demo = model.behrt_demo(*b[:6])
        lab = model.behrt_lab(b[6])
        o = model(*b[:8], old_eddi_weights=w, return_modality_logits=True, return_gated_vector=True,
                  return_intermediate=True)
        o_def = model(*b[:8])
    assert _rel(demo, g["demo_embedding"]) < 3e-2
    assert _rel(lab, g["lab_embedding"]) < 3e-2
    assert _rel(o["fused_logits"], g["fused_logits"]) < 1e-2
    assert _rel(o_def["fused_logits"], g["fused_logits_default"]) < 1e-2
    assert _rel(o["gated_vector"], g["gated_vector"]) < 2e-2
    assert _rel(o["fusion_pre_relu"], g["fusion_pre_relu"]) < 2e-2
    np.testing.assert_allclose(o["sigmoid_weights"].cpu().numpy(), g["sigmoid_weights"], rtol=1e-6)
    for m in ("demo", "lab", "text"):
        assert _rel(o["modality_logits"][m], g[f"modality_logits_{m}"]) < 2e-2
    assert o["dynamic_weights"] == w["mortality"]
    assert set(o) == {"fused_logits", "dynamic_weights", "sigmoid_weights", "modality_logits", "gated_vector",
                      "fusion_pre_relu"}


def test_fusion_kernel_fp32_exactness():
    """Given identical fp32 embeddings the fusion kernel is an fp32 computation: 1e-5 relative to the oracle."""
    from fairmultimodal_b200 import modules, ops, synth
    from oracle import fame_oracle as O
    model, w = _build_model(12, 11)
    torch.manual_seed(0)
    for B in (1, 7, 32, 100):
        e = [torch.randn(B, 768) * s for s in (1.0, 0.7, 0.5)]
        ref = O.fusion(w, *e, weights=(0.5, 0.2, 0.3))
        o = ops.fusion_fwd([x.cuda() for x in e], model._pack_fusion(), (0.5, 0.2, 0.3), True, True)
        assert _rel(o["logits"], ref["fused_logits"]) < 1e-5
        assert _rel(o["gated"], ref["gated_vector"]) < 1e-5
        assert _rel(o["pre_relu"], ref["fusion_pre_relu"]) < 1e-5
        for i, m in enumerate(("demo", "lab", "text")):
            assert _rel(o["mod_logits"][i], ref["modality_logits"][m]) < 1e-5
            assert _rel(o["proj"][:, 256 * i:256 * (i + 1)], ref["proj"][m]) < 1e-5


@pytest.mark.parametrize("B,seed", [(32, 0), (12, 1), (1, 2), (257, 3), (5000, 4)])
def test_loss_matches_oracle(B, seed):
    """fp32 loss within 1e-4 (north_star); subgroup counts bit-exact; gradient vs autograd of the oracle."""
    from fairmultimodal_b200 import ops, synth
    from oracle import fame_oracle as O
    co = synth.make_cohort(B, lab_tokens=4, chunks=0, with_tokens=False, seed=seed)
    torch.manual_seed(seed)
    z = (torch.randn(B, 3) * 2.0).requires_grad_(True)
    y = torch.from_numpy(co["labels"])
    attrs = [torch.from_numpy(co[k]) for k in ("age_ids", "ethnicity_ids", "insurance_ids")]
    sig = torch.randn(768)
    pw = torch.tensor([4.9, 1.3, 0.55])
    total, bce, leddi = O.fame_loss(z, y, attrs, sig, pw, 0.8, 0.01)
    total.backward()
    zc, yc, ac = z.detach().cuda(), y.cuda(), [a.cuda() for a in attrs]
    stats = ops.loss_stats(zc, yc, ac, pw.cuda())
    loss, dz = ops.loss_fwd_bwd(zc, yc, ac, pw.cuda(), stats, sig.cuda(), 0.8, 0.01)
    loss = loss.cpu().numpy()
    assert abs(loss[0] - float(total)) < 1e-4
    assert abs(loss[1] - float(bce)) < 1e-4
    assert abs(loss[2] - float(leddi)) < 1e-5
    counts, sums = O.loss_group_stats(z.detach(), y, attrs)
    s = stats.cpu().numpy()
    np.testing.assert_array_equal(s[78:102].reshape(3, 8), counts)                 # membership counts: bit-exact
    assert s[102] == B and s[103] == 0
    np.testing.assert_allclose(s[6:78].reshape(3, 3, 8) / 2.0 ** 24, sums, atol=1e-5 * max(1, B / 100))
    gref = z.grad
    assert (dz.cpu() - gref).abs().max().item() <= 2e-6 + 1e-3 * gref.abs().max().item()


def test_loss_stats_shard_sum_equals_global():
    """Virtual ranks: statistics of shards add up to the statistics of the concatenated batch, bit for bit."""
    from fairmultimodal_b200 import ops, synth
    co = synth.make_cohort(128, lab_tokens=4, chunks=0, with_tokens=False, seed=8)
    z = torch.randn(128, 3, device="cuda")
    y = torch.from_numpy(co["labels"]).cuda()
    attrs = [torch.from_numpy(co[k]).cuda() for k in ("age_ids", "ethnicity_ids", "insurance_ids")]
    pw = torch.tensor([4.9, 1.3, 0.55], device="cuda")
    full = ops.loss_stats(z, y, attrs, pw)
    acc = torch.zeros_like(full)
    for r in range(4):
        sl = slice(32 * r, 32 * (r + 1))
        acc += ops.loss_stats(z[sl], y[sl], [a[sl] for a in attrs], pw)
    f, a = full.cpu().numpy(), acc.cpu().numpy()
    np.testing.assert_array_equal(f, a)     # per-patient fixed point + integer sums: every statistic is bit-exact
    # and the loss evaluated from either is the same to 1e-6
    l1, _ = ops.loss_fwd_bwd(z, y, attrs, pw, full, None, 0.8, 0.0)
    l2, _ = ops.loss_fwd_bwd(z, y, attrs, pw, acc, None, 0.8, 0.0)
    assert abs(l1[0].item() - l2[0].item()) < 1e-6


def test_loss_rejects_bad_codes():
    from fairmultimodal_b200 import ops
    z = torch.zeros(4, 3, device="cuda")
    a = torch.tensor([0, 1, 9, 2], device="cuda")
    st = ops.loss_stats(z, z, [a, a, a], torch.ones(3, device="cuda"))
    assert st[103].item() == 1


def test_metrics_match_reference_golden(golden_dir):
    from fairmultimodal_b200 import metrics as M, ops
    g = _golden(golden_dir, "metrics.npz")
    z, y = torch.from_numpy(g["logits"]).cuda(), torch.from_numpy(g["labels"]).cuda()
    attrs = [torch.from_numpy(g[k]).cuda() for k in ("age", "eth", "ins")]
    sweep = torch.from_numpy(np.linspace(0, 1, 101)).cuda()
    c = M.Counts(ops.eval_counts(z, y, attrs, (0.5, 0.5, 0.5), sweep=sweep))
    th = M.thresholds_from_hist(c.hist)
    np.testing.assert_array_equal([th[n] for n in NAMES], g["thresholds"])          # calibrated thresholds bit-exact
    metrics, fair, eddi = M.evaluate_from_logits(z, y, attrs, th, verbose=False)
    np.testing.assert_allclose([metrics[n]["aucroc"] for n in NAMES], g["aucroc"], atol=1e-9)   # north_star: abs 1e-3
    np.testing.assert_allclose([metrics[n]["auprc"] for n in NAMES], g["auprc"], atol=1e-9)
    np.testing.assert_allclose([metrics[n]["f1"] for n in NAMES], g["f1"], atol=1e-12)
    np.testing.assert_allclose([metrics[n]["TPR"] for n in NAMES], g["tpr"], atol=1e-15)
    np.testing.assert_allclose([metrics[n]["fpr"] for n in NAMES], g["fpr"], atol=1e-15)
    np.testing.assert_allclose([metrics[n]["precision"] for n in NAMES], g["precision"], atol=1e-12)
    got = [[fair[n][a]["eo_metric"] for a in ("age", "ethnicity", "insurance")] for n in NAMES]
    np.testing.assert_allclose(got, g["eo"], atol=1e-14)
    np.testing.assert_allclose([fair[n]["overall_eo"] for n in NAMES], g["overall_eo"], atol=1e-14)
    got = [[eddi[n][a] for a in ("age", "ethnicity", "insurance")] for n in NAMES]
    np.testing.assert_allclose(got, g["eddi"], atol=1e-14)                           # north_star: EDDI abs 1e-3
    # weight update from modality logits
    ml = torch.from_numpy(g["mod_logits"]).cuda()
    counts = {m: M.Counts(ops.eval_counts(ml[:, 3 * i:3 * i + 3].contiguous(), y, attrs, (0.5,) * 3))
              for i, m in enumerate(M.MODALITIES)}
    w0 = {n: {m: 0.33 for m in M.MODALITIES} for n in NAMES}
    w1 = M.weights_from_modality_counts(counts, w0, 1.0, verbose=False)
    w2 = M.weights_from_modality_counts(counts, w1, 1.0, verbose=False)
    np.testing.assert_allclose([[w1[n][m] for m in M.MODALITIES] for n in NAMES], g["weights_epoch1"], atol=1e-14)
    np.testing.assert_allclose([[w2[n][m] for m in M.MODALITIES] for n in NAMES], g["weights_epoch2"], atol=1e-14)


def test_count_vector_bit_exact_vs_restatement(golden_dir):
    """All 914 integers of fame_eval_counts (confusion cells, totals, the 101-threshold F1 histogram, sample count)
    equal the numpy restatement that the CPU suite feeds to the host formulas."""
    from fairmultimodal_b200 import ops
    from oracle import count_vector as CV
    g = _golden(golden_dir, "metrics.npz")
    attrs_np = [g["age"], g["eth"], g["ins"]]
    attrs = [torch.from_numpy(a).cuda() for a in attrs_np]
    sweep = np.linspace(0, 1, 101)
    for logits, thr in ((g["logits"], (0.5, 0.5, 0.5)), (g["logits"], (0.23, 0.61, 0.5)), (g["mod_logits"][:, 3:6], (0.4, 0.5, 0.7))):
        got = ops.eval_counts(torch.from_numpy(np.ascontiguousarray(logits)).cuda(), torch.from_numpy(g["labels"]).cuda(),
                              attrs, thr, sweep=torch.from_numpy(sweep).cuda()).cpu().numpy()
        np.testing.assert_array_equal(got, CV.eval_count_vector(logits, g["labels"], attrs_np, thr, sweep=sweep))


def test_compute_eddi_dropin_and_counts_vs_oracle():
    from fairmultimodal_b200 import metrics as M, ops, synth
    from oracle import fame_oracle as O
    e, per = M.compute_eddi(np.array([0, 1, 1, 0, 1, 0]), np.array([.1, .9, .2, .8, .7, .3]),
                            np.array([0, 0, 1, 1, 2, 2]), complete_groups=[0, 1, 2, 3])
    assert e == pytest.approx(0.40824829046386296, abs=1e-15)
    assert {int(k): round(v, 12) for k, v in per.items()} == {0: -0.5, 1: 1.0, 2: -0.5}
    assert M.compute_eddi(np.array([]), np.array([]), np.array([])) == (0.0, {})
    co = synth.make_cohort(3001, lab_tokens=4, chunks=0, with_tokens=False, seed=2)
    rng = np.random.default_rng(0)
    z = torch.from_numpy(rng.standard_normal((3001, 3)).astype(np.float32) * 2)
    y = co["labels"]
    attrs_np = [co[k] for k in ("age_ids", "ethnicity_ids", "insurance_ids")]
    thr = (0.37, 0.5, 0.81)
    c = M.Counts(ops.eval_counts(z.cuda(), torch.from_numpy(y).cuda(), [torch.from_numpy(a).cuda() for a in attrs_np], thr))
    probs = torch.sigmoid(z).numpy()
    for o in range(3):
        pred = (probs[:, o] > thr[o]).astype(int)
        for a in range(3):
            for code in range(8):
                sel = attrs_np[a] == code
                _, _, cells = O.group_rates(y[:, o], pred, sel)
                assert tuple(int(v) for v in c.conf[o, a, code]) == cells          # TP, FN, FP, TN: bit-exact
        e_ref = O.compute_eddi(y[:, o], probs[:, o], attrs_np[0], thr[o], O.AGE_GROUPS)[0]
        assert M.eddi_from_counts(c.conf[o, 0], c.tot[o], O.AGE_GROUPS)[0] == pytest.approx(e_ref, abs=1e-15)
    assert c.n == 3001


def test_calculate_tpr_and_fpr_dropin():
    """10_FAME.py:84-97 semantics: rates inside the mask, 0 where a denominator is 0 (checked against the oracle)."""
    from fairmultimodal_b200 import metrics as M
    from oracle import fame_oracle as O
    rng = np.random.default_rng(4)
    y = (rng.random(997) < 0.3).astype(np.int64)
    pred = (rng.random(997) < 0.4).astype(np.int64)
    for mask in (rng.random(997) < 0.5, np.zeros(997, bool), y == 0, y == 1):
        tpr, fpr = M.calculate_tpr_and_fpr(y, pred, mask)
        rt, rf, _ = O.group_rates(y, pred, mask)
        assert tpr == rt and fpr == rf


def test_rank_metrics_ties_and_large_n():
    from fairmultimodal_b200 import metrics as M
    from oracle import fame_oracle as O
    rng = np.random.default_rng(5)
    N = 46000                                                 # BASELINE config 5 cohort size
    y = (rng.random((N, 3)) < np.array([0.1, 0.38, 0.9])).astype(np.float32)
    z = (np.round((rng.standard_normal((N, 3)) + y) * 8) / 8).astype(np.float32)   # heavy ties
    z[:, 2] = np.clip(z[:, 2] * 20, -30, 30)                  # saturation: many probabilities equal 0 or 1
    res = M.rank_metrics(torch.from_numpy(z).cuda(), torch.from_numpy(y).cuda())
    probs = torch.sigmoid(torch.from_numpy(z)).numpy()
    for o in range(3):
        # outcomes 0/1: every distinct logit maps to a distinct float32 probability -> exact agreement.
        # outcome 2 saturates: which logits collapse onto 1.0 / 0.0 depends on the last bit of the host's exp(), so
        # the tie structure may differ from torch's CPU sigmoid by a few pairs (north_star tolerance: abs 1e-3)
        tol = 1e-9 if o < 2 else 1e-5
        assert res[o][0] == pytest.approx(O.auroc(y[:, o], probs[:, o]), abs=tol)
        assert res[o][1] == pytest.approx(O.auprc(y[:, o], probs[:, o]), abs=tol)
    # size-independent properties: label flip <-> 1 - AUROC; a constant score gives exactly 0.5
    res_flip = M.rank_metrics(torch.from_numpy(z).cuda(), torch.from_numpy(1 - y).cuda())
    assert res_flip[0][0] == pytest.approx(1 - res[0][0], abs=1e-12)
    res_c = M.rank_metrics(torch.zeros(1000, 3).cuda(), torch.from_numpy(y[:1000]).cuda())
    assert res_c[0][0] == 0.5


@pytest.mark.parametrize("N", [1, 2, 777, 4096, 4097, 46000, 100003])
def test_rank_counts_sorted_equals_brute_force(N):
    """fame_rank_counts over the full range (key sort + tie-run scan, O(N log N)) produces the integers of the exact
    O(N^2) compare (run here as two sub-ranges): auroc2, positives / negatives bit-exact, the float64 AP sum to rounding."""
    from fairmultimodal_b200 import ops
    rng = np.random.default_rng(N)
    y = torch.from_numpy((rng.random(N) < 0.3).astype(np.uint8)).cuda()
    s = rng.random(N).astype(np.float32)
    s = np.round(s * 500) / 500 if N % 2 else s                 # odd sizes: heavy ties (501 distinct scores, incl. 0 and 1)
    s = torch.from_numpy(s.astype(np.float32)).cuda()
    full = ops.rank_counts(s, y)
    half = ops.rank_counts(s, y, 0, N // 2)
    half = ops.rank_counts(s, y, N // 2, N, acc=half)
    torch.cuda.synchronize()
    if N == 1:                                                   # [0, 0) is empty and [0, 1) is the full range itself
        assert int(full["pn"].sum()) == 1
        return
    assert torch.equal(full["auroc2"], half["auroc2"]) and torch.equal(full["pn"], half["pn"])
    assert abs(full["ap"].item() - half["ap"].item()) <= 1e-12 * max(1.0, abs(half["ap"].item()))


"""The drop-in ENTRY POINTS of the evaluation half (model + DataLoader in; dicts, tuples and the experiment log out)
against golden vectors made by calling the unmodified reference functions the same way
(oracle/make_golden_wrappers.py): calibrate_thresholds, evaluate_model_multi, evaluate_model,
update_dynamic_weights_all_tasks, print_fairness_metrics (10_FAME.py:99-122, 315-399, 451-557)."""
import contextlib
import io
import os
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ("mortality", "los", "mechanical_ventilation")
ATTRS = ("age", "ethnicity", "insurance")
MODS = ("demo", "lab", "text")


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


class _LogitStub(torch.nn.Module):
    """Same stand-in as oracle/make_golden.py: logits travel through the aggregated_text_embedding slot."""

    beta = 1.0

    def __init__(self):
        super().__init__()
        self.calls = []

    def forward(self, a, b, c, d, e, f, g, text, **kw):
        self.calls.append((self.training, torch.is_grad_enabled(), dict(kw)))
        return {"fused_logits": text[:, 0:3].contiguous(),
                "modality_logits": {"demo": text[:, 3:6], "lab": text[:, 6:9], "text": text[:, 9:12]}}


def _loader(golden_dir):
    from fairmultimodal_b200 import synth
    g = np.load(os.path.join(golden_dir, "metrics.npz"), allow_pickle=False)
    N = g["logits"].shape[0]
    co = synth.make_cohort(N, lab_tokens=4, chunks=0, with_tokens=False, seed=5)
    text = np.concatenate([g["logits"], g["mod_logits"]], axis=1).astype(np.float32)
    t = lambda k: torch.from_numpy(co[k])
    ds = torch.utils.data.TensorDataset(t("demo_dummy_ids"), t("demo_attn_mask"), t("age_ids"), t("gender_ids"),
                                        t("ethnicity_ids"), t("insurance_ids"), t("lab_features"),
                                        torch.from_numpy(text), t("labels"))
    return torch.utils.data.DataLoader(ds, batch_size=64, shuffle=False)     # 23 full batches + a ragged one of 28


def _capture(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **kw)
    return r, buf.getvalue()


def _check_eval(prefix, res, w):
    assert isinstance(res, tuple) and len(res) == 7                      # the reference's 7-tuple
    metrics, lg, lb, ag, et, ins, fair = res
    for k in ("aucroc", "auprc", "f1", "recall (TPR)", "TPR", "precision", "fpr", "optimal_threshold"):
        ref = w[f"{prefix}_{k.split(' ')[0]}"]
        got = np.array([metrics[n][k] for n in NAMES], dtype=np.float64)
        np.testing.assert_allclose(got, ref, atol=1e-9 if k in ("aucroc", "auprc") else 1e-12, rtol=0, err_msg=k)
    assert set(metrics["los"].keys()) == {"aucroc", "auprc", "f1", "recall (TPR)", "TPR", "precision", "fpr",
                                          "optimal_threshold"}
    for k in ("avg_tpr_diff", "avg_fpr_diff", "eo_metric"):
        got = np.array([[fair[n][a][k] for a in ATTRS] for n in NAMES])
        np.testing.assert_allclose(got, w[f"{prefix}_{k}"], atol=1e-14, rtol=0, err_msg=k)
    np.testing.assert_allclose([fair[n]["overall_eo"] for n in NAMES], w[f"{prefix}_overall_eo"], atol=1e-14, rtol=0)
    # returned arrays: numpy, host, same dtypes / shapes / values (logits and labels float32 [N,3]; codes int64 [N])
    for got, key in ((lg, "ret_logits"), (lb, "ret_labels"), (ag, "ret_age"), (et, "ret_eth"), (ins, "ret_ins")):
        ref = w[f"{prefix}_{key}"]
        assert isinstance(got, np.ndarray) and got.dtype == ref.dtype and got.shape == ref.shape, key
        np.testing.assert_array_equal(got, ref)


def test_calibrate_and_evaluate_entry_points(golden_dir):
    from fairmultimodal_b200 import metrics as M
    w = np.load(os.path.join(golden_dir, "metric_wrappers.npz"), allow_pickle=False)
    loader = _loader(golden_dir)
    stub = _LogitStub().cuda().train()
    th = M.calibrate_thresholds(stub, loader, "cuda")
    assert stub.training is False                                        # switched to eval(), as the reference does
    assert all(c[0] is False and c[1] is False and c[2] == {} for c in stub.calls)   # eval mode, no_grad, no kwargs
    assert list(th.keys()) == list(NAMES)
    np.testing.assert_array_equal([th[n] for n in NAMES], w["thresholds"])   # bit-exact (first strict F1 maximum)
    res, txt = _capture(M.evaluate_model_multi, stub, loader, "cuda", th)
    _check_eval("multi", res, w)
    assert txt == str(w["multi_stdout"])                                 # the experiment log, byte for byte
    stub.train()
    res, txt = _capture(M.evaluate_model, stub, loader, "cuda", threshold=0.5, old_eddi_weights={"ignored": 1})
    assert stub.training is False
    _check_eval("single", res, w)
    assert txt == str(w["single_stdout"])


_FLOAT = re.compile(r"np\.float64\(([^)]*)\)")


def _split_repr_floats(text):
    """The weight-update log prints a dict of np.float64: compare its numbers to 1e-12 and the rest exactly."""
    nums = [float(x) for x in _FLOAT.findall(text)]
    return _FLOAT.sub("np.float64(#)", text), nums


def test_update_dynamic_weights_entry_point(golden_dir):
    from fairmultimodal_b200 import metrics as M
    w = np.load(os.path.join(golden_dir, "metric_wrappers.npz"), allow_pickle=False)
    loader = _loader(golden_dir)
    stub = _LogitStub().cuda().train()
    w0 = {n: {m: 0.33 for m in MODS} for n in NAMES}
    w1, txt1 = _capture(M.update_dynamic_weights_all_tasks, stub, loader, "cuda", w0, beta=1.0)
    assert stub.training is True                                         # NOT switched to eval() (reference quirk)
    # forward called with the reference's keyword arguments (10_FAME.py:327-331)
    assert all(set(c[2].keys()) == {"beta", "old_eddi_weights", "return_modality_logits"} and
               c[2]["return_modality_logits"] is True and c[2]["beta"] == 1.0 and c[2]["old_eddi_weights"] is w0
               for c in stub.calls)
    w2, txt2 = _capture(M.update_dynamic_weights_all_tasks, stub, loader, "cuda", w1, beta=0.5, threshold=0.4)
    for got, key, txt in ((w1, "weights1", txt1), (w2, "weights2", txt2)):
        assert list(got.keys()) == list(NAMES) and all(list(got[n].keys()) == list(MODS) for n in NAMES)
        np.testing.assert_allclose([[got[n][m] for m in MODS] for n in NAMES], w[key], atol=1e-14, rtol=0)
        a, an = _split_repr_floats(txt)
        b, bn = _split_repr_floats(str(w[key + "_stdout"]))
        assert a == b
        np.testing.assert_allclose(an, bn, atol=1e-12, rtol=0)


def test_print_fairness_metrics_entry_point(golden_dir):
    from fairmultimodal_b200 import metrics as M
    g = np.load(os.path.join(golden_dir, "metrics.npz"), allow_pickle=False)
    w = np.load(os.path.join(golden_dir, "metric_wrappers.npz"), allow_pickle=False)
    probs = torch.sigmoid(torch.from_numpy(g["logits"]))[:, 1].numpy()
    preds = (probs > 0.5).astype(int)
    r, txt = _capture(M.print_fairness_metrics, g["labels"][:, 1], preds, g["eth"], "ethnicity")
    np.testing.assert_allclose(np.array(r, dtype=np.float64), w["pfm"], atol=1e-14, rtol=0)
    assert txt == str(w["pfm_stdout"])

"""Golden vectors for the DROP-IN ENTRY POINTS of the evaluation half of 10_FAME.py, produced by calling the
unmodified reference functions exactly as run_experiment does (model + DataLoader in, dicts / tuples / stdout out):

    calibrate_thresholds(model, loader, device)                      10_FAME.py:451-482
    evaluate_model_multi(model, loader, device, thresholds)          10_FAME.py:484-552
    evaluate_model(model, loader, device, threshold=0.5)             10_FAME.py:554-557  (scalar threshold)
    update_dynamic_weights_all_tasks(model, loader, device, w, beta) 10_FAME.py:315-399  (model left in train() mode)
    print_fairness_metrics(y, preds, demographics, name)             10_FAME.py:99-122

Run in the build container only (needs /root/reference):   python oracle/make_golden_wrappers.py
TEST INFRASTRUCTURE ONLY.  The model is make_golden._LogitStub (controlled logits travel through the text slot); the
cohort / logits are the ones of tests/golden/metrics.npz, batches of 64 with a ragged last batch of 28.  Besides the
returned values the fixture keeps the reference's stdout, which is part of its contract (the experiment log).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fairmultimodal_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import _LogitStub, make_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "metric_wrappers.npz")
NAMES = ("mortality", "los", "mechanical_ventilation")
ATTRS = ("age", "ethnicity", "insurance")


def _capture(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **kw)
    return r, buf.getvalue()


def _pack_eval(prefix, res, out):
    metrics, lg, lb, ag, et, ins, fair = res
    for k in ("aucroc", "auprc", "f1", "recall (TPR)", "TPR", "precision", "fpr", "optimal_threshold"):
        out[f"{prefix}_{k.split(' ')[0]}"] = np.array([metrics[n][k] for n in NAMES], dtype=np.float64)
    for k in ("avg_tpr_diff", "avg_fpr_diff", "eo_metric"):
        out[f"{prefix}_{k}"] = np.array([[fair[n][a][k] for a in ATTRS] for n in NAMES], dtype=np.float64)
    out[f"{prefix}_overall_eo"] = np.array([fair[n]["overall_eo"] for n in NAMES], dtype=np.float64)
    out[f"{prefix}_ret_logits"], out[f"{prefix}_ret_labels"] = lg, lb
    out[f"{prefix}_ret_age"], out[f"{prefix}_ret_eth"], out[f"{prefix}_ret_ins"] = ag, et, ins


def main():
    ref = ref_loader.load()
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    N = g["logits"].shape[0]
    co = synth.make_cohort(N, lab_tokens=4, chunks=0, with_tokens=False, seed=5)
    assert np.array_equal(co["labels"], g["labels"]) and np.array_equal(co["age_ids"], g["age"])
    text = np.concatenate([g["logits"], g["mod_logits"]], axis=1).astype(np.float32)
    loader = make_loader(ref, co, text, 64)
    stub = _LogitStub()
    out = {}
    th, _ = _capture(ref.calibrate_thresholds, stub, loader, "cpu")
    assert stub.training is False                      # calibrate_thresholds switched the model to eval()
    out["thresholds"] = np.array([th[n] for n in NAMES])
    res, txt = _capture(ref.evaluate_model_multi, stub, loader, "cpu", th)
    _pack_eval("multi", res, out)
    out["multi_stdout"] = np.array(txt)
    res, txt = _capture(ref.evaluate_model, stub, loader, "cpu", threshold=0.5)
    _pack_eval("single", res, out)
    out["single_stdout"] = np.array(txt)
    stub.train()
    w0 = {n: {"demo": 0.33, "lab": 0.33, "text": 0.33} for n in NAMES}
    w1, txt1 = _capture(ref.update_dynamic_weights_all_tasks, stub, loader, "cpu", w0, beta=1.0)
    assert stub.training is True                       # the reference does NOT switch to eval() here (A.3-11)
    w2, txt2 = _capture(ref.update_dynamic_weights_all_tasks, stub, loader, "cpu", w1, beta=0.5, threshold=0.4)
    out["weights1"] = np.array([[w1[n][m] for m in ("demo", "lab", "text")] for n in NAMES])
    out["weights2"] = np.array([[w2[n][m] for m in ("demo", "lab", "text")] for n in NAMES])
    out["weights1_stdout"], out["weights2_stdout"] = np.array(txt1), np.array(txt2)
    # print_fairness_metrics called directly on 0/1 predictions
    probs = torch.sigmoid(torch.from_numpy(g["logits"]))[:, 1].numpy()
    preds = (probs > 0.5).astype(int)
    r, txt = _capture(ref.print_fairness_metrics, g["labels"][:, 1], preds, g["eth"], "ethnicity")
    out["pfm"] = np.array(r, dtype=np.float64)
    out["pfm_stdout"] = np.array(txt)
    np.savez_compressed(OUT, **out)
    print("metric_wrappers.npz", {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()

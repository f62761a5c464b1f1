"""Fairness / evaluation functions of 10_FAME.py on the B200: the data passes (thresholding, per-subgroup confusion
counts, the 101-threshold F1 sweep, AUROC / AUPRC rank statistics) run in the CUDA metric kernels and produce INTEGER
counts; the handful of float64 divisions that turn counts into EDDI / EO / F1 / AUROC / AP are done on the host, in
the same order as the reference.

    compute_eddi                      10_FAME.py:54-82        print_fairness_metrics   10_FAME.py:99-122
    update_dynamic_weights_all_tasks  10_FAME.py:315-399      calibrate_thresholds     10_FAME.py:451-482
    evaluate_model_multi / evaluate_model  10_FAME.py:484-557

Subgroup codes must lie in 0..7 (the reference's codes are 0..5); other codes raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

OUTCOMES = ("mortality", "los", "mechanical_ventilation")
MODALITIES = ("demo", "lab", "text")
ATTRS = ("age", "ethnicity", "insurance")
AGE_GROUPS, ETH_GROUPS, INS_GROUPS = (0, 1, 2, 3), (0, 1, 2, 3, 4), (0, 1, 2, 3, 4, 5)   # 10_FAME.py:353-355
_SWEEP = np.linspace(0, 1, 101)                                                        # 10_FAME.py:475


class Counts:
    """View over the uint64 vector written by fame_eval_counts (layout: include/fame_b200.h)."""

    def __init__(self, vec):
        v = vec.cpu().numpy().astype(np.int64) if isinstance(vec, torch.Tensor) else np.asarray(vec, dtype=np.int64)
        if v[913]:
            raise ValueError("sensitive-attribute code outside 0..7")
        self.conf = v[:288].reshape(3, 3, 8, 4)          # [outcome, attr, code, (TP, FN, FP, TN)]
        self.tot = v[288:300].reshape(3, 4)              # [outcome, (TP, FN, FP, TN)]
        self.hist = v[300:912].reshape(3, 2, 102)        # [outcome, label, #sweep thresholds strictly below p]
        self.n = int(v[912])


def eddi_from_counts(conf_oa, tot_o, complete_groups=None):
    """compute_eddi (10_FAME.py:54-82) from integer counts.  conf_oa [8,4] for one (outcome, attr); tot_o [4]."""
    n = int(tot_o.sum())
    overall = (int(tot_o[1]) + int(tot_o[2])) / n                   # mean(pred != y): FN + FP
    den = 1 - overall if overall < 0.5 else overall
    sizes = conf_oa.sum(axis=1)
    groups = [g for g in range(8) if sizes[g] > 0] if complete_groups is None else [g for g in complete_groups]
    sub = {}
    for g in groups:
        if g < 0 or g >= 8 or sizes[g] == 0:
            continue
        sub[g] = ((int(conf_oa[g, 1]) + int(conf_oa[g, 2])) / int(sizes[g]) - overall) / den
    if not sub:
        return 0.0, sub
    return float(np.sqrt(np.sum(np.array(list(sub.values())) ** 2)) / len(sub)), sub


def eo_from_counts(conf_oa):
    """print_fairness_metrics (10_FAME.py:99-122) from counts: groups = codes present."""
    tpr, fpr = [], []
    for g in range(8):
        tp, fn, fp, tn = (int(x) for x in conf_oa[g])
        if tp + fn + fp + tn == 0:
            continue
        tpr.append(tp / (tp + fn) if tp + fn > 0 else 0)
        fpr.append(fp / (fp + tn) if fp + tn > 0 else 0)
    dt = [abs(tpr[i] - tpr[j]) for i in range(len(tpr)) for j in range(i + 1, len(tpr))]
    df = [abs(fpr[i] - fpr[j]) for i in range(len(fpr)) for j in range(i + 1, len(fpr))]
    a = float(np.mean(dt)) if dt else 0.0
    b = float(np.mean(df)) if df else 0.0
    return a, b, (a + b) / 2.0, tpr, fpr


def _dev(x, dtype, device):
    t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
    return t.to(device=device, dtype=dtype).contiguous()


def compute_eddi(y_true, y_pred, sensitive_labels, threshold=0.5, complete_groups=None, device="cuda"):
    """Drop-in for 10_FAME.py:54-82 (numpy in, (float, dict) out); the counting runs on the GPU.
    y_pred holds scores (float32 probabilities, or 0/1 predictions as in update_dynamic_weights_all_tasks)."""
    y_true = np.asarray(y_true)
    n = y_true.shape[0]
    if n == 0:
        return 0.0, {}
    sens = np.asarray(sensitive_labels)
    scores = np.asarray(y_pred)
    # numpy compares a float32 array with a python float in float32 and with an np.float64 scalar in float64;
    # float64 scores are compared in float64 either way -> carry the scores at their own precision
    if scores.dtype == np.float64:
        pred = (scores > threshold).astype(np.float32)             # host thresholding keeps float64 semantics
        probs, thr = pred, 0.5
    else:
        probs = scores.astype(np.float32)
        thr = float(np.float32(threshold)) if isinstance(threshold, float) and not isinstance(threshold, np.floating) else float(threshold)
    lab3 = np.zeros((n, 3), np.float32)
    lab3[:, 0] = y_true
    pr3 = np.zeros((n, 3), np.float32)
    pr3[:, 0] = probs
    attrs = [_dev(sens, torch.int64, device)] * 3
    vec = ops.eval_counts(_dev(pr3, torch.float32, device), _dev(lab3, torch.float32, device), attrs,
                          (thr, 2.0, 2.0), logits_are_probs=True)
    c = Counts(vec)
    groups = None if complete_groups is None else [int(g) for g in np.asarray(complete_groups)]
    e, sub = eddi_from_counts(c.conf[0, 0], c.tot[0], groups)
    keyt = (lambda g: g) if complete_groups is None else (lambda g: np.asarray(complete_groups).dtype.type(g))
    return e, {keyt(g): v for g, v in sub.items()}


def calculate_tpr_and_fpr(y_true, y_pred, group_mask, device="cuda"):
    """Drop-in for 10_FAME.py:84-97: (TPR, FPR) of the 0/1 predictions inside `group_mask`; 0 where a denominator
    is 0.  The four confusion counts come from the count kernel (membership code 1 = inside the mask)."""
    n = len(y_true)
    lab3 = np.zeros((n, 3), np.float32)
    lab3[:, 0] = np.asarray(y_true)
    pr3 = np.zeros((n, 3), np.float32)
    pr3[:, 0] = np.asarray(y_pred)
    member = _dev(np.asarray(group_mask).astype(np.int64), torch.int64, device)
    c = Counts(ops.eval_counts(_dev(pr3, torch.float32, device), _dev(lab3, torch.float32, device), [member] * 3,
                               (0.5, 2.0, 2.0), logits_are_probs=True))
    tp, fn, fp, tn = (int(v) for v in c.conf[0, 0, 1])
    tpr = tp / (tp + fn) if (tp + fn) > 0 else 0
    fpr = fp / (fp + tn) if (fp + tn) > 0 else 0
    return tpr, fpr


def print_fairness_metrics(y_true, y_pred, demographics, sensitive_attr_name, device="cuda", verbose=True):
    """Drop-in for 10_FAME.py:99-122: y_pred are 0/1 predictions."""
    n = len(y_true)
    lab3 = np.zeros((n, 3), np.float32)
    lab3[:, 0] = np.asarray(y_true)
    pr3 = np.zeros((n, 3), np.float32)
    pr3[:, 0] = np.asarray(y_pred)
    attrs = [_dev(demographics, torch.int64, device)] * 3
    c = Counts(ops.eval_counts(_dev(pr3, torch.float32, device), _dev(lab3, torch.float32, device), attrs,
                               (0.5, 2.0, 2.0), logits_are_probs=True))
    a, b, eo, tpr, fpr = eo_from_counts(c.conf[0, 0])
    if verbose:
        print(f"Fairness metrics for sensitive attribute: {sensitive_attr_name}")
        present = [g for g in range(8) if c.conf[0, 0, g].sum() > 0]
        for g, t, f in zip(present, tpr, fpr):
            print(f"  Group {g}: TPR = {t:.3f}, FPR = {f:.3f}")
        print(f"  Average TPR difference across groups: {a:.3f}")
        print(f"  Average FPR difference across groups: {b:.3f}")
        print(f"  EO fairness metric (average of TPR and FPR differences): {eo:.3f}\n")
    return a, b, eo


# ------------------------------------------------------------------------------------------------ device-resident
def _f1(tp, fp, fn):
    d = 2 * tp + fp + fn
    return 0.0 if d == 0 else 2.0 * tp / d


def thresholds_from_hist(hist):
    """First strict F1 maximum over linspace(0, 1, 101), starting from (0.5, 0.0) (10_FAME.py:473-481).
    hist [3, 2, 102]: hist[o, y, k] = samples of label y with exactly k sweep thresholds strictly below p."""
    out = {}
    for o, name in enumerate(OUTCOMES):
        pos, neg = hist[o, 1], hist[o, 0]
        # prediction at threshold index k is positive iff kk > k
        tp_k = pos[::-1].cumsum()[::-1]          # tp_k[k] = sum_{kk >= k} pos[kk]
        fp_k = neg[::-1].cumsum()[::-1]
        npos = int(pos.sum())
        best_t, best_f = 0.5, 0.0
        for k in range(101):
            tp, fp = int(tp_k[k + 1]), int(fp_k[k + 1])
            f = _f1(tp, fp, npos - tp)
            if f > best_f:
                best_f, best_t = f, _SWEEP[k]
        out[name] = best_t
    return out


def rank_metrics(logits, labels, group=None):
    """AUROC / average precision per outcome from exact rank counts (sklearn semantics on float32 probabilities):
    fame_rank_counts over the full range = key sort + tie-run scan, O(N log N).  With a process group the caller has
    all-gathered logits / labels, so every rank holds the whole cohort and computes the same integers itself (a sort of
    46 k keys costs less than an all-reduce would); `group` is accepted for interface stability."""
    probs, y8 = ops.sigmoid_probs(logits, labels)
    accs = [ops.rank_counts(probs[o], y8[o]) for o in range(3)]
    au = torch.cat([a["auroc2"] for a in accs] + [a["pn"] for a in accs]).cpu().numpy()
    ap = torch.cat([a["ap"] for a in accs]).cpu().numpy()
    res = []
    for o in range(3):
        npos, nneg = int(au[3 + 2 * o]), int(au[3 + 2 * o + 1])
        auroc = float("nan") if npos == 0 or nneg == 0 else 1.0 - int(au[o]) / (2.0 * npos * nneg)
        res.append((auroc, float(ap[o]) / npos if npos else 0.0))
    return res


@torch.no_grad()
def _collect(model, dataloader, device, want_modality=False, old_eddi_weights=None, beta=None):
    logits, mods, labels, attrs = [], [], [], [[], [], []]
    for batch in dataloader:
        b = [x.to(device, non_blocking=True) for x in batch]
        kw = {}
        if want_modality:
            kw = dict(beta=beta, old_eddi_weights=old_eddi_weights, return_modality_logits=True)
        out = model(*b[:8], **kw)
        logits.append(out["fused_logits"])
        if want_modality:
            mods.append(torch.stack([out["modality_logits"][m] for m in MODALITIES]))
        labels.append(b[8].float())
        for k, idx in enumerate((2, 4, 5)):
            attrs[k].append(b[idx])
    logits, labels = torch.cat(logits), torch.cat(labels)
    attrs = [torch.cat(a).to(torch.int64) for a in attrs]
    mods = torch.cat(mods, dim=1) if want_modality else None
    return logits, mods, labels, attrs


def calibrate_thresholds(model, dataloader, device):
    """Drop-in for 10_FAME.py:451-482."""
    model.eval()
    logits, _, labels, attrs = _collect(model, dataloader, device)
    sweep = torch.from_numpy(_SWEEP).to(device)
    c = Counts(ops.eval_counts(logits, labels, attrs, (0.5, 0.5, 0.5), sweep=sweep))
    return thresholds_from_hist(c.hist)


def evaluate_from_logits(logits, labels, attrs, thresholds, verbose=True, counts=None, rank_group=None, ranks=None):
    """Metric half of evaluate_model_multi (10_FAME.py:511-552) + the EDDI tail of run_experiment (887-915) on
    device tensors.  Returns (metrics, fairness_details, eddi).  `counts`: a precomputed (e.g. all-reduced) count
    vector; `rank_group`: process group over which the AUROC / AP rank counting of `logits` is sharded; `ranks`:
    precomputed [(auroc, ap)] per outcome (then `logits` is not touched at all)."""
    th = [thresholds[n] if isinstance(thresholds, dict) else thresholds for n in OUTCOMES]
    vec = counts if counts is not None else ops.eval_counts(logits, labels, attrs, th)
    c = Counts(vec)
    if ranks is None:
        ranks = rank_metrics(logits, labels, group=rank_group)
    metrics, fair, eddi = {}, {}, {}
    for o, name in enumerate(OUTCOMES):
        tp, fn, fp, tn = (int(x) for x in c.tot[o])
        metrics[name] = {"aucroc": ranks[o][0], "auprc": ranks[o][1], "f1": _f1(tp, fp, fn),
                         "recall (TPR)": tp / (tp + fn) if tp + fn else 0.0, "TPR": tp / (tp + fn) if tp + fn else 0,
                         "precision": tp / (tp + fp) if tp + fp else 0.0, "fpr": fp / (fp + tn) if fp + tn else 0,
                         "optimal_threshold": th[o]}
        fair[name] = {}
        if verbose:
            print(f"\nOutcome: {name} (Threshold: {th[o]:.2f})")
        eos = []
        for a, an in enumerate(ATTRS):
            dt, df, eo, tpr, fpr = eo_from_counts(c.conf[o, a])
            if verbose:
                print(f"Fairness metrics for sensitive attribute: {an}")
                present = [g for g in range(8) if c.conf[o, a, g].sum() > 0]
                for g, t, f in zip(present, tpr, fpr):
                    print(f"  Group {g}: TPR = {t:.3f}, FPR = {f:.3f}")
                print(f"  Average TPR difference across groups: {dt:.3f}")
                print(f"  Average FPR difference across groups: {df:.3f}")
                print(f"  EO fairness metric (average of TPR and FPR differences): {eo:.3f}\n")
            fair[name][an] = {"avg_tpr_diff": dt, "avg_fpr_diff": df, "eo_metric": eo}
            eos.append(eo)
        fair[name]["overall_eo"] = float(np.mean(eos))
        if verbose:
            print(f"Overall EO fairness metric for outcome {name}: {fair[name]['overall_eo']:.3f}")
        parts = [eddi_from_counts(c.conf[o, a], c.tot[o], gl)[0]
                 for a, gl in enumerate((AGE_GROUPS, ETH_GROUPS, INS_GROUPS))]
        eddi[name] = {"age": parts[0], "ethnicity": parts[1], "insurance": parts[2],
                      "combined": float(np.sqrt(sum(p * p for p in parts)) / 3.0)}
    eddi["overall"] = float(np.mean([eddi[n]["combined"] for n in OUTCOMES]))
    return metrics, fair, eddi


def evaluate_model_multi(model, dataloader, device, thresholds, print_eddi=False):
    """Drop-in for 10_FAME.py:484-552 (same 7-tuple)."""
    model.eval()
    logits, _, labels, attrs = _collect(model, dataloader, device)
    metrics, fair, _ = evaluate_from_logits(logits, labels, attrs, thresholds, verbose=True)
    return (metrics, logits.cpu().numpy(), labels.cpu().numpy(), attrs[0].cpu().numpy().squeeze(),
            attrs[1].cpu().numpy().squeeze(), attrs[2].cpu().numpy().squeeze(), fair)


def evaluate_model(model, dataloader, device, threshold=0.5, old_eddi_weights=None):
    """Drop-in for 10_FAME.py:554-557 (old_eddi_weights is ignored there too)."""
    return evaluate_model_multi(model, dataloader, device, thresholds=threshold, print_eddi=True)


def weights_from_modality_counts(counts_by_modality, old_eddi_weights, beta, verbose=True):
    """The weight update of 10_FAME.py:357-397 from per-modality confusion counts (threshold 0.5)."""
    new = {}
    for o, name in enumerate(OUTCOMES):
        e = {}
        for m in MODALITIES:
            c = counts_by_modality[m]
            parts = [eddi_from_counts(c.conf[o, a], c.tot[o], gl)[0]
                     for a, gl in enumerate((AGE_GROUPS, ETH_GROUPS, INS_GROUPS))]
            e[m] = np.sqrt(parts[0] ** 2 + parts[1] ** 2 + parts[2] ** 2) / 3.0
        top = max(e["demo"], e["lab"], e["text"])
        if verbose:
            print(f"[{name} Weight Update] EDDI:")
            print("  Demo modality - Overall EDDI: {:.4f}".format(e["demo"]))
            print("  Lab modality  - Overall EDDI: {:.4f}".format(e["lab"]))
            print("  Text modality - Overall EDDI: {:.4f}".format(e["text"]))
            print("  Maximum EDDI among modalities: {:.4f}".format(top))
        prev = old_eddi_weights.get(name, {"demo": 0.33, "lab": 0.33, "text": 0.33})
        raw = {m: max(prev[m] + np.clip(beta * (top - e[m]), -0.05, 0.05), 0.1) for m in MODALITIES}
        tot = raw["demo"] + raw["lab"] + raw["text"]
        new[name] = {m: raw[m] / tot for m in MODALITIES}
        if verbose:
            print(f"[{name} Weight Update] New dynamic weights: {new[name]}\n")
    return new


def update_dynamic_weights_all_tasks(model, dataloader, device, old_eddi_weights, beta, threshold=0.5):
    """Drop-in for 10_FAME.py:315-399.  As in the reference the model is NOT switched to eval mode here; the
    forward runs under no_grad."""
    _, mods, labels, attrs = _collect(model, dataloader, device, want_modality=True,
                                      old_eddi_weights=old_eddi_weights, beta=getattr(model, "beta", beta))
    counts = {m: Counts(ops.eval_counts(mods[i].contiguous(), labels, attrs, (threshold,) * 3))
              for i, m in enumerate(MODALITIES)}
    return weights_from_modality_counts(counts, old_eddi_weights, beta)

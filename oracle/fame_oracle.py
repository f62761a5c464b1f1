"""CPU oracle for the FAME hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 restatement (torch CPU tensors for the encoders, numpy for the integer / metric work) of what the
reference computes on the path named by BASELINE.json.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package (fairmultimodal_b200/) never
does and has no CPU path at all.

Reference = AI-for-Health-Data/FairMultimodal, FinalCode/New/Final/10_FAME.py ("FAME:<line>" below) and the
third-party modules it calls, which the reference does not vendor or pin:
  * transformers.models.bert.modeling_bert (installed here: 5.5.0; "HF:<line>")
  * torch.nn.TransformerEncoderLayer / MultiheadAttention (torch 2.11.0)
  * sklearn.metrics roc_auc_score / average_precision_score / f1_score (sklearn 1.9.0)

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the unmodified 10_FAME.py in the build
container, runs it on seeded synthetic inputs and stores inputs + outputs under tests/golden/;
tests/test_oracle_golden.py checks every function below against those files (and, where /root/reference is
present, against the live reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

OUTCOMES = ("mortality", "los", "mechanical_ventilation")
MODALITIES = ("demo", "lab", "text")
AGE_GROUPS = (0, 1, 2, 3)            # FAME:353
ETH_GROUPS = (0, 1, 2, 3, 4)         # FAME:354
INS_GROUPS = (0, 1, 2, 3, 4, 5)      # FAME:355


# ------------------------------------------------------------------------------------------------ encoders
def layer_norm(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def bert_encode(sd, prefix, input_ids, attention_mask, num_layers=12, num_heads=12, eps=1e-12):
    """last_hidden_state of a HF BertModel in eval mode (HF:102-112 embeddings, HF:179-206 attention,
    HF:294-298 / 339-342 / 352-356 MLP blocks).  sd: state_dict, prefix e.g. 'BioBert.' or 'behrt_demo.bert.'."""
    g = lambda k: sd[prefix + k].float()
    B, S = input_ids.shape
    # nn.Embedding(padding_idx=pad_token_id=0): row 0 is looked up normally but never receives gradient
    x = F.embedding(input_ids, g("embeddings.word_embeddings.weight"), padding_idx=0)
    x = x + g("embeddings.token_type_embeddings.weight")[0]
    x = x + g("embeddings.position_embeddings.weight")[:S][None]
    x = layer_norm(x, g("embeddings.LayerNorm.weight"), g("embeddings.LayerNorm.bias"), eps)
    H = x.shape[-1]
    D = H // num_heads
    # additive key bias: 0 for attended keys, -inf otherwise (HF:709-713 builds the equivalent sdpa mask)
    bias = torch.zeros(B, 1, 1, S, device=input_ids.device)
    bias = bias.masked_fill(attention_mask[:, None, None, :] == 0, float("-inf"))
    for i in range(num_layers):
        p = f"encoder.layer.{i}."
        q = F.linear(x, g(p + "attention.self.query.weight"), g(p + "attention.self.query.bias"))
        k = F.linear(x, g(p + "attention.self.key.weight"), g(p + "attention.self.key.bias"))
        v = F.linear(x, g(p + "attention.self.value.weight"), g(p + "attention.self.value.bias"))
        q, k, v = (t.view(B, S, num_heads, D).transpose(1, 2) for t in (q, k, v))
        s = q @ k.transpose(-1, -2) * (D ** -0.5) + bias
        ctx = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, S, H)
        a = F.linear(ctx, g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias"))
        x = layer_norm(a + x, g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias"), eps)
        h = F.gelu(F.linear(x, g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")))
        o = F.linear(h, g(p + "output.dense.weight"), g(p + "output.dense.bias"))
        x = layer_norm(o + x, g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"), eps)
    return x


def note_cls(sd, input_ids, attention_mask, prefix="BioBert."):
    """BioClinicalBERT_FT.forward (FAME:139-142): CLS row of the last hidden state."""
    return bert_encode(sd, prefix, input_ids, attention_mask)[:, 0, :]


def pool_patient_notes(cls_rows, offsets, hidden=768):
    """Chunk -> patient mean (FAME:153-154, 170-172): np.mean over the patient's CLS rows, zeros if none.
    cls_rows: float32 [C, hidden] in chunk order; offsets: int [P+1] CSR.  Returns float32 [P, hidden]
    (the reference casts to float32 at FAME:731)."""
    cls_rows = np.asarray(cls_rows, dtype=np.float32)
    P = len(offsets) - 1
    out = np.zeros((P, hidden), dtype=np.float32)
    for p in range(P):
        a, b = int(offsets[p]), int(offsets[p + 1])
        if b > a:
            out[p] = np.mean(cls_rows[a:b], axis=0)
    return out


def behrt_demo(sd, input_ids, attention_mask, age, gender, eth, ins, prefix="behrt_demo.", num_layers=12, num_heads=12):
    """BEHRTModel_Demo.forward (FAME:194-206)."""
    tab = {n: sd[prefix + n + "_embedding.weight"].float() for n in ("age", "gender", "ethnicity", "insurance")}
    cl = lambda ids, t: ids.clamp(0, t.shape[0] - 1)
    cls = bert_encode(sd, prefix + "bert.", input_ids, attention_mask, num_layers=num_layers, num_heads=num_heads)[:, 0, :]
    extra = (tab["age"][cl(age, tab["age"])] + tab["gender"][cl(gender, tab["gender"])]
             + tab["ethnicity"][cl(eth, tab["ethnicity"])] + tab["insurance"][cl(ins, tab["insurance"])]) / 4.0
    return cls + extra


def behrt_lab(sd, lab, prefix="behrt_lab.", nhead=8, num_layers=2, eps=1e-5):
    """BEHRTModel_Lab.forward (FAME:217-224) with nn.TransformerEncoderLayer defaults (post-norm, ReLU, ff 2048)
    in eval mode.  lab: f32 [B, L]."""
    g = lambda k: sd[prefix + k].float()
    B, L = lab.shape
    x = lab[..., None] * g("token_embedding.weight")[:, 0] + g("token_embedding.bias")
    x = x + g("pos_embedding")[None]
    H = x.shape[-1]
    D = H // nhead
    for i in range(num_layers):
        p = f"transformer_encoder.layers.{i}."
        qkv = F.linear(x, g(p + "self_attn.in_proj_weight"), g(p + "self_attn.in_proj_bias"))
        q, k, v = (t.reshape(B, L, nhead, D).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        s = q @ k.transpose(-1, -2) / math.sqrt(D)
        ctx = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, H)
        a = F.linear(ctx, g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"))
        x = layer_norm(x + a, g(p + "norm1.weight"), g(p + "norm1.bias"), eps)
        f = F.linear(torch.relu(F.linear(x, g(p + "linear1.weight"), g(p + "linear1.bias"))),
                     g(p + "linear2.weight"), g(p + "linear2.bias"))
        x = layer_norm(x + f, g(p + "norm2.weight"), g(p + "norm2.bias"), eps)
    return x.mean(dim=1)


def behrt_combined(sd, lab):
    """BEHRTModel_Combined.forward (01_BEHRT.py:122-131), eval mode: logits f32 [B, 3] (mort, los, mech)."""
    e = behrt_lab(sd, lab, prefix="lab_model.")
    fused = F.linear(e, sd["fusion_fc.weight"].float(), sd["fusion_fc.bias"].float())
    w = torch.cat([sd[h + ".weight"].float() for h in ("classifier_mort", "classifier_los", "classifier_mech")])
    b = torch.cat([sd[h + ".bias"].float() for h in ("classifier_mort", "classifier_los", "classifier_mech")])
    return F.linear(fused, w, b)


def behrt_combined_loss(logits, labels, pos_weight):
    """Training objective of 01_BEHRT.py:218-222: the SUM of three BCEWithLogitsLoss(pos_weight_i) batch means."""
    total = 0.0
    for i in range(3):
        total = total + F.binary_cross_entropy_with_logits(logits[:, i], labels[:, i], pos_weight=pos_weight[i])
    return total


def sigmoid_fusion_forward(sd, batch8):
    """MultimodalTransformer.forward of 09_multimodal_sigmoid_fusion.py:186-222, eval mode: (logits [B,3], aggregated)."""
    ids, mask, age, gender, eth, ins, lab, text = batch8
    d = behrt_demo(sd, ids, mask, age, gender, eth, ins, prefix="BEHRT.")
    l = behrt_lab(sd, lab)
    parts = []
    for m, e in (("demo", d), ("lab", l), ("text", text)):
        pr = torch.relu(F.linear(e, sd[f"{m}_projector.0.weight"].float(), sd[f"{m}_projector.0.bias"].float()))
        parts.append(pr * torch.sigmoid(sd[f"sig_weights_{m}"].float()))
    agg = torch.relu(F.linear(torch.cat(parts, dim=1), sd["aggregate_projector.0.weight"].float(),
                              sd["aggregate_projector.0.bias"].float()))
    hid = torch.relu(F.linear(agg, sd["classifier.0.weight"].float(), sd["classifier.0.bias"].float()))
    return F.linear(hid, sd["classifier.3.weight"].float(), sd["classifier.3.bias"].float()), agg


def average_fusion_forward(sd, ids, mask, codes7, text):
    """MultimodalTransformer.forward of 07_multimodal_average_fusion.py:221-238 with its BEHRTModel (156-203), eval
    mode.  codes7 = (age, segment, admission_loc, discharge_loc, gender, ethnicity, insurance) int64 [B] each.
    Returns (logits [B,3], fused_embedding_pre_relu [B,512])."""
    names = ("age", "segment", "admission_loc", "discharge_loc", "gender", "ethnicity", "insurance")
    cls = bert_encode(sd, "BEHRT.bert.", ids, mask)[:, 0, :]
    extra = 0
    for n, c in zip(names, codes7):
        tab = sd[f"BEHRT.{n}_embedding.weight"].float()
        extra = extra + tab[c.clamp(0, tab.shape[0] - 1)]
    emb = cls + extra / 7.0
    ts_pre = F.linear(emb, sd["ts_linear.weight"].float(), sd["ts_linear.bias"].float())
    tx_pre = F.linear(text, sd["text_linear.weight"].float(), sd["text_linear.bias"].float())
    comb = torch.cat([torch.relu(ts_pre), torch.relu(tx_pre)], dim=1)
    hid = torch.relu(F.linear(comb, sd["classifier.0.weight"].float(), sd["classifier.0.bias"].float()))
    return F.linear(hid, sd["classifier.3.weight"].float(), sd["classifier.3.bias"].float()), torch.cat([ts_pre, tx_pre], dim=1)


def eddi_08(y_true, y_prob, sensitive, threshold=0.5):
    """compute_eddi of 08_multimodal_eddi_fusion.py:45-59 (np.unique groups; NaN-free here; sum, not nansum)."""
    pred = (y_prob > threshold).astype(int)
    err = np.mean(pred != y_true)
    denom = max(err, 1 - err) if err not in [0, 1] else 1.0
    groups = np.unique(sensitive)
    sub = [(np.mean(pred[sensitive == g] != y_true[sensitive == g]) - err) / denom for g in groups]
    return float(np.sqrt(np.sum(np.array(sub) ** 2)) / len(groups))


def eddi_fusion_forward(sd, batch8, labels=None, sensitive=None, beta=0.3, old_weights=None):
    """MultimodalTransformer.forward of 08_multimodal_eddi_fusion.py:348-449: per outcome, the three modality logits are
    fused with weights  w_m = (old_w_m | 0.33) + beta * (max_m' EDDI_m' - EDDI_m), where EDDI_m is computed INSIDE the
    forward from the batch's thresholded modality predictions over `sensitive` (the weights are constants for autograd).
    labels f32 [B,3] / sensitive int [B] may be None (EDDI = 0).  Returns (logits [B,3], weights [3][3], eddi [3][3])."""
    ids, mask, age, gender, eth, ins, lab, text = batch8
    d = behrt_demo(sd, ids, mask, age, gender, eth, ins, num_layers=6, num_heads=6)       # 08:264-265
    l = behrt_lab(sd, lab)
    proj = {}
    for m, e in (("demo", d), ("lab", l), ("text", text)):
        proj[m] = torch.relu(F.linear(e, sd[f"{m}_projector.0.weight"].float(), sd[f"{m}_projector.0.bias"].float()))
    logits, weights, eddis = [], [], []
    for oi, o in enumerate(("mort", "los", "mv")):
        raw = {m: F.linear(proj[m], sd[f"classifier_{m}_{o}.weight"].float(), sd[f"classifier_{m}_{o}.bias"].float())
               for m in ("demo", "lab", "text")}
        if labels is not None and sensitive is not None:
            y = np.asarray(labels[:, oi])
            e = [eddi_08(y, torch.sigmoid(raw[m].detach()).numpy().squeeze(), np.asarray(sensitive)) for m in ("demo", "lab", "text")]
        else:
            e = [0.0, 0.0, 0.0]
        top = max(e)
        base = old_weights[oi] if old_weights is not None else (0.33, 0.33, 0.33)
        w = [base[k] + beta * (top - e[k]) for k in range(3)]
        logits.append(raw["demo"] * w[0] + raw["lab"] * w[1] + raw["text"] * w[2])
        weights.append(w)
        eddis.append(e)
    return torch.cat(logits, dim=1), weights, eddis


def eddi_fusion_loss(logits, labels, pos_weight, loss_gamma=1.0, target=1.0, gamma=1.0):
    """train_step objective of 08:475-479: three FocalLoss(gamma=1, pos_weight_i) + loss_gamma * mean((mort - target)^2)."""
    return text_classifier_loss(logits, labels, pos_weight, gamma) + loss_gamma * ((logits[:, 0:1] - target) ** 2).mean()


def text_classifier(sd, x):
    """UnstructuredClassifier.forward (02_BioClinicalBERT.py:122-134), eval mode: logits f32 [B, 3]."""
    h = torch.relu(F.linear(x, sd["classifier.0.weight"].float(), sd["classifier.0.bias"].float()))
    return F.linear(h, sd["classifier.3.weight"].float(), sd["classifier.3.bias"].float())


def focal_loss(logits, targets, pos_weight, gamma=2.0, alpha=None):
    """FocalLoss.forward, reduction='mean' (02_BioClinicalBERT.py:26-38)."""
    bce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none", pos_weight=pos_weight)
    fl = (1 - torch.exp(-bce)) ** gamma * bce
    return (fl if alpha is None else alpha * fl).mean()


def text_classifier_loss(logits, labels, pos_weight, gamma=2.0):
    """train_model's objective (02_BioClinicalBERT.py:143-146): the sum of the three focal losses."""
    return sum(focal_loss(logits[:, i:i + 1], labels[:, i:i + 1], pos_weight[i], gamma) for i in range(3))


def eo_difference_n2(tpr, fpr):
    """calculate_equalized_odds_difference (01_BEHRT.py:27-42): sum over i < j of |d| divided by n^2."""
    g = list(tpr.keys())
    n = len(g)
    if n == 0:
        return 0.0, 0.0, 0.0
    t = sum(abs(tpr[g[i]] - tpr[g[j]]) for i in range(n) for j in range(i + 1, n)) / n ** 2
    f = sum(abs(fpr[g[i]] - fpr[g[j]]) for i in range(n) for j in range(i + 1, n)) / n ** 2
    return t, f, (t + f) / 2.0


def eddi_unique_groups(sensitive, y_true, score, threshold=0.5):
    """compute_eddi of 01_BEHRT.py:85-100 (groups = np.unique, denominator 1.0 when the error rate is 0 or 1)."""
    pred = (score > threshold).astype(int)
    err = np.mean(pred != y_true)
    denom = max(err, 1 - err) if err not in [0, 1] else 1.0
    sub = {}
    for gval in np.unique(sensitive):
        m = sensitive == gval
        sub[gval] = (np.mean(pred[m] != y_true[m]) - err) / denom
    return float(np.sqrt(np.nansum(np.array(list(sub.values())) ** 2)) / len(sub)), sub


def fusion(sd, demo_emb, lab_emb, text_emb, weights=(0.33, 0.33, 0.33)):
    """MultimodalTransformer_EDDI_Sigmoid.forward after the encoders (FAME:276-308), eval mode.
    weights = (w_demo, w_lab, w_text): the 'mortality' entry of old_eddi_weights (FAME:283-285) or 0.33."""
    g = lambda k: sd[k].float()
    proj = lambda e, n: torch.relu(F.linear(e, g(f"{n}_projector.0.weight"), g(f"{n}_projector.0.bias")))
    pd_, pl, pt = proj(demo_emb, "demo"), proj(lab_emb, "lab"), proj(text_emb, "text")
    fused = torch.cat([weights[0] * pd_, weights[1] * pl, weights[2] * pt], dim=1)
    sig = torch.sigmoid(g("sig_weights"))
    gated = fused * sig
    pre = F.linear(gated, g("fusion_mlp.0.weight"), g("fusion_mlp.0.bias"))
    logits = F.linear(torch.relu(pre), g("fusion_mlp.3.weight"), g("fusion_mlp.3.bias"))
    mod = {n: F.linear(x, g(f"classifier_{n}.weight"), g(f"classifier_{n}.bias"))
           for n, x in (("demo", pd_), ("lab", pl), ("text", pt))}
    return {"fused_logits": logits, "sigmoid_weights": sig, "gated_vector": gated, "fusion_pre_relu": pre,
            "modality_logits": mod, "proj": {"demo": pd_, "lab": pl, "text": pt}}


def fame_forward(sd, batch, weights=(0.33, 0.33, 0.33)):
    """Whole model forward, eval mode.  batch = the 9 tensors of SURVEY.md appendix C (labels unused)."""
    ids, mask, age, gender, eth, ins, lab, text = batch[:8]
    d = behrt_demo(sd, ids, mask, age, gender, eth, ins)
    l = behrt_lab(sd, lab)
    out = fusion(sd, d, l, text, weights)
    out["demo_embedding"], out["lab_embedding"] = d, l
    return out


# ------------------------------------------------------------------------------------------------ loss
def fame_loss(logits, labels, attrs, sig_weights, pos_weight, lambda_edd, lambda_l1):
    """train_step's objective (FAME:420-444) on torch tensors (differentiable):
    BCEWithLogits(pos_weight) mean + lambda_edd * 10 * LEDDI + lambda_l1 * |sig_weights|_1.
    attrs = (age_ids, ethnicity_ids, insurance_ids).  Returns (total, bce, leddi)."""
    z, y = logits, labels
    logsig = F.logsigmoid(z)
    bce = (-(pos_weight * y * logsig + (1 - y) * (logsig - z))).mean()
    err = (torch.sigmoid(z) - y).abs()
    terms = []
    for i in range(z.shape[1]):
        e = err[:, i]
        overall = e.mean()
        for a in attrs:
            devs = [(e[a == gval].mean() - overall) ** 2 for gval in torch.unique(a).tolist()]
            terms.append(torch.sqrt(torch.stack(devs).mean() + 1e-8))
    leddi = torch.stack(terms).mean()
    total = bce + lambda_edd * (10.0 * leddi) + lambda_l1 * sig_weights.abs().sum()
    return total, bce, leddi


def loss_group_stats(logits, labels, attrs, n_slots=8):
    """The integer/fp statistics behind LEDDI: per (outcome, attr, group-code) count and sum |sigmoid(z)-y|.
    Returns (counts int64 [3 attrs, n_slots], sums float64 [n_out, 3, n_slots])."""
    err = (torch.sigmoid(logits.double()) - labels.double()).abs().numpy()
    counts = np.zeros((len(attrs), n_slots), dtype=np.int64)
    sums = np.zeros((err.shape[1], len(attrs), n_slots), dtype=np.float64)
    for ai, a in enumerate(attrs):
        a = np.asarray(a)
        for gval in range(n_slots):
            m = a == gval
            counts[ai, gval] = int(m.sum())
            sums[:, ai, gval] = err[m].sum(axis=0)
    return counts, sums


# ------------------------------------------------------------------------------------------------ metrics
def compute_eddi(y_true, y_score, groups_of, threshold=0.5, complete_groups=None):
    """FAME:54-82.  Error rate disparity: d_g = (err_g - err) / max(err, 1 - err) over non-empty groups,
    EDDI = ||d||_2 / #non-empty groups."""
    y_true = np.asarray(y_true)
    wrong = (np.asarray(y_score) > threshold).astype(int) != y_true
    groups_of = np.asarray(groups_of)
    cand = np.unique(groups_of) if complete_groups is None else np.asarray(complete_groups)
    err = wrong.mean()
    den = 1.0 - err if err < 0.5 else err
    per = {}
    for gval in cand:
        sel = groups_of == gval
        if sel.any():
            per[gval] = (wrong[sel].mean() - err) / den
    if not per:
        return 0.0, per
    return float(np.sqrt(np.sum(np.square(list(per.values())))) / len(per)), per


def combined_eddi(y_true, y_score, age, eth, ins, threshold):
    """sqrt(EDDI_age^2 + EDDI_eth^2 + EDDI_ins^2) / 3 over the fixed code lists (FAME:359-364, 897-901)."""
    parts = [compute_eddi(y_true, y_score, a, threshold, gl)[0]
             for a, gl in ((age, AGE_GROUPS), (eth, ETH_GROUPS), (ins, INS_GROUPS))]
    return float(np.sqrt(sum(p * p for p in parts)) / 3.0), parts


def group_rates(y_true, y_pred, sel):
    """TPR / FPR inside a group, 0 when the denominator is empty (FAME:84-97)."""
    yt, yp = np.asarray(y_true)[sel], np.asarray(y_pred)[sel]
    tp = int(((yt == 1) & (yp == 1)).sum()); fn = int(((yt == 1) & (yp == 0)).sum())
    fp = int(((yt == 0) & (yp == 1)).sum()); tn = int(((yt == 0) & (yp == 0)).sum())
    return (tp / (tp + fn) if tp + fn else 0), (fp / (fp + tn) if fp + tn else 0), (tp, fn, fp, tn)


def equalized_odds(y_true, y_pred, groups_of):
    """Mean pairwise |dTPR|, |dFPR| over the groups present, EO = their average (FAME:99-122)."""
    groups_of = np.asarray(groups_of)
    rates = [group_rates(y_true, y_pred, groups_of == gval)[:2] for gval in np.unique(groups_of)]
    n = len(rates)
    dt = [abs(rates[i][0] - rates[j][0]) for i in range(n) for j in range(i + 1, n)]
    df = [abs(rates[i][1] - rates[j][1]) for i in range(n) for j in range(i + 1, n)]
    a, b = (float(np.mean(dt)) if dt else 0.0), (float(np.mean(df)) if df else 0.0)
    return a, b, (a + b) / 2.0


def f1_from_counts(tp, fp, fn):
    return 0.0 if 2 * tp + fp + fn == 0 else 2.0 * tp / (2 * tp + fp + fn)


def calibrate_thresholds(logits, labels):
    """FAME:468-482: per outcome the first threshold of linspace(0,1,101) with strictly larger F1, from (0.5, 0)."""
    out = {}
    for i, name in enumerate(OUTCOMES):
        probs = torch.sigmoid(torch.as_tensor(logits)[:, i]).numpy()
        y = np.asarray(labels)[:, i]
        best_t, best_f = 0.5, 0.0
        for t in np.linspace(0, 1, 101):
            pred = probs > t
            tp = int((pred & (y == 1)).sum()); fp = int((pred & (y == 0)).sum()); fn = int((~pred & (y == 1)).sum())
            f = f1_from_counts(tp, fp, fn)
            if f > best_f:
                best_f, best_t = f, t
        out[name] = best_t
    return out


def auroc(y_true, score):
    """sklearn.metrics.roc_auc_score for binary labels: trapezoid over distinct thresholds (ties share a point)."""
    y = np.asarray(y_true).astype(np.float64)
    s = np.asarray(score)
    if y.min() == y.max():
        return float("nan")
    order = np.argsort(-s, kind="mergesort")
    s, y = s[order], y[order]
    last = np.r_[np.nonzero(np.diff(s))[0], len(s) - 1]
    tps = np.cumsum(y)[last]
    fps = (1 + last) - tps
    tps, fps = np.r_[0, tps], np.r_[0, fps]
    return float(np.trapezoid(tps / tps[-1], fps / fps[-1]))


def auprc(y_true, score):
    """sklearn.metrics.average_precision_score: sum_n (R_n - R_{n-1}) P_n over distinct thresholds."""
    y = np.asarray(y_true).astype(np.float64)
    s = np.asarray(score)
    if y.sum() == 0:
        return 0.0
    order = np.argsort(-s, kind="mergesort")
    s, y = s[order], y[order]
    last = np.r_[np.nonzero(np.diff(s))[0], len(s) - 1]
    tps = np.cumsum(y)[last]
    fps = (1 + last) - tps
    prec = tps / (tps + fps)
    rec = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0, rec]) * prec))


def evaluate(logits, labels, age, eth, ins, thresholds):
    """evaluate_model_multi (FAME:511-552) + the EDDI tail of run_experiment (FAME:887-915), on gathered arrays."""
    logits = torch.as_tensor(logits)
    labels = np.asarray(labels)
    metrics, fair, eddi = {}, {}, {}
    for i, name in enumerate(OUTCOMES):
        th = thresholds[name] if isinstance(thresholds, dict) else thresholds
        probs = torch.sigmoid(logits[:, i]).numpy()
        y = labels[:, i]
        pred = (probs > th).astype(int)
        tp = int(((pred == 1) & (y == 1)).sum()); fp = int(((pred == 1) & (y == 0)).sum())
        fn = int(((pred == 0) & (y == 1)).sum()); tn = int(((pred == 0) & (y == 0)).sum())
        metrics[name] = {
            "aucroc": auroc(y, probs), "auprc": auprc(y, probs), "f1": f1_from_counts(tp, fp, fn),
            "recall (TPR)": tp / (tp + fn) if tp + fn else 0.0, "TPR": tp / (tp + fn) if tp + fn else 0,
            "precision": tp / (tp + fp) if tp + fp else 0.0, "fpr": fp / (fp + tn) if fp + tn else 0,
            "optimal_threshold": th, "counts": (tp, fn, fp, tn),
        }
        fair[name] = {}
        eos = []
        for an, av in (("age", age), ("ethnicity", eth), ("insurance", ins)):
            a, b, eo = equalized_odds(y, pred, av)
            fair[name][an] = {"avg_tpr_diff": a, "avg_fpr_diff": b, "eo_metric": eo}
            eos.append(eo)
        fair[name]["overall_eo"] = float(np.mean(eos))
        comb, parts = combined_eddi(y, probs, age, eth, ins, th)
        eddi[name] = {"age": parts[0], "ethnicity": parts[1], "insurance": parts[2], "combined": comb}
    eddi["overall"] = float(np.mean([eddi[n]["combined"] for n in OUTCOMES]))
    return metrics, fair, eddi


def update_dynamic_weights(mod_preds, labels, age, eth, ins, old_weights, beta, threshold=0.5):
    """update_dynamic_weights_all_tasks after the forward passes (FAME:346-399).
    mod_preds[outcome][modality]: 0/1 float arrays ((sigmoid(logit) > threshold), FAME:335-337)."""
    new = {}
    for oi, name in enumerate(OUTCOMES):
        y = np.asarray(labels)[:, oi]
        e = {m: combined_eddi(y, mod_preds[name][m], age, eth, ins, threshold)[0] for m in MODALITIES}
        top = max(e.values())
        prev = old_weights.get(name, {m: 0.33 for m in MODALITIES})
        raw = {m: max(prev[m] + float(np.clip(beta * (top - e[m]), -0.05, 0.05)), 0.1) for m in MODALITIES}
        tot = sum(raw.values())
        new[name] = {m: raw[m] / tot for m in MODALITIES}
    return new


# ------------------------------------------------------------------------------------------------ optimizer
def clip_and_adamw(params, grads, m, v, step, lr, wd, max_norm=1.0, b1=0.9, b2=0.999, eps=1e-8):
    """clip_grad_norm_(max_norm) + one torch.optim.AdamW step (FAME:446-447) on lists of float32 tensors.
    Entries whose grad is None are skipped entirely (no decay), as torch does."""
    live = [i for i, g in enumerate(grads) if g is not None]
    total = torch.sqrt(sum((grads[i].double() ** 2).sum() for i in live)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for i in live:
        g = grads[i] * coef
        params[i].mul_(1 - lr * wd)
        m[i].mul_(b1).add_(g, alpha=1 - b1)
        v[i].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = (v[i].sqrt() / math.sqrt(bc2)).add_(eps)
        params[i].addcdiv_(m[i], denom, value=-lr / bc1)
    return float(total)

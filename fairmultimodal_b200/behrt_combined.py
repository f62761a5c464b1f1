"""Structured-only baseline of the reference (FinalCode/New/Final/01_BEHRT.py) on the B200 kernels: SURVEY.md 8(f-2).

    BEHRTModel_Combined                     01_BEHRT.py:112-131   lab tower (= BEHRTModel_Lab of 10_FAME.py) + fusion_fc
                                                                  768 -> 768 + Dropout(0.1) + three Linear(768, 1) heads
    train_epoch / optimisation_step         01_BEHRT.py:215-233   sum of three BCEWithLogitsLoss(pos_weight_i) ->
                                                                  backward -> clip_grad_norm_(1.0) -> AdamW
    calculate_equalized_odds_difference     01_BEHRT.py:27-42     EO variant  sum_{i<j} |d| / n^2
    compute_eddi                            01_BEHRT.py:85-100    EDDI variant over np.unique groups (denominator 1.0 when
                                                                  the overall error is exactly 0 or 1)
    compute_attribute_eddi                  01_BEHRT.py:102-103

The lab tower, its hand-written backward, the loss kernels, clip + AdamW and the dropout machinery are the ones of
train.py; only the 768 -> 768 -> 3 head (fp32, a few small products) is specific to this model.  Same class name,
constructor, forward signature (three [B, 1] logit tensors) and state_dict keys as the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import ops_train as T
from . import train
from .modules import BEHRTModel_Lab

HEADS = ("classifier_mort", "classifier_los", "classifier_mech")


class BEHRTModel_Combined(nn.Module):
    def __init__(self, lab_token_count, hidden_size=768):
        super().__init__()
        if hidden_size != 768:
            raise ValueError("the B200 lab-tower kernels are built for hidden_size 768 (8 heads of 96)")
        self.lab_model = BEHRTModel_Lab(lab_token_count, hidden_size, nhead=8, num_layers=2)
        self.fusion_fc = nn.Linear(hidden_size, hidden_size)
        self.dropout = nn.Dropout(0.1)
        self.classifier_mort = nn.Linear(hidden_size, 1)
        self.classifier_los = nn.Linear(hidden_size, 1)
        self.classifier_mech = nn.Linear(hidden_size, 1)

    def _head(self, emb):
        """logits f32 [B, 3] from the lab embedding (eval: no dropout)."""
        with torch.no_grad():
            wf, bf = self.fusion_fc.weight.detach().float(), self.fusion_fc.bias.detach().float()
            wc = torch.cat([getattr(self, h).weight.detach().float() for h in HEADS]).contiguous()
            bc = torch.cat([getattr(self, h).bias.detach().float() for h in HEADS]).contiguous()
            return _head_forward(emb, wf, bf, wc, bc, None)[1]

    def forward(self, lab_features):
        if not lab_features.is_cuda:
            raise RuntimeError("BEHRTModel_Combined runs on a B200 only: move inputs to cuda (no CPU fallback)")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # outputs of the training-mode forward (dropout active); gradients come from forward_backward(), not autograd
            st = get_state(self)
            with torch.no_grad():
                ds = _drop_sites(self, st)
                emb, _ = train._lab_forward(st, self, lab_features, ds, lab_module=self.lab_model, pre="lab_model.")
                logits = _head_forward(emb, st.f("fusion_fc.weight"), st.f("fusion_fc.bias"), _wc(st), _bc(st),
                                       ds.site("combined.head", ds.p_fusion) if ds is not None else None)[1]
        else:
            with torch.no_grad():
                logits = self._head(self.lab_model(lab_features))
        return logits[:, 0:1], logits[:, 1:2], logits[:, 2:3]


def _head_forward(emb, wf, bf, wc, bc, drop):
    """fused = dropout(emb Wf^T + bf) [B,768];  logits = fused Wc^T + bc [B,3]   (01_BEHRT.py:124-130)."""
    B = emb.shape[0]
    emb = emb.float().contiguous()
    fused = bf.repeat(B, 1)
    T.sgemm(emb, 768, 1, wf, 1, 768, fused, B, 768, 768, accumulate=True)
    T.dropout_apply(fused, drop)
    logits = bc.repeat(B, 1)
    T.sgemm(fused, 768, 1, wc, 1, 768, logits, B, 3, 768, accumulate=True)
    return fused, logits


def _wc(st):
    return torch.cat([st.f(h + ".weight") for h in HEADS]).contiguous()        # [3, 768]


def _bc(st):
    return torch.cat([st.f(h + ".bias") for h in HEADS]).contiguous()          # [3]


def get_state(model) -> train.FlatTrainState:
    st = getattr(model, "_fame_train_state", None)
    if st is None or st.model is not model:
        st = train.FlatTrainState(model, no_grad_prefixes=(), fame_layout=False)
        object.__setattr__(model, "_fame_train_state", st)
    return st


def _drop_sites(model, st):
    ds = train.DropSites(model, st.step_dev, lab_module=model.lab_model, head_dropout=model.dropout)
    return ds if ds.any else None


def forward_backward(model, lab, labels, pos_weight, group=None):
    """Forward + loss + backward of one batch (01_BEHRT.py:217-229); gradients land in the flat buffer.
    loss = sum_i BCEWithLogits(pos_weight_i)(logits_i, labels_i), each a mean over the (global) batch.
    Returns (loss f32 [1] on the device, logits [B,3])."""
    st = get_state(model)
    post = st.post_stream()
    post.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(post):                       # beside the forward pass (see train.forward_backward)
        st.zero_grad()
        st.sumsq.zero_()
    ds = _drop_sites(model, st)
    emb, saved = train._lab_forward(st, model, lab, ds, lab_module=model.lab_model, pre="lab_model.")
    wf, wc = st.f("fusion_fc.weight"), _wc(st)
    d_head = ds.site("combined.head", ds.p_fusion) if ds is not None else None
    fused, logits = _head_forward(emb, wf, st.f("fusion_fc.bias"), wc, _bc(st), d_head)
    B = logits.shape[0]
    dev = logits.device
    labels = labels.float().contiguous()
    # the BCE kernels of the FAME loss: mean over B * 3 entries, so the sum of the three per-outcome means is 3 x that;
    # the LEDDI / L1 terms are switched off (lambda = 0) and every patient sits in subgroup 0
    zeros = torch.zeros(B, device=dev, dtype=torch.int64)
    stats = ops.loss_stats(logits, labels, [zeros, zeros, zeros], pos_weight)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(stats, group=group)
    loss4, dlogits = ops.loss_fwd_bwd(logits, labels, [zeros, zeros, zeros], pos_weight, stats, None, 0.0, 0.0)
    loss = loss4[1:2] * 3.0
    dlogits = dlogits * 3.0
    red = train._GradReducer(st, group)
    torch.cuda.current_stream().wait_stream(post)
    # head backward (fp32): dWc = dlogits^T fused, dbc = colsum(dlogits); dfused = dlogits Wc (through the dropout);
    # dWf = dfused^T emb, dbf = colsum(dfused); demb = dfused Wf
    dwc = torch.empty((3, 768), device=dev, dtype=torch.float32)
    T.sgemm(dlogits, 1, 3, fused, 768, 1, dwc, 3, 768, B)
    dbc = torch.zeros(3, device=dev, dtype=torch.float32)
    T.colsum(dlogits, dbc)
    for i, h in enumerate(HEADS):
        st.gr(h + ".weight").copy_(dwc[i:i + 1])
        st.gr(h + ".bias").copy_(dbc[i:i + 1])
    dfused = torch.empty((B, 768), device=dev, dtype=torch.float32)
    T.sgemm(dlogits, 3, 1, wc, 768, 1, dfused, B, 768, 3)
    T.dropout_apply(dfused, d_head)
    emb32 = emb.float().contiguous()
    T.sgemm(dfused, 1, 768, emb32, 768, 1, st.gr("fusion_fc.weight"), 768, 768, B)
    T.colsum(dfused, st.gr("fusion_fc.bias"))
    demb = torch.empty((B, 768), device=dev, dtype=torch.float32)
    T.sgemm(dfused, 768, 1, wf, 768, 1, demb, B, 768, 768)
    train._lab_backward(st, model, saved, demb, ds, lab_module=model.lab_model, pre="lab_model.")
    red.ready("tail")
    red.finish()
    return loss, logits


def optimisation_step(model, lab, labels, pos_weight, hp, group=None):
    """One batch of the reference's training loop (01_BEHRT.py:217-232): forward, summed BCE, backward, clip, AdamW."""
    st = get_state(model)
    loss, _ = forward_backward(model, lab, labels, pos_weight, group=group)
    st.clip_and_step(hp["lr"], hp["weight_decay"], hp.get("betas", (0.9, 0.999)), hp.get("eps", 1e-8), max_norm=1.0)
    return loss


def train_epoch(model, train_loader, optimizer, device, pos_weight, group=None):
    """The training half of one epoch of train_model (01_BEHRT.py:214-233): returns the mean batch loss.  Batches are
    the reference's 6-tuples (lab_features, age, gender, ethnicity, insurance, labels)."""
    model.train()
    g = optimizer.param_groups[0]
    hp = dict(lr=g["lr"], weight_decay=g.get("weight_decay", 0.01), betas=tuple(g.get("betas", (0.9, 0.999))),
              eps=g.get("eps", 1e-8))
    pw = torch.as_tensor(pos_weight, dtype=torch.float32, device=device)
    losses = []
    for batch in train_loader:
        lab, labels = batch[0].to(device, non_blocking=True), batch[-1].to(device, non_blocking=True)
        losses.append(optimisation_step(model, lab, labels, pw, hp, group=group))
    return float(torch.cat(losses).mean().item()) if losses else float("inf")


# ------------------------------------------------------------------------------------------------ metric variants
def calculate_equalized_odds_difference(tpr_dict, fpr_dict):
    """01_BEHRT.py:27-42: pairwise absolute differences summed over i < j and divided by n^2 (not by the pair count)."""
    groups = list(tpr_dict.keys())
    n = len(groups)
    if n == 0:
        return {"EOTPR": 0.0, "EOFPR": 0.0, "EO": 0.0}
    t = f = 0.0
    for i in range(n):
        for j in range(i + 1, n):
            t += abs(tpr_dict[groups[i]] - tpr_dict[groups[j]])
            f += abs(fpr_dict[groups[i]] - fpr_dict[groups[j]])
    return {"EOTPR": t / n ** 2, "EOFPR": f / n ** 2, "EO": (t / n ** 2 + f / n ** 2) / 2.0}


def group_counts(sensitive_codes, true_labels, scores, threshold=0.5, scores_are_probs=True, device="cuda"):
    """Integer confusion counts [group code 0..7][TP, FN, FP, TN] of ONE outcome from the evaluation count kernel
    (predictions = scores > threshold, strict, as 01_BEHRT.py:86)."""
    n = len(true_labels)
    z = torch.zeros((n, 3), device=device, dtype=torch.float32)
    z[:, 0] = torch.as_tensor(np.asarray(scores, dtype=np.float32).reshape(-1), device=device)
    y = torch.zeros((n, 3), device=device, dtype=torch.float32)
    y[:, 0] = torch.as_tensor(np.asarray(true_labels, dtype=np.float32).reshape(-1), device=device)
    a = torch.as_tensor(np.asarray(sensitive_codes, dtype=np.int64).reshape(-1), device=device)
    zero = torch.zeros_like(a)
    from .metrics import Counts
    c = Counts(ops.eval_counts(z, y, [a, zero, zero], (threshold, 0.5, 0.5), logits_are_probs=scores_are_probs))
    return np.asarray(c.conf[0, 0], dtype=np.int64)


def compute_eddi(sensitive_codes, true_labels, pred_scores, threshold=0.5, device="cuda"):
    """01_BEHRT.py:85-100 on integer group codes 0..7: overall EDDI = sqrt(sum_g d_g^2) / #groups with
    d_g = (err_g - err) / max(err, 1 - err) (denominator 1.0 if err is exactly 0 or 1); groups = codes present."""
    conf = group_counts(sensitive_codes, true_labels, pred_scores, threshold, True, device)
    n_g = conf.sum(axis=1)
    wrong = conf[:, 1] + conf[:, 2]                                   # FN + FP
    total = int(n_g.sum())
    overall = wrong.sum() / total
    denom = max(overall, 1 - overall) if overall not in (0, 1) else 1.0
    sub = {int(g): (wrong[g] / n_g[g] - overall) / denom for g in range(conf.shape[0]) if n_g[g] > 0}
    vals = np.array(list(sub.values()), dtype=np.float64)
    return float(np.sqrt(np.sum(vals ** 2)) / len(sub)), sub


def compute_attribute_eddi(age_eddi, ethnicity_eddi, insurance_eddi):
    return float(np.sqrt(age_eddi ** 2 + ethnicity_eddi ** 2 + insurance_eddi ** 2) / 3.0)


def group_tpr_fpr(sensitive_codes, true_labels, pred_labels, device="cuda"):
    """Per-group (TPR, FPR) dicts as built around calculate_tpr_and_fpr (01_BEHRT.py:17-25); 0.0 on empty denominators."""
    conf = group_counts(sensitive_codes, true_labels, pred_labels, 0.5, True, device)
    tpr, fpr = {}, {}
    for g in range(conf.shape[0]):
        tp, fn, fp, tn = (int(x) for x in conf[g])
        if tp + fn + fp + tn == 0:
            continue
        tpr[g] = tp / (tp + fn) if tp + fn > 0 else 0.0
        fpr[g] = fp / (fp + tn) if fp + tn > 0 else 0.0
    return tpr, fpr

"""Text-only baseline of the reference (FinalCode/New/Final/02_BioClinicalBERT.py) on the B200 kernels: SURVEY.md 8(f-1).

    UnstructuredClassifier   02_BioClinicalBERT.py:122-134   Linear(768, 256), ReLU, Dropout(0.1), Linear(256, 3) over the
                                                             per-patient note embedding (the a1 + a2 pipeline)
    FocalLoss                02_BioClinicalBERT.py:18-38     (1 - exp(-bce))^gamma * bce with pos_weight, mean
    train_model              02_BioClinicalBERT.py:137-152   three summed focal losses -> backward -> AdamW (no clipping)

The note embeddings come from modules.apply_bioclinicalbert_on_patient_notes (same function as in 10_FAME.py; this
script's variant also returns the patient ids).  Same class names, constructor arguments, forward signature and
state_dict keys (classifier.0.*, classifier.3.*) as the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import modules
from . import ops_train as T
from . import train


class FocalLoss(nn.Module):
    """Holder of (gamma, alpha, reduction='mean', pos_weight) as in the reference; the arithmetic runs in
    fame_focal_loss_fwd_bwd.  Calling it on CUDA logits [B, 1] / targets [B, 1] returns the loss value (no autograd)."""

    def __init__(self, gamma=2, alpha=None, reduction="mean", pos_weight=None):
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("the reference trains with reduction='mean' (02_BioClinicalBERT.py:495-497)")
        self.gamma, self.alpha, self.reduction, self.pos_weight = gamma, alpha, reduction, pos_weight

    def forward(self, logits, targets):
        B = logits.shape[0]
        z = torch.zeros((B, 3), device=logits.device, dtype=torch.float32)
        y = torch.zeros((B, 3), device=logits.device, dtype=torch.float32)
        z[:, 0], y[:, 0] = logits.reshape(-1).float(), targets.reshape(-1).float()
        pw = torch.ones(3, device=logits.device)
        if self.pos_weight is not None:
            pw[0] = torch.as_tensor(self.pos_weight, dtype=torch.float32, device=logits.device).reshape(-1)[0]
        full, _ = T.focal_loss_fwd_bwd(z, y, pw, self.gamma, 1.0 if self.alpha is None else self.alpha, want_grad=False)
        # columns 1 and 2 hold z = 0, y = 0: bce = log 2, pt = 1/2 -> a known constant each
        pad = (1.0 if self.alpha is None else self.alpha) * (0.5 ** self.gamma) * np.log(2.0)
        return (full - 2.0 * pad).float().squeeze(0)


class UnstructuredClassifier(nn.Module):
    def __init__(self, input_size=768, hidden_size=256):
        super().__init__()
        self.classifier = nn.Sequential(nn.Linear(input_size, hidden_size), nn.ReLU(), nn.Dropout(0.1),
                                        nn.Linear(hidden_size, 3))

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("UnstructuredClassifier runs on a B200 only: move inputs to cuda (no CPU fallback)")
        with torch.no_grad():
            c = self.classifier
            drop = None
            if self.training and c[2].p > 0:
                drop = _site(self, get_state(self))
            return _mlp_forward(x, c[0].weight.detach().float(), c[0].bias.detach().float(),
                                c[3].weight.detach().float(), c[3].bias.detach().float(), drop)[2]


def _site(model, st):
    """DropoutCfg of the classifier's only dropout (None in eval mode / p = 0)."""
    from . import _lib
    p = float(model.classifier[2].p) if model.training else 0.0
    if p <= 0.0:
        return None
    c = getattr(st, "_drop_cfg", None)
    if c is None or c.thresh16 != max(1, min(65535, int(round(p * 65536.0)))):
        c = _lib.DropoutCfg()
        c.step, c.seed, c.group_shift = st.step_dev.data_ptr(), 0x7E57C1A5, 0
        c.thresh16 = max(1, min(65535, int(round(p * 65536.0))))
        st._drop_cfg = c
    return c


def _mlp_forward(x, w1, b1, w2, b2, drop):
    """pre = x W1^T + b1; h = dropout(relu(pre)); logits = h W2^T + b2.  Returns (pre, h, logits), fp32."""
    B, K = x.shape
    Hd = w1.shape[0]
    x = x.float().contiguous()
    pre = b1.repeat(B, 1)
    T.sgemm(x, K, 1, w1, 1, K, pre, B, Hd, K, accumulate=True)
    h = T.relu_(pre.clone())
    T.dropout_apply(h, drop)
    logits = b2.repeat(B, 1)
    T.sgemm(h, Hd, 1, w2, 1, Hd, logits, B, 3, Hd, accumulate=True)
    return pre, h, logits


def get_state(model) -> train.FlatTrainState:
    st = getattr(model, "_fame_train_state", None)
    if st is None or st.model is not model:
        st = train.FlatTrainState(model, no_grad_prefixes=(), fame_layout=False)
        object.__setattr__(model, "_fame_train_state", st)
    return st


def forward_backward(model, x, labels, pos_weight, gamma=2.0, alpha=None):
    """One batch of train_model (02_BioClinicalBERT.py:141-148): loss = sum of the three focal losses; gradients land
    in the flat buffer.  labels f32 [B, 3] (mortality, los, mech); pos_weight f32 [3].  Returns (loss f64 [1], logits)."""
    st = get_state(model)
    st.zero_grad()
    f, g = st.f, st.gr
    w1, w2 = f("classifier.0.weight"), f("classifier.3.weight")
    drop = _site(model, st)
    pre, h, logits = _mlp_forward(x, w1, f("classifier.0.bias"), w2, f("classifier.3.bias"), drop)
    B, K = x.shape
    Hd = w1.shape[0]
    loss, dlogits = T.focal_loss_fwd_bwd(logits, labels, pos_weight, gamma, 1.0 if alpha is None else alpha)
    # dW2 = dlogits^T h, db2; dh = dlogits W2 (through the dropout and the ReLU); dW1 = dh^T x, db1
    T.sgemm(dlogits, 1, 3, h, Hd, 1, g("classifier.3.weight"), 3, Hd, B)
    T.colsum(dlogits, g("classifier.3.bias"))
    dh = torch.empty((B, Hd), device=x.device, dtype=torch.float32)
    T.sgemm(dlogits, 3, 1, w2, Hd, 1, dh, B, Hd, 3)
    T.dropout_apply(dh, drop)
    T.relu_bwd_(dh, pre)
    x32 = x.float().contiguous()
    T.sgemm(dh, 1, Hd, x32, K, 1, g("classifier.0.weight"), Hd, K, B)
    T.colsum(dh, g("classifier.0.bias"))
    return loss, logits


def train_model(model, dataloader, optimizer, device, criterion_mort, criterion_los, criterion_mech):
    """Drop-in for 02_BioClinicalBERT.py:137-152: one epoch, returns the mean batch loss.  The three criteria supply
    pos_weight (and the shared gamma / alpha); the optimiser supplies lr / betas / eps / weight_decay; there is no
    gradient clipping in this script."""
    model.train()
    gp = optimizer.param_groups[0]
    crits = (criterion_mort, criterion_los, criterion_mech)
    if len({(float(c.gamma), c.alpha) for c in crits}) != 1:
        raise NotImplementedError("the three focal losses must share gamma and alpha (as in the reference)")
    pw = torch.stack([torch.as_tensor(1.0 if c.pos_weight is None else c.pos_weight, dtype=torch.float32).reshape(-1)[0]
                      for c in crits]).to(device)
    st = get_state(model)
    total = torch.zeros(1, device=device, dtype=torch.float64)
    n = 0
    for batch in dataloader:
        emb, lm, ll, lc = [b.to(device, non_blocking=True) for b in batch]
        labels = torch.cat([lm.reshape(-1, 1), ll.reshape(-1, 1), lc.reshape(-1, 1)], dim=1).float()
        loss, _ = forward_backward(model, emb, labels, pw, float(crits[0].gamma), crits[0].alpha)
        st.clip_and_step(gp["lr"], gp.get("weight_decay", 0.01), tuple(gp.get("betas", (0.9, 0.999))), gp.get("eps", 1e-8),
                         max_norm=1e30)                       # optimizer.step() without clip_grad_norm_
        total += loss
        n += 1
    return float(total.item()) / max(n, 1)


def apply_bioclinicalbert_on_patient_notes(df, note_columns, tokenizer, model, device, aggregation="mean", max_length=512):
    """02_BioClinicalBERT.py:72-105: as the FAME variant, but returns (embeddings, patient_ids)."""
    emb = modules.apply_bioclinicalbert_on_patient_notes(df, note_columns, tokenizer, model, device, aggregation=aggregation,
                                                         max_length=max_length)
    return emb, df["subject_id"].unique()

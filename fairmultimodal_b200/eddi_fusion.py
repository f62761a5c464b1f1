"""Per-batch EDDI-weighted logit-fusion ablation of the reference (FinalCode/New/Final/08_multimodal_eddi_fusion.py) on
the B200 kernels: SURVEY.md 8(f-3).

    BEHRTModel_Demo        08:253-292   the demographic encoder with a SIX-layer, six-head BERT (128 positions)
    BEHRTModel_Lab         08:294-312   the lab tower of 10_FAME.py
    MultimodalTransformer  08:315-452   ReLU(Linear 768->256) per modality -> nine scalar heads classifier_{demo,lab,
                                        text}_{mort,los,mv} -> per outcome, compute_eddi (08:45-59) of the three
                                        modalities' thresholded predictions over the batch's sensitive attribute, weights
                                        w_m = (old_w_m | 0.33) + beta * (max EDDI - EDDI_m), fused logit = sum_m w_m raw_m
                                        (the weights are constants for autograd: they come from detached logits)
    train_step             08:454-493   three FocalLoss(gamma=1, pos_weight_i) + loss_gamma * mean((mortality logit -
                                        target)^2) -> backward -> clip 1.0 -> optimizer.step(); returns the SUM of losses
    validate_step          08:495-533

The in-forward EDDI is integer counting: fame_eval_counts (the evaluation count kernel, K10) is launched once per outcome
with the three modality logits as its three columns and the outcome's labels repeated; under data parallel the count
vectors are SUM all-reduced, so every rank derives the same weights as a single process on the concatenated batch.  The
float64 formula then runs on the host exactly as the reference's numpy code does (the reference also leaves the device
here: 08:368-376) -- one 22 KB device->host read per forward.
The towers, their hand-written backward, dropout, clip + AdamW and the flat training state are train.py's; the head is
fp32 and specific to this model.  Same class names, constructor arguments, forward signature, returned structure,
stdout lines and state_dict keys as the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import modules, ops
from . import ops_train as T
from . import train
from .bert import BertModelB200
from .unstructured import FocalLoss  # noqa: F401  (08:23-43 defines the same class)

OUTCOMES = ("mortality", "los", "mechanical_ventilation")
_O = ("mort", "los", "mv")
_M = ("demo", "lab", "text")
NAMES = dict(demo="behrt_demo.", lab="behrt_lab.",
             head=("demo_projector.", "lab_projector.", "text_projector.", "classifier_"))
NO_GRAD = ("behrt_demo.bert.pooler.",)            # computed by HF BertModel, unused by the model: grad None in the reference
_PROJ = ("demo_projector.0.", "lab_projector.0.", "text_projector.0.")


class BEHRTModel_Demo(modules.BEHRTModel_Demo):
    """08:253-292: as 10_FAME.py's class but num_hidden_layers=6, num_attention_heads=6, max_position_embeddings=128."""

    def __init__(self, num_ages, num_genders, num_ethnicities, num_insurances, hidden_size=768):
        nn.Module.__init__(self)
        vocab_size = num_ages + num_genders + num_ethnicities + num_insurances + 2
        self.bert = BertModelB200(vocab_size, hidden_size, 6, 6, 3072, 128)
        self.age_embedding = nn.Embedding(num_ages, hidden_size)
        self.gender_embedding = nn.Embedding(num_genders, hidden_size)
        self.ethnicity_embedding = nn.Embedding(num_ethnicities, hidden_size)
        self.insurance_embedding = nn.Embedding(num_insurances, hidden_size)


BEHRTModel_Lab = modules.BEHRTModel_Lab


def eddi_from_counts(cells):
    """compute_eddi (08:45-59) from the integer confusion cells of one (outcome, modality): cells int [8 codes][TP,FN,FP,TN].
    Groups = the codes present (np.unique); returns (eddi, {code: subgroup value})."""
    cells = np.asarray(cells, dtype=np.int64).reshape(8, 4)
    n_g = cells.sum(axis=1)
    e_g = cells[:, 1] + cells[:, 2]
    n = int(n_g.sum())
    if n == 0:
        return float("nan"), {}
    err = np.float64(int(e_g.sum())) / np.float64(n)
    denom = max(err, 1 - err) if err not in [0, 1] else 1.0
    sub = {}
    for c in range(8):
        if n_g[c] > 0:
            sub[c] = (np.float64(int(e_g[c])) / np.float64(int(n_g[c])) - err) / denom
    return float(np.sqrt(np.sum(np.array(list(sub.values())) ** 2)) / len(sub)), sub


def fusion_weights(eddi3, beta, old=None):
    """08:385-395."""
    top = max(eddi3)
    base = old if old is not None else (0.33, 0.33, 0.33)
    return tuple(base[k] + beta * (top - eddi3[k]) for k in range(3)), top


class MultimodalTransformer(nn.Module):
    def __init__(self, text_embed_size, behrt_demo, behrt_lab, device, beta=0.3):
        super().__init__()
        if text_embed_size != 768:
            raise ValueError("built for the reference's sizes (768 -> 3 x 256 -> 9 scalar heads)")
        self.beta = beta
        self.device = device
        self.behrt_demo = behrt_demo
        self.behrt_lab = behrt_lab
        self.demo_projector = nn.Sequential(nn.Linear(behrt_demo.bert.config.hidden_size, 256), nn.ReLU())
        self.lab_projector = nn.Sequential(nn.Linear(behrt_lab.hidden_size, 256), nn.ReLU())
        self.text_projector = nn.Sequential(nn.Linear(text_embed_size, 256), nn.ReLU())
        for o in _O:                                                   # registration order of 08:338-348
            for m in _M:
                setattr(self, f"classifier_{m}_{o}", nn.Linear(256, 1))

    # ---- 08:350-411 on given projections and heads (API parity; forward() below batches the three outcomes)
    def compute_weighted_logit(self, demo_proj, lab_proj, text_proj, classifier_demo, classifier_lab, classifier_text, beta,
                               y_true, sensitive_labels, old_weights=None):
        with torch.no_grad():
            raw = [_lin(p.float().contiguous(), c.weight.detach().float().contiguous(), c.bias.detach().float().contiguous())
                   for p, c in ((demo_proj, classifier_demo), (lab_proj, classifier_lab), (text_proj, classifier_text))]
            raw3 = torch.cat(raw, dim=1).contiguous()                    # [B,3]: demo | lab | text
            if y_true is not None and sensitive_labels is not None:
                counts = _outcome_counts(raw3, _dev(y_true, raw3.device, torch.float32), _dev(sensitive_labels, raw3.device, torch.int64))
                cells = counts.cpu().numpy()[:288].reshape(3, 3, 8, 4)[:, 0]
            else:
                cells = None
            fused, det = _fuse_one(raw3, cells, beta, old_weights)
        return fused, det

    def forward(self, demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids, lab_features,
                aggregated_text_embedding, beta=None, y_true_dict=None, sensitive_labels_dict=None, old_eddi_weights=None):
        if not lab_features.is_cuda:
            raise RuntimeError("runs on a B200 only: move inputs to cuda (no CPU fallback)")
        beta = self.beta if beta is None else beta
        with torch.no_grad():
            if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                st = get_state(self)
                ds = _drop_sites(self, st)
                demo, _ = train._demo_forward(st, self, demo_dummy_ids, age_ids, gender_ids, ethnicity_ids, insurance_ids, ds)
                lab, _ = train._lab_forward(st, self, lab_features, ds)
                w = _head_weights(st.f)
            else:
                demo = self.behrt_demo(demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids)
                lab = self.behrt_lab(lab_features)
                sd = dict(self.named_parameters())
                w = _head_weights(lambda n: sd[n].detach().float().contiguous())
            h = _head_forward((demo, lab, aggregated_text_embedding), w)
            cells = _batch_cells(h["raw"], y_true_dict, sensitive_labels_dict, None)
            logits, details = _fuse(h["raw"], cells, beta, old_eddi_weights)
        return logits[:, 0:1], logits[:, 1:2], logits[:, 2:3], details


def _dev(x, device, dtype):
    t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype).reshape(-1).contiguous()


def _head_weights(f):
    """Projector weights + the nine scalar heads stacked per modality: wc[m] [3 outcomes, 256], bc[m] [3]."""
    wc = [torch.cat([f(f"classifier_{m}_{o}.weight") for o in _O], dim=0).contiguous() for m in _M]
    bc = [torch.cat([f(f"classifier_{m}_{o}.bias") for o in _O], dim=0).contiguous() for m in _M]
    return dict(wp=[f(p + "weight") for p in _PROJ], bp=[f(p + "bias") for p in _PROJ], wc=wc, bc=bc)


def _lin(x, w, b):
    """x [B,K] fp32 @ w[N,K]^T + b -> [B,N] (fame_sgemm_small)."""
    B, K = x.shape
    N = w.shape[0]
    y = b.repeat(B, 1)
    T.sgemm(x, K, 1, w, 1, K, y, B, N, K, accumulate=True)
    return y


def _head_forward(embs, w):
    """08:420-423 + the raw modality logits of 08:353-355.  raw [B, 3 modalities, 3 outcomes] fp32."""
    embs = [e.float().contiguous() for e in embs]
    pre_p = [_lin(embs[m], w["wp"][m], w["bp"][m]) for m in range(3)]                 # [B,256] pre-ReLU
    proj = [T.relu_(p.clone()) for p in pre_p]
    raw = torch.stack([_lin(proj[m], w["wc"][m], w["bc"][m]) for m in range(3)], dim=1).contiguous()
    return dict(embs=embs, pre_p=pre_p, proj=proj, raw=raw)


def _outcome_counts(raw3, y, sens, out=None):
    """fame_eval_counts over one outcome: columns = the three modality logits, labels = y repeated, attribute = sens.
    The cells of (column m, attribute 0, code) are out[((m*3 + 0)*8 + code)*4 + {TP,FN,FP,TN}]."""
    labels3 = y.reshape(-1, 1).expand(-1, 3).contiguous()
    return ops.eval_counts(raw3, labels3, [sens, sens, sens], (0.5, 0.5, 0.5), out=out)


def _batch_cells(raw, y_true_dict, sensitive_labels_dict, group):
    """Integer confusion cells [3 outcomes][3 modalities][8 codes][4] of the batch (global batch under data parallel), or
    None per outcome whose labels / sensitive attribute are missing (08:367-380)."""
    have = [y_true_dict is not None and sensitive_labels_dict is not None and o in y_true_dict and o in sensitive_labels_dict
            and y_true_dict[o] is not None and sensitive_labels_dict[o] is not None for o in OUTCOMES]
    if not any(have):
        return [None, None, None]
    dev = raw.device
    counts = torch.zeros((3, ops._lib.EVAL_COUNTS_LEN), device=dev, dtype=torch.int64)
    for oi, o in enumerate(OUTCOMES):
        if have[oi]:
            _outcome_counts(raw[:, :, oi].contiguous(), _dev(y_true_dict[o], dev, torch.float32),
                            _dev(sensitive_labels_dict[o], dev, torch.int64), out=counts[oi])
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(counts, group=group)
    host = counts.cpu().numpy()
    if host[:, 913].any():
        raise ValueError("sensitive-attribute codes must lie in 0..7")
    return [host[oi, :288].reshape(3, 3, 8, 4)[:, 0] if have[oi] else None for oi in range(3)]


def _fuse_one(raw3, cells, beta, old):
    """One outcome: raw3 [B,3] (demo, lab, text), cells [3 modalities][8][4] or None.  Prints what 08:377, 380, 397 print."""
    if cells is not None:
        vals = [eddi_from_counts(cells[m]) for m in range(3)]
        e3, subs = [v[0] for v in vals], tuple(v[1] for v in vals)
        print(f"Computed EDDI - Demo: {e3[0]:.4f}, Lab: {e3[1]:.4f}, Text: {e3[2]:.4f}")
    else:
        e3, subs = [0.0, 0.0, 0.0], ({}, {}, {})
        print("No y_true or sensitive_labels provided, setting EDDI values to 0.")
    wts, top = fusion_weights(e3, beta, old)
    print(f"Modality weights - Demo: {wts[0]:.4f}, Lab: {wts[1]:.4f}, Text: {wts[2]:.4f}")
    wt = torch.tensor(wts, dtype=torch.float32, device=raw3.device)                # python floats meet fp32 tensors
    fused = (raw3 * wt).sum(dim=1, keepdim=True)
    probs = torch.sigmoid(raw3)
    det = {"eddi": (e3[0], e3[1], e3[2], top), "weights": wts,
           "probs": (probs[:, 0:1], probs[:, 1:2], probs[:, 2:3]), "subgroups": subs}
    return fused, det


def _fuse(raw, cells, beta, old_eddi_weights):
    """All three outcomes.  Returns (logits [B,3], eddi_details as 08:446-449)."""
    cols, details = [], {}
    for oi, o in enumerate(OUTCOMES):
        old = old_eddi_weights.get(o) if old_eddi_weights is not None else None
        fused, det = _fuse_one(raw[:, :, oi], cells[oi], beta, old)
        cols.append(fused)
        details[o] = det
    return torch.cat(cols, dim=1).contiguous(), details


def get_state(model) -> train.FlatTrainState:
    st = getattr(model, "_fame_train_state", None)
    if st is None or st.model is not model:
        st = train.FlatTrainState(model, no_grad_prefixes=NO_GRAD, fame_layout=True, names=NAMES)
        object.__setattr__(model, "_fame_train_state", st)
    return st


def _drop_sites(model, st):
    ds = train.DropSites(model, st.step_dev, lab_module=model.behrt_lab, head_dropout=None, demo_module=model.behrt_demo)
    return ds if ds.any else None


def _wgrad(dy, x, gw, gb):
    """dW[N,K] = dy^T x, db = colsum(dy) (fp32)."""
    B, N = dy.shape
    K = x.shape[1]
    T.sgemm(dy, 1, N, x, K, 1, gw, N, K, B)
    T.colsum(dy, gb)


def _dgrad(dy, w):
    B, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((B, K), device=dy.device, dtype=torch.float32)
    T.sgemm(dy, N, 1, w, K, 1, dx, B, K, N)
    return dx


def objective(logits, labels, pos_weight, loss_gamma=1.0, target=1.0, gamma=1.0, alpha=None, batch_total=None, want_grad=True):
    """08:475-479: three summed focal losses + loss_gamma * mean((mortality logit - target)^2), means over the global
    batch.  Returns (loss f64 [1] -- this rank's share of the global mean, dlogits [B,3] | None)."""
    loss, dl = T.focal_loss_fwd_bwd(logits, labels.float().contiguous(), pos_weight, gamma, 1.0 if alpha is None else alpha,
                                    want_grad=want_grad, batch_total=batch_total)
    n = float(logits.shape[0]) if batch_total is None else batch_total.to(torch.float64)
    d = logits[:, 0].double() - target
    loss = loss + loss_gamma * (d * d).sum() / n
    if dl is not None:
        dl[:, 0] += (2.0 * loss_gamma * d / n).float()
    return loss, dl


def forward_backward(model, batch8, labels, pos_weight, beta=0.3, loss_gamma=1.0, target=1.0, old_eddi_weights=None,
                     gamma=1.0, alpha=None, group=None, sensitive=None):
    """Forward (with the in-forward EDDI over `sensitive`, default gender_ids as 08:468-472) + objective + backward of
    one batch; gradients land in the flat buffer.  batch8 = (demo_dummy_ids, demo_attn_mask, age, gender, ethnicity,
    insurance, lab_features, text_embedding); labels f32 [B,3].  Returns (loss f64 [1], logits [B,3], eddi_details)."""
    st = get_state(model)
    ids, _, age, gender, eth, ins, lab, text = batch8
    sensitive = gender if sensitive is None else sensitive
    post = st.post_stream()
    for w_ in train.begin_step(st, group).values():
        w_.wait()
    ds = _drop_sites(model, st)
    demo, sv_d = train._demo_forward(st, model, ids, age, gender, eth, ins, ds)
    labe, sv_l = train._lab_forward(st, model, lab, ds)
    w = _head_weights(st.f)
    h = _head_forward((demo, labe, text), w)
    yd = {o: labels[:, oi] for oi, o in enumerate(OUTCOMES)}
    cells = _batch_cells(h["raw"], yd, {o: sensitive for o in OUTCOMES}, group)
    logits, details = _fuse(h["raw"], cells, beta, old_eddi_weights)
    btot = None
    if group is not None:
        import torch.distributed as dist
        btot = torch.full((1,), labels.shape[0], device=labels.device, dtype=torch.int64)
        dist.all_reduce(btot, group=group)
    loss, dlogits = objective(logits, labels, pos_weight, loss_gamma, target, gamma, alpha, batch_total=btot)
    if group is not None:
        dist.all_reduce(loss, group=group)
    torch.cuda.current_stream().wait_stream(post)
    red = train._GradReducer(st, group)
    g = st.gr
    # fused[:, o] = sum_m w[o][m] raw[:, m, o]  with constant weights
    wmat = torch.tensor([details[o]["weights"] for o in OUTCOMES], dtype=torch.float32, device=logits.device)   # [o][m]
    demb = []
    for m in range(3):
        draw = (dlogits * wmat[:, m]).contiguous()                                   # [B,3 outcomes]
        gw = torch.zeros((3, 256), device=logits.device, dtype=torch.float32)      # colsum accumulates: start from zero
        gb = torch.zeros(3, device=logits.device, dtype=torch.float32)
        _wgrad(draw, h["proj"][m], gw, gb)
        for oi, o in enumerate(_O):
            g(f"classifier_{_M[m]}_{o}.weight").copy_(gw[oi:oi + 1])
            g(f"classifier_{_M[m]}_{o}.bias").copy_(gb[oi:oi + 1])
        dproj = _dgrad(draw, w["wc"][m])                                             # [B,256]
        T.relu_bwd_(dproj, h["pre_p"][m])
        _wgrad(dproj, h["embs"][m], g(_PROJ[m] + "weight"), g(_PROJ[m] + "bias"))
        if m < 2:
            demb.append(_dgrad(dproj, w["wp"][m]))
    train._demo_backward(st, model, sv_d, demb[0], red, ds)
    train._lab_backward(st, model, sv_l, demb[1], ds, reducer=red)
    red.ready("tail")
    red.finish()
    return loss, logits, details


def _pos_weights(crits, device):
    if len({(float(c.gamma), c.alpha) for c in crits}) != 1:
        raise NotImplementedError("the three focal losses must share gamma and alpha (as in the reference)")
    return torch.stack([torch.as_tensor(1.0 if c.pos_weight is None else c.pos_weight, dtype=torch.float32).reshape(-1)[0]
                        for c in crits]).to(device)


def train_step(model, dataloader, optimizer, device, beta=0.3, loss_gamma=1.0, target=1.0, old_eddi_weights=None,
               criterion_mortality=None, criterion_los=None, criterion_mech=None, group=None):
    """Drop-in for 08:454-493 (the reference reads its three criteria from module globals; here they are arguments):
    one epoch, returns the SUM of the batch losses."""
    model.train()
    crits = (criterion_mortality, criterion_los, criterion_mech)
    if any(c is None for c in crits):
        raise ValueError("pass criterion_mortality / criterion_los / criterion_mech (FocalLoss objects)")
    pw = _pos_weights(crits, device)
    gp = optimizer.param_groups[0]
    st = get_state(model)
    total = torch.zeros(1, device=device, dtype=torch.float64)
    for batch in dataloader:
        b = [x.to(device, non_blocking=True) for x in batch]
        labels = torch.stack([b[8].reshape(-1), b[9].reshape(-1), b[10].reshape(-1)], dim=1).float()
        loss, _, _ = forward_backward(model, b[:8], labels, pw, beta, loss_gamma, target, old_eddi_weights,
                                      float(crits[0].gamma), crits[0].alpha, group=group)
        st.clip_and_step(gp["lr"], gp.get("weight_decay", 0.01), tuple(gp.get("betas", (0.9, 0.999))), gp.get("eps", 1e-8),
                         max_norm=1.0)
        total += loss
    train.sync_parameters(model, group)
    return float(total.item())


def validate_step(model, dataloader, device, beta=0.3, loss_gamma=1.0, target=1.0, old_eddi_weights=None,
                  criterion_mortality=None, criterion_los=None, criterion_mech=None):
    """Drop-in for 08:495-533: (sum of batch losses, eddi_details of the last batch)."""
    model.eval()
    crits = (criterion_mortality, criterion_los, criterion_mech)
    if any(c is None for c in crits):
        raise ValueError("pass criterion_mortality / criterion_los / criterion_mech (FocalLoss objects)")
    pw = _pos_weights(crits, device)
    total, last = 0.0, None
    with torch.no_grad():
        for batch in dataloader:
            b = [x.to(device, non_blocking=True) for x in batch]
            labels = torch.stack([b[8].reshape(-1), b[9].reshape(-1), b[10].reshape(-1)], dim=1).float()
            yd = {o: labels[:, oi] for oi, o in enumerate(OUTCOMES)}
            a, c, d, last = model(*b[:8], beta=beta, y_true_dict=yd, sensitive_labels_dict={o: b[3] for o in OUTCOMES},
                                  old_eddi_weights=old_eddi_weights)
            loss, _ = objective(torch.cat([a, c, d], dim=1).contiguous(), labels, pw, loss_gamma, target,
                                float(crits[0].gamma), crits[0].alpha, want_grad=False)
            total += float(loss.item())
    return total, last

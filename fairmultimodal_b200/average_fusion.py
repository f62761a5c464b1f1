"""Concat-fusion ablation of the reference (FinalCode/New/Final/07_multimodal_average_fusion.py) on the B200 kernels:
SURVEY.md 8(f-3).

    BEHRTModel             07:156-203   a 12-layer BERT over the length-1 dummy token + the mean of SEVEN clamped code
                                        embeddings (age, segment, admission / discharge location, gender, ethnicity,
                                        insurance)
    MultimodalTransformer  07:205-238   ReLU(ts_linear 768->256) | ReLU(text_linear 768->256) -> concat 512 ->
                                        classifier (512->512, ReLU, Dropout .1, 512->3); returns (mortality, los, vent
                                        logits [B,1], fused_embedding_pre_relu [B,512])
    train_step             07:240-264   three summed FocalLoss(gamma=1, pos_weight_i) -> backward -> optimizer.step()
                                        (torch.optim.Adam, NO gradient clipping, 07:720); returns the SUM of batch losses

The BERT tower, its hand-written backward, dropout, the flat training state and the fused Adam kernel are the ones of
train.py (the tower is the demographic encoder of 10_FAME.py with seven instead of four code tables:
fame_embed_mean_add); the fp32 head is specific to this model.  Same class names, constructor arguments, forward
signatures and state_dict keys as the reference.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import ops
from . import ops_train as T
from . import train
from .bert import BertModelB200
from .unstructured import FocalLoss  # noqa: F401  (07:18-38 defines the same class)

TABLES = ("age", "segment", "admission_loc", "discharge_loc", "gender", "ethnicity", "insurance")
NAMES = dict(demo="BEHRT.", lab="(no lab tower).", head=("ts_linear.", "text_linear.", "classifier."))
NO_GRAD = ("BEHRT.bert.pooler.",)                 # computed by HF BertModel, unused by the model: grad None in the reference
_NO_LAB = types.SimpleNamespace(transformer_encoder=types.SimpleNamespace(layers=[]))


class BEHRTModel(nn.Module):
    def __init__(self, num_diseases, num_ages, num_segments, num_admission_locs, num_discharge_locs, num_genders,
                 num_ethnicities, num_insurances, hidden_size=768):
        super().__init__()
        vocab_size = num_diseases + num_ages + num_segments + num_admission_locs + num_discharge_locs + 2
        self.bert = BertModelB200(vocab_size, hidden_size, 12, 12, 3072, 512)
        self.age_embedding = nn.Embedding(num_ages, hidden_size)
        self.segment_embedding = nn.Embedding(num_segments, hidden_size)
        self.admission_loc_embedding = nn.Embedding(num_admission_locs, hidden_size)
        self.discharge_loc_embedding = nn.Embedding(num_discharge_locs, hidden_size)
        self.gender_embedding = nn.Embedding(num_genders, hidden_size)
        self.ethnicity_embedding = nn.Embedding(num_ethnicities, hidden_size)
        self.insurance_embedding = nn.Embedding(num_insurances, hidden_size)

    def _tables(self):
        return [getattr(self, n + "_embedding").weight for n in TABLES]

    def forward(self, input_ids, attention_mask, age_ids, segment_ids, adm_loc_ids, disch_loc_ids, gender_ids,
                ethnicity_ids, insurance_ids):
        if not input_ids.is_cuda:
            raise RuntimeError("runs on a B200 only: move inputs to cuda (no CPU fallback)")
        B, S = input_ids.shape
        h = self.bert.encode_f32(input_ids, attention_mask)           # f32 [B*S, hidden], fp32 residual stream
        with torch.no_grad():
            return ops.embed_mean_add(h, S * h.shape[1],
                                      [age_ids, segment_ids, adm_loc_ids, disch_loc_ids, gender_ids, ethnicity_ids, insurance_ids],
                                      [t.detach().float() for t in self._tables()])


class MultimodalTransformer(nn.Module):
    def __init__(self, text_embed_size, BEHRT, device, hidden_size=512):
        super().__init__()
        if text_embed_size != 768 or hidden_size != 512:
            raise ValueError("built for the reference's sizes (768 -> 2 x 256 -> 512 -> 3)")
        self.BEHRT = BEHRT
        self.device = device
        self.ts_linear = nn.Linear(BEHRT.bert.config.hidden_size, 256)
        self.text_linear = nn.Linear(text_embed_size, 256)
        self.classifier = nn.Sequential(nn.Linear(256 + 256, hidden_size), nn.ReLU(), nn.Dropout(0.1),
                                        nn.Linear(hidden_size, 3))

    def forward(self, dummy_input_ids, dummy_attn_mask, age_ids, segment_ids, adm_loc_ids, discharge_loc_ids, gender_ids,
                ethnicity_ids, insurance_ids, aggregated_text_embedding):
        codes = (age_ids, segment_ids, adm_loc_ids, discharge_loc_ids, gender_ids, ethnicity_ids, insurance_ids)
        with torch.no_grad():
            if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                st = get_state(self)
                ds = _drop_sites(self, st)
                emb, _ = train._demo_forward(st, self, dummy_input_ids, None, None, None, None, ds, demo_module=self.BEHRT,
                                             dpre="BEHRT.", codes=codes, table_names=TABLES)
                w = _head_weights(st.f)
                drop = ds.site("avgfusion.head", ds.p_fusion) if ds is not None else None
            else:
                emb = self.BEHRT(dummy_input_ids, dummy_attn_mask, *codes)
                sd = dict(self.named_parameters())
                w = _head_weights(lambda n: sd[n].detach().float().contiguous())
                drop = None
            h = _head_forward(emb, aggregated_text_embedding, w, drop)
        lg = h["logits"]
        return lg[:, 0:1], lg[:, 1:2], lg[:, 2:3], torch.cat([h["ts_pre"], h["tx_pre"]], dim=1)


def _head_weights(f):
    return dict(wts=f("ts_linear.weight"), bts=f("ts_linear.bias"), wtx=f("text_linear.weight"), btx=f("text_linear.bias"),
                w1=f("classifier.0.weight"), b1=f("classifier.0.bias"), w2=f("classifier.3.weight"), b2=f("classifier.3.bias"))


def _lin(x, w, b):
    """x [B,K] fp32 @ w[N,K]^T + b -> [B,N] (fame_sgemm_small)."""
    B, K = x.shape
    N = w.shape[0]
    y = b.repeat(B, 1)
    T.sgemm(x, K, 1, w, 1, K, y, B, N, K, accumulate=True)
    return y


def _head_forward(emb, text, w, drop):
    """07:226-238.  Returns every intermediate the backward needs."""
    emb, text = emb.float().contiguous(), text.float().contiguous()
    ts_pre = _lin(emb, w["wts"], w["bts"])                               # [B,256]
    tx_pre = _lin(text, w["wtx"], w["btx"])
    comb = torch.cat([T.relu_(ts_pre.clone()), T.relu_(tx_pre.clone())], dim=1).contiguous()
    pre_h = _lin(comb, w["w1"], w["b1"])                                 # [B,512]
    hid = T.relu_(pre_h.clone())
    T.dropout_apply(hid, drop)
    logits = _lin(hid, w["w2"], w["b2"])
    return dict(emb=emb, text=text, ts_pre=ts_pre, tx_pre=tx_pre, comb=comb, pre_h=pre_h, hid=hid, logits=logits)


def get_state(model) -> train.FlatTrainState:
    st = getattr(model, "_fame_train_state", None)
    if st is None or st.model is not model:
        st = train.FlatTrainState(model, no_grad_prefixes=NO_GRAD, fame_layout=True, names=NAMES)
        object.__setattr__(model, "_fame_train_state", st)
    return st


def _drop_sites(model, st):
    ds = train.DropSites(model, st.step_dev, lab_module=_NO_LAB, head_dropout=model.classifier[2], demo_module=model.BEHRT)
    return ds if ds.any else None


def _wgrad(dy, x, gw, gb):
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


def forward_backward(model, batch10, labels, pos_weight, gamma=1.0, alpha=None, group=None):
    """Forward + summed focal loss + backward of one batch (07:251-262); gradients land in the flat buffer.
    batch10 = (dummy_input_ids, dummy_attn_mask, age, segment, adm_loc, disch_loc, gender, ethnicity, insurance,
    text_embedding); labels f32 [B,3].  Returns (loss f64 [1], logits [B,3])."""
    st = get_state(model)
    ids, _, *codes, text = batch10
    post = st.post_stream()
    for w_ in train.begin_step(st, group).values():
        w_.wait()
    ds = _drop_sites(model, st)
    emb, saved = train._demo_forward(st, model, ids, None, None, None, None, ds, demo_module=model.BEHRT, dpre="BEHRT.",
                                     codes=codes, table_names=TABLES)
    w = _head_weights(st.f)
    drop = ds.site("avgfusion.head", ds.p_fusion) if ds is not None else None
    h = _head_forward(emb, text, w, drop)
    btot = None
    if group is not None:
        import torch.distributed as dist
        btot = torch.full((1,), labels.shape[0], device=labels.device, dtype=torch.int64)
        dist.all_reduce(btot, group=group)
    loss, dlogits = T.focal_loss_fwd_bwd(h["logits"], labels.float().contiguous(), pos_weight, gamma,
                                         1.0 if alpha is None else alpha, batch_total=btot)
    if group is not None:
        dist.all_reduce(loss, group=group)
    torch.cuda.current_stream().wait_stream(post)
    red = train._GradReducer(st, group)
    g = st.gr
    _wgrad(dlogits, h["hid"], g("classifier.3.weight"), g("classifier.3.bias"))
    dhid = _dgrad(dlogits, w["w2"])
    T.dropout_apply(dhid, drop)
    T.relu_bwd_(dhid, h["pre_h"])
    _wgrad(dhid, h["comb"], g("classifier.0.weight"), g("classifier.0.bias"))
    dcomb = _dgrad(dhid, w["w1"])                                                   # [B,512]
    dts = dcomb[:, :256].contiguous()
    T.relu_bwd_(dts, h["ts_pre"])
    dtx = dcomb[:, 256:].contiguous()
    T.relu_bwd_(dtx, h["tx_pre"])
    _wgrad(dts, h["emb"], g("ts_linear.weight"), g("ts_linear.bias"))
    _wgrad(dtx, h["text"], g("text_linear.weight"), g("text_linear.bias"))
    demb = _dgrad(dts, w["wts"])
    train._demo_backward(st, model, saved, demb, red, ds, dpre="BEHRT.")
    red.ready("tail")
    red.finish()
    return loss, h["logits"]


def _adam_hyper(optimizer):
    gp = optimizer.param_groups[0]
    wd = gp.get("weight_decay", 0.0)
    if isinstance(optimizer, torch.optim.Adam) and not isinstance(optimizer, torch.optim.AdamW) and wd != 0:
        raise NotImplementedError("torch.optim.Adam with L2 weight_decay (the reference uses weight_decay = 0, 07:720); "
                                  "the fused kernel implements decoupled decay only")
    return gp["lr"], wd, tuple(gp.get("betas", (0.9, 0.999))), gp.get("eps", 1e-8)


def train_step(model, dataloader, optimizer, device, crit_mort, crit_los, crit_vent, group=None):
    """Drop-in for 07:240-264: one epoch, returns the SUM of the batch losses.  The criteria are FocalLoss objects (gamma
    / alpha shared, pos_weight per outcome); the optimiser (torch.optim.Adam in the reference) supplies lr / betas / eps;
    no gradient clipping."""
    model.train()
    crits = (crit_mort, crit_los, crit_vent)
    if len({(float(c.gamma), c.alpha) for c in crits}) != 1:
        raise NotImplementedError("the three focal losses must share gamma and alpha (as in the reference)")
    pw = torch.stack([torch.as_tensor(1.0 if c.pos_weight is None else c.pos_weight, dtype=torch.float32).reshape(-1)[0]
                      for c in crits]).to(device)
    lr, wd, betas, eps = _adam_hyper(optimizer)
    st = get_state(model)
    total = torch.zeros(1, device=device, dtype=torch.float64)
    for batch in dataloader:
        b = [x.to(device, non_blocking=True) for x in batch]
        labels = torch.stack([b[10].reshape(-1), b[11].reshape(-1), b[12].reshape(-1)], dim=1).float()
        loss, _ = forward_backward(model, b[:10], labels, pw, float(crits[0].gamma), crits[0].alpha, group=group)
        st.clip_and_step(lr, wd, betas, eps, max_norm=1e30)              # optimizer.step() without clip_grad_norm_
        total += loss
    train.sync_parameters(model, group)
    return float(total.item())

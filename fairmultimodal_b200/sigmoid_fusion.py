"""Per-modality sigmoid-gate ablation of the reference (FinalCode/New/Final/09_multimodal_sigmoid_fusion.py) on the B200
kernels: SURVEY.md 8(f-3).

    MultimodalTransformer   09:162-222   demo tower (attribute BEHRT) + lab tower (behrt_lab) + text embedding ->
                                         ReLU(Linear 768->256) x3 -> * sigmoid(sig_weights_{demo,lab,text}) -> concat 768 ->
                                         aggregate_projector (768->512, ReLU) -> classifier (512->512, ReLU, Dropout .1,
                                         512->3); returns (mortality, los, mech logits [B,1], aggregated [B,512])
    train_step              09:464-488   three summed FocalLoss(gamma=1, pos_weight_i) -> backward -> clip 1.0 -> AdamW

The two towers, their hand-written backward passes, dropout, clip + AdamW and the flat training state are the ones of
train.py (the towers are the same classes as in 10_FAME.py); the head is fp32 and specific to this model.  Same class
name, constructor, forward signature and state_dict keys as the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from . import ops_train as T
from . import train

NAMES = dict(demo="BEHRT.", lab="behrt_lab.",
             head=("demo_projector.", "lab_projector.", "text_projector.", "sig_weights_", "aggregate_projector.",
                   "classifier."))
NO_GRAD = ("BEHRT.bert.pooler.",)                 # computed by HF BertModel, unused by the model: grad None in the reference
_PROJ = ("demo_projector.0.", "lab_projector.0.", "text_projector.0.")
_SIG = ("sig_weights_demo", "sig_weights_lab", "sig_weights_text")


class MultimodalTransformer(nn.Module):
    def __init__(self, text_embed_size, BEHRT, behrt_lab, device, hidden_size=512):
        super().__init__()
        if text_embed_size != 768 or hidden_size != 512:
            raise ValueError("built for the reference's sizes (768 -> 3 x 256 -> 512 -> 512 -> 3)")
        self.BEHRT = BEHRT
        self.behrt_lab = behrt_lab
        self.device = device
        self.demo_projector = nn.Sequential(nn.Linear(BEHRT.bert.config.hidden_size, 256), nn.ReLU())
        self.lab_projector = nn.Sequential(nn.Linear(behrt_lab.hidden_size, 256), nn.ReLU())
        self.text_projector = nn.Sequential(nn.Linear(text_embed_size, 256), nn.ReLU())
        self.sig_weights_demo = nn.Parameter(torch.randn(256))
        self.sig_weights_lab = nn.Parameter(torch.randn(256))
        self.sig_weights_text = nn.Parameter(torch.randn(256))
        self.aggregate_projector = nn.Sequential(nn.Linear(768, 512), nn.ReLU())
        self.classifier = nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.Dropout(0.1), nn.Linear(512, 3))

    def forward(self, demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids, lab_features,
                aggregated_text_embedding):
        if not lab_features.is_cuda:
            raise RuntimeError("runs on a B200 only: move inputs to cuda (no CPU fallback)")
        with torch.no_grad():
            if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                st = get_state(self)
                ds = _drop_sites(self, st)
                demo, _ = train._demo_forward(st, self, demo_dummy_ids, age_ids, gender_ids, ethnicity_ids, insurance_ids, ds,
                                              demo_module=self.BEHRT, dpre="BEHRT.")
                lab, _ = train._lab_forward(st, self, lab_features, ds)
                w = _head_weights(st.f)
                drop = ds.site("sigfusion.head", ds.p_fusion) if ds is not None else None
            else:
                demo = self.BEHRT(demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids)
                lab = self.behrt_lab(lab_features)
                sd = dict(self.named_parameters())
                w = _head_weights(lambda n: sd[n].detach().float().contiguous())
                drop = None
            h = _head_forward((demo, lab, aggregated_text_embedding), w, drop)
        lg = h["logits"]
        return lg[:, 0:1], lg[:, 1:2], lg[:, 2:3], h["agg"]


def _head_weights(f):
    return dict(wp=[f(p + "weight") for p in _PROJ], bp=[f(p + "bias") for p in _PROJ], sig=[f(n) for n in _SIG],
                wa=f("aggregate_projector.0.weight"), ba=f("aggregate_projector.0.bias"),
                w1=f("classifier.0.weight"), b1=f("classifier.0.bias"), w2=f("classifier.3.weight"), b2=f("classifier.3.bias"))


def _lin(x, w, b):
    """x [B,K] fp32 @ w[N,K]^T + b -> [B,N] (fame_sgemm_small)."""
    B, K = x.shape
    N = w.shape[0]
    y = b.repeat(B, 1)
    T.sgemm(x, K, 1, w, 1, K, y, B, N, K, accumulate=True)
    return y


def _head_forward(embs, w, drop):
    """09:197-216.  Returns every intermediate the backward needs."""
    pre_p, gate, cat = [], [], []
    for m in range(3):
        pp = _lin(embs[m].float().contiguous(), w["wp"][m], w["bp"][m])            # [B,256] pre-ReLU
        pre_p.append(pp)
        g = torch.sigmoid(w["sig"][m])
        gate.append(g)
        cat.append(T.relu_(pp.clone()) * g)
    conc = torch.cat(cat, dim=1).contiguous()                                         # [B,768]
    pre_a = _lin(conc, w["wa"], w["ba"])
    agg = T.relu_(pre_a.clone())                                                      # [B,512] (returned by the model)
    pre_h = _lin(agg, w["w1"], w["b1"])
    hid = T.relu_(pre_h.clone())
    T.dropout_apply(hid, drop)
    logits = _lin(hid, w["w2"], w["b2"])
    return dict(pre_p=pre_p, gate=gate, conc=conc, pre_a=pre_a, agg=agg, pre_h=pre_h, hid=hid, logits=logits)


def get_state(model) -> train.FlatTrainState:
    st = getattr(model, "_fame_train_state", None)
    if st is None or st.model is not model:
        st = train.FlatTrainState(model, no_grad_prefixes=NO_GRAD, fame_layout=True, names=NAMES)
        object.__setattr__(model, "_fame_train_state", st)
    return st


def _drop_sites(model, st):
    ds = train.DropSites(model, st.step_dev, lab_module=model.behrt_lab, head_dropout=model.classifier[2],
                         demo_module=model.BEHRT)
    return ds if ds.any else None


def _wgrad(dy, x, gw, gb):
    """dW[N,K] = dy^T x, db = colsum(dy) into the flat gradient buffer (fp32)."""
    B, N = dy.shape
    K = x.shape[1]
    T.sgemm(dy, 1, N, x, K, 1, gw, N, K, B)
    T.colsum(dy, gb)


def _dgrad(dy, w):
    """dx[B,K] = dy[B,N] @ w[N,K]."""
    B, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((B, K), device=dy.device, dtype=torch.float32)
    T.sgemm(dy, N, 1, w, K, 1, dx, B, K, N)
    return dx


def forward_backward(model, batch8, labels, pos_weight, gamma=1.0, alpha=None, group=None):
    """Forward + summed focal loss + backward of one batch (09:474-484); gradients land in the flat buffer.
    batch8 = (demo_dummy_ids, demo_attn_mask, age, gender, ethnicity, insurance, lab_features, text_embedding);
    labels f32 [B,3].  Returns (loss f64 [1], logits [B,3])."""
    st = get_state(model)
    ids, _, age, gender, eth, ins, lab, text = batch8
    post = st.post_stream()
    for w_ in train.begin_step(st, group).values():      # sharded optimizer: bf16 shadows of the large matrices
        w_.wait()
    ds = _drop_sites(model, st)
    demo, sv_d = train._demo_forward(st, model, ids, age, gender, eth, ins, ds, demo_module=model.BEHRT, dpre="BEHRT.")
    labe, sv_l = train._lab_forward(st, model, lab, ds)
    embs = (demo, labe, text.float().contiguous())
    w = _head_weights(st.f)
    drop = ds.site("sigfusion.head", ds.p_fusion) if ds is not None else None
    h = _head_forward(embs, w, drop)
    btot = None
    if group is not None:
        # N ranks == the single-process step on the concatenated batch: the focal mean runs over the GLOBAL batch (the
        # gradient all-reduce below is a SUM), and the returned loss is the global one
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
    # classifier: Linear 512->3 after dropout(relu(Linear 512->512))
    _wgrad(dlogits, h["hid"], g("classifier.3.weight"), g("classifier.3.bias"))
    dhid = _dgrad(dlogits, w["w2"])
    T.dropout_apply(dhid, drop)
    T.relu_bwd_(dhid, h["pre_h"])
    _wgrad(dhid, h["agg"], g("classifier.0.weight"), g("classifier.0.bias"))
    dagg = _dgrad(dhid, w["w1"])
    T.relu_bwd_(dagg, h["pre_a"])
    _wgrad(dagg, h["conc"], g("aggregate_projector.0.weight"), g("aggregate_projector.0.bias"))
    dconc = _dgrad(dagg, w["wa"])                                                   # [B,768]
    demb = []
    for m in range(3):
        dc = dconc[:, 256 * m:256 * (m + 1)].contiguous()
        relu_p = torch.relu(h["pre_p"][m])
        gate = h["gate"][m]
        # weighted = relu(pre) * sigmoid(s):  ds = sum_b dc * relu(pre) * sig (1 - sig);  dpre = dc * sig * [pre > 0]
        g(_SIG[m]).copy_((dc * relu_p).sum(0) * gate * (1 - gate))
        dpre = dc * gate
        T.relu_bwd_(dpre, h["pre_p"][m])
        _wgrad(dpre, embs[m].float().contiguous(), g(_PROJ[m] + "weight"), g(_PROJ[m] + "bias"))
        if m < 2:
            demb.append(_dgrad(dpre, w["wp"][m]))
    train._demo_backward(st, model, sv_d, demb[0], red, ds, dpre="BEHRT.")
    train._lab_backward(st, model, sv_l, demb[1], ds, reducer=red)
    red.ready("tail")
    red.finish()
    return loss, h["logits"]


def train_step(model, dataloader, optimizer, device, criterion_mortality, criterion_los, criterion_mech, group=None):
    """Drop-in for 09:464-488: one epoch, returns the mean batch loss.  The criteria are unstructured.FocalLoss objects
    (gamma / alpha shared, pos_weight per outcome); the optimiser supplies lr / betas / eps / weight_decay."""
    model.train()
    crits = (criterion_mortality, criterion_los, criterion_mech)
    if len({(float(c.gamma), c.alpha) for c in crits}) != 1:
        raise NotImplementedError("the three focal losses must share gamma and alpha (as in the reference)")
    pw = torch.stack([torch.as_tensor(1.0 if c.pos_weight is None else c.pos_weight, dtype=torch.float32).reshape(-1)[0]
                      for c in crits]).to(device)
    gp = optimizer.param_groups[0]
    st = get_state(model)
    total = torch.zeros(1, device=device, dtype=torch.float64)
    n = 0
    for batch in dataloader:
        b = [x.to(device, non_blocking=True) for x in batch]
        labels = torch.stack([b[8].reshape(-1), b[9].reshape(-1), b[10].reshape(-1)], dim=1).float()
        loss, _ = forward_backward(model, b[:8], labels, pw, float(crits[0].gamma), crits[0].alpha, group=group)
        st.clip_and_step(gp["lr"], gp.get("weight_decay", 0.01), tuple(gp.get("betas", (0.9, 0.999))), gp.get("eps", 1e-8),
                         max_norm=1.0)
        total += loss
        n += 1
    return float(total.item()) / max(n, 1)

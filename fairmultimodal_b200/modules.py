"""Drop-in modules for the classes of the reference's 10_FAME.py (same names, constructor arguments, forward
signatures, returned keys and ``state_dict`` layout), executed by the sm_100a kernel library.

    BioClinicalBERT_FT                     10_FAME.py:133-142
    apply_bioclinicalbert_on_patient_notes 10_FAME.py:144-173
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .bert import BertModelB200


class BioClinicalBERT_FT(nn.Module):
    """Frozen note encoder: CLS hidden state of each 512-token chunk (10_FAME.py:133-142).

    ``base_model`` may be a ``transformers.BertModel`` (as in the reference, 10_FAME.py:726-728; its weights are
    copied into the B200 encoder) or a ``BertModelB200``.  ``forward`` accepts any number of chunks per call;
    the reference calls it with one chunk at a time, here the natural batch is 256."""

    def __init__(self, base_model, config=None, device=None):
        super().__init__()
        if not isinstance(base_model, BertModelB200):
            base_model = BertModelB200.from_hf(base_model)
        self.BioBert = base_model
        self.device = device

    @property
    def bert(self):
        """Name of the encoder attribute in 02_BioClinicalBERT.py:62 (10_FAME.py calls it BioBert)."""
        return self.BioBert

    @classmethod
    def from_state_dict(cls, sd, prefix="BioBert.", **cfg):
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        vocab, hidden = sd["embeddings.word_embeddings.weight"].shape
        cfg.setdefault("num_hidden_layers", 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layer.")))
        bert = BertModelB200(vocab, hidden, **cfg)
        bert.load_state_dict(sd, strict=False)   # a pooler may be absent from hand-built dicts; it is unused
        return cls(bert)

    def encode_chunks(self, input_ids, attention_mask):
        """bf16 last hidden state of EVERY token [chunks*seq, hidden] (kept on device; CLS rows are every seq-th
        row).  The reference never reads the other rows: encode_cls is the hot path."""
        return self.BioBert.encode(input_ids, attention_mask)

    def encode_cls(self, input_ids, attention_mask):
        """bf16 CLS hidden state [chunks, hidden] (on device): what forward returns, before the cast to float32."""
        return self.BioBert.encode(input_ids, attention_mask, cls_only=True)

    def forward(self, input_ids, attention_mask):
        return self.encode_cls(input_ids, attention_mask).float()


def pool_chunks(cls_rows, offsets, mode="mean", ldx=None, cols=None):
    """Chunk -> patient aggregation on the device (10_FAME.py:153-154,170-172).  cls_rows [C, 768] f32/bf16 (or the
    full hidden state with ldx = seq*hidden), offsets int32 [P+1]."""
    return ops.segment_mean(cls_rows, offsets.to(torch.int32), cols=cols, ldx=ldx, mode=mode)


def _tokenize(tokenizer, note, max_length):
    kw = dict(add_special_tokens=True, max_length=max_length, padding="max_length", truncation=True,
              return_attention_mask=True, return_tensors="pt")
    enc = tokenizer.encode_plus(text=note, **kw) if hasattr(tokenizer, "encode_plus") else tokenizer(note, **kw)
    return enc["input_ids"].view(-1), enc["attention_mask"].view(-1)


def apply_bioclinicalbert_on_patient_notes(df, note_columns, tokenizer, model, device, aggregation="mean",
                                           chunks_per_batch=256, max_length=512):
    """Same contract as 10_FAME.py:144-173: one 768-vector per patient (first-appearance order of subject_id),
    the mean (or max) of the CLS vectors of that patient's non-empty note chunks in column-major order, zeros
    for note-less patients.  Returns np.ndarray [P, hidden].

    Differences in execution only: notes are gathered with one groupby instead of an O(P^2) filter, tokenised
    once, encoded 256 chunks per call and pooled on the device by CSR offsets."""
    pids = df["subject_id"].unique()
    groups = df.groupby("subject_id", sort=False).indices
    ids_l, mask_l, counts = [], [], []
    for pid in pids:
        rows = df.iloc[groups[pid]]
        n = 0
        for col in note_columns:
            for v in rows[col].dropna().tolist():
                if isinstance(v, str) and v.strip() != "":
                    i, m = _tokenize(tokenizer, v, max_length)
                    ids_l.append(i)
                    mask_l.append(m)
                    n += 1
        counts.append(n)
    hidden = model.BioBert.config.hidden_size
    offsets = np.zeros(len(pids) + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)
    if not ids_l:
        return np.zeros((len(pids), hidden))
    ids = torch.stack(ids_l).to(torch.int64).pin_memory()
    mask = torch.stack(mask_l).to(torch.int64).pin_memory()
    C = ids.shape[0]
    cls = torch.empty((C, hidden), device=device, dtype=torch.float32)
    for s in range(0, C, chunks_per_batch):
        e = min(C, s + chunks_per_batch)
        cls[s:e] = model(ids[s:e].to(device, non_blocking=True), mask[s:e].to(device, non_blocking=True))
    pooled = pool_chunks(cls, torch.from_numpy(offsets).to(device), mode="mean" if aggregation == "mean" else "max")
    return pooled.cpu().numpy()


def _versions(params):
    return tuple((p.data_ptr(), p._version) for p in params)


def set_dropout(model, p):
    """Set every dropout probability of a FAME model (or one of its towers) to p -- what the parity scripts do to the
    reference (`m.p = 0` on every nn.Dropout, `m.dropout = 0` on nn.MultiheadAttention, oracle/make_golden.py) --
    including the BERT config entries that stand in for HF's dropout modules here."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = p
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = p
        cfg = getattr(m, "config", None)
        if cfg is not None and hasattr(cfg, "hidden_dropout_prob"):
            cfg.hidden_dropout_prob = p
            cfg.attention_probs_dropout_prob = p
    return model


class BEHRTModel_Demo(nn.Module):
    """Demographic encoder (10_FAME.py:175-206): a 12-layer BERT over a length-1 sequence (token id 0) whose CLS
    state is added to the mean of four demographic embedding rows.  Same constructor, forward signature and
    state_dict keys (bert.*, age_embedding.weight, ...) as the reference class."""

    def __init__(self, num_ages, num_genders, num_ethnicities, num_insurances, hidden_size=768):
        super().__init__()
        vocab_size = num_ages + num_genders + num_ethnicities + num_insurances + 2
        self.bert = BertModelB200(vocab_size, hidden_size, 12, 12, 3072, 512)
        self.age_embedding = nn.Embedding(num_ages, hidden_size)
        self.gender_embedding = nn.Embedding(num_genders, hidden_size)
        self.ethnicity_embedding = nn.Embedding(num_ethnicities, hidden_size)
        self.insurance_embedding = nn.Embedding(num_insurances, hidden_size)

    def _tables(self):
        return [self.age_embedding.weight, self.gender_embedding.weight, self.ethnicity_embedding.weight,
                self.insurance_embedding.weight]

    def forward(self, input_ids, attention_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids):
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import demo_forward_train
            return demo_forward_train(self, input_ids, attention_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids)
        B, S = input_ids.shape
        h = self.bert.encode_f32(input_ids, attention_mask)           # f32 [B*S, hidden], fp32 residual stream
        with torch.no_grad():
            return ops.demo_add(h, S * h.shape[1], [age_ids, gender_ids, ethnicity_ids, insurance_ids],
                                [t.detach().float() for t in self._tables()])


class BEHRTModel_Lab(nn.Module):
    """Structured (lab / chart feature) encoder (10_FAME.py:208-224): one token per numeric feature,
    Linear(1, 768) + learned position, 2 post-norm TransformerEncoder layers (8 heads => head_dim 96, ff 2048,
    ReLU, eps 1e-5), mean over tokens.  torch's own nn.TransformerEncoder object is kept as the PARAMETER
    CONTAINER (identical keys and initialisation to the reference); every op runs in the sm_100a kernels."""

    def __init__(self, lab_token_count, hidden_size=768, nhead=8, num_layers=2):
        super().__init__()
        self.hidden_size = hidden_size
        self.nhead = nhead
        self.token_embedding = nn.Linear(1, hidden_size)
        self.pos_embedding = nn.Parameter(torch.randn(lab_token_count, hidden_size))
        encoder_layer = nn.TransformerEncoderLayer(d_model=hidden_size, nhead=nhead)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.transformer_encoder = nn.TransformerEncoder(encoder_layer, num_layers=num_layers)
        self._packed, self._packed_key = None, None

    def _pack(self):
        key = _versions(self.parameters())
        if self._packed is not None and self._packed_key == key:
            return self._packed
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        layers = []
        for l in self.transformer_encoder.layers:
            layers.append(dict(
                wqkv=bf(l.self_attn.in_proj_weight), bqkv=f32(l.self_attn.in_proj_bias),
                wo=bf(l.self_attn.out_proj.weight), bo=f32(l.self_attn.out_proj.bias),
                ln1=(f32(l.norm1.weight), f32(l.norm1.bias), l.norm1.eps),
                w1=bf(l.linear1.weight), b1=f32(l.linear1.bias), w2=bf(l.linear2.weight), b2=f32(l.linear2.bias),
                ln2=(f32(l.norm2.weight), f32(l.norm2.bias), l.norm2.eps)))
        self._packed = dict(layers=layers, w_tok=f32(self.token_embedding.weight[:, 0]),
                            b_tok=f32(self.token_embedding.bias), pos=f32(self.pos_embedding))
        self._packed_key = key
        return self._packed

    def forward(self, lab_features):
        if not lab_features.is_cuda:
            raise RuntimeError("BEHRTModel_Lab runs on a B200 only: move inputs to cuda (no CPU fallback)")
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import lab_forward_train
            return lab_forward_train(self, lab_features)
        with torch.no_grad():
            pk = self._pack()
            B, L = lab_features.shape
            H, nh = self.hidden_size, self.nhead
            x = ops.lab_embed(lab_features.float(), pk["w_tok"], pk["b_tok"], pk["pos"])
            for l in pk["layers"]:
                qkv = ops.gemm_bias_act(x, l["wqkv"], l["bqkv"])
                ctx = ops.attn_fwd(qkv, B, L, nh, H // nh)
                t = ops.gemm_bias_act(ctx, l["wo"], l["bo"], residual=x)
                x = ops.layernorm(t, l["ln1"][0], l["ln1"][1], l["ln1"][2], out=t)
                h = ops.gemm_bias_act(x, l["w1"], l["b1"], act=ops.ACT_RELU)
                t = ops.gemm_bias_act(h, l["w2"], l["b2"], residual=x)
                x = ops.layernorm(t, l["ln2"][0], l["ln2"][1], l["ln2"][2], out=t)
            return ops.seq_mean(x, B, L)


class MultimodalTransformer_EDDI_Sigmoid(nn.Module):
    """FAME fusion model (10_FAME.py:226-313): same constructor, forward signature, returned dict keys and
    state_dict layout.  Reproduced quirks: the 'mortality' EDDI weights gate all three
    outcomes (283-285); the modality classifiers are outside the loss; fusion_mlp[0] is evaluated once."""

    def __init__(self, text_embed_size, behrt_demo, behrt_lab, device, fusion_hidden=512, beta=1.0):
        super().__init__()
        if text_embed_size != 768 or fusion_hidden != 512:
            raise ValueError("the B200 fusion kernel is built for the reference's fixed sizes (768 -> 3x256 -> 512 -> 3)")
        self.behrt_demo = behrt_demo
        self.behrt_lab = behrt_lab
        self.device = device
        self.beta = beta
        self.demo_projector = nn.Sequential(nn.Linear(behrt_demo.bert.config.hidden_size, 256), nn.ReLU())
        self.lab_projector = nn.Sequential(nn.Linear(behrt_lab.hidden_size, 256), nn.ReLU())
        self.text_projector = nn.Sequential(nn.Linear(text_embed_size, 256), nn.ReLU())
        self.classifier_demo = nn.Linear(256, 3)
        self.classifier_lab = nn.Linear(256, 3)
        self.classifier_text = nn.Linear(256, 3)
        self.sig_weights = nn.Parameter(torch.randn(768))
        self.fusion_mlp = nn.Sequential(nn.Linear(768, fusion_hidden), nn.ReLU(), nn.Dropout(0.1),
                                        nn.Linear(fusion_hidden, 3))
        self._packed, self._packed_key = None, None

    def head_parameters(self):
        return [self.sig_weights, *self.demo_projector.parameters(), *self.lab_projector.parameters(),
                *self.text_projector.parameters(), *self.classifier_demo.parameters(),
                *self.classifier_lab.parameters(), *self.classifier_text.parameters(), *self.fusion_mlp.parameters()]

    def _pack_fusion(self):
        ps = self.head_parameters()
        key = _versions(ps)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous()
        projs = (self.demo_projector[0], self.lab_projector[0], self.text_projector[0])
        cls = (self.classifier_demo, self.classifier_lab, self.classifier_text)
        self._packed = dict(
            wp_t=torch.stack([f32(p.weight).t().contiguous() for p in projs]).contiguous(),   # [3,768,256]
            bp=torch.stack([f32(p.bias) for p in projs]).contiguous(),
            sig_w=f32(self.sig_weights),
            w3_t=f32(self.fusion_mlp[0].weight).t().contiguous(), b3=f32(self.fusion_mlp[0].bias),
            w4=f32(self.fusion_mlp[3].weight), b4=f32(self.fusion_mlp[3].bias),
            wc=torch.stack([f32(c.weight) for c in cls]).contiguous(), bc=torch.stack([f32(c.bias) for c in cls]).contiguous())
        self._packed_key = key
        return self._packed

    @staticmethod
    def modality_weights(old_eddi_weights):
        if old_eddi_weights is None:
            return 0.33, 0.33, 0.33
        return (old_eddi_weights.get("mortality", {"demo": 0.33})["demo"],
                old_eddi_weights.get("mortality", {"lab": 0.33})["lab"],
                old_eddi_weights.get("mortality", {"text": 0.33})["text"])

    def forward(self, demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids,
                lab_features, aggregated_text_embedding, beta=None, old_eddi_weights=None,
                return_modality_logits=False, return_gated_vector=False, return_intermediate=False):
        if beta is None:
            beta = self.beta
        w = self.modality_weights(old_eddi_weights)
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            from .train import fame_forward_train
            return fame_forward_train(self, (demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids,
                                             insurance_ids, lab_features, aggregated_text_embedding), w,
                                      return_modality_logits, return_gated_vector, return_intermediate)
        with torch.no_grad():
            demo = self.behrt_demo(demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids)
            lab = self.behrt_lab(lab_features)
            o = ops.fusion_fwd((demo, lab, aggregated_text_embedding.float()), self._pack_fusion(), w,
                               want_mod_logits=return_modality_logits,
                               want_intermediates=return_gated_vector or return_intermediate)
        outputs = {"fused_logits": o["logits"],
                   "dynamic_weights": {"demo": w[0], "lab": w[1], "text": w[2]},
                   "sigmoid_weights": o["sig"]}
        if return_modality_logits:
            outputs["modality_logits"] = {"demo": o["mod_logits"][0], "lab": o["mod_logits"][1], "text": o["mod_logits"][2]}
        if return_gated_vector:
            outputs["gated_vector"] = o["gated"]
        if return_intermediate:
            outputs["fusion_pre_relu"] = o["pre_relu"]
        return outputs

"""BERT encoder (HF ``BertModel`` parameter layout) executed by the sm_100a kernel library.

The module tree exists to reproduce the reference's ``state_dict`` keys exactly (SURVEY.md appendix A.1:
``embeddings.word_embeddings.weight``, ``encoder.layer.{i}.attention.self.query.weight`` ...), so checkpoints
round-trip with ``transformers.BertModel`` / the reference modules.  torch.nn.Linear / LayerNorm / Embedding are
used as PARAMETER CONTAINERS only; all arithmetic goes through fairmultimodal_b200.ops (C ABI -> CUDA).

Forward (inference, what 10_FAME.py:139-142 and :199 need): bf16 activations, fp32 accumulation, per layer
    qkv = GEMM(x, [Wq;Wk;Wv]) -> fused attention -> GEMM(+bias +residual) -> LN -> GEMM(+bias, GELU-erf)
    -> GEMM(+bias +residual) -> LN          (HF modeling_bert.py:179-206, 294-298, 339-342, 352-356)
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from . import ops


class _SelfAttention(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)


class _DenseLN(nn.Module):
    def __init__(self, fin, fout, eps):
        super().__init__()
        self.dense = nn.Linear(fin, fout)
        self.LayerNorm = nn.LayerNorm(fout, eps=eps)


class _Dense(nn.Module):
    def __init__(self, fin, fout):
        super().__init__()
        self.dense = nn.Linear(fin, fout)


class _Attention(nn.Module):
    def __init__(self, h, eps):
        super().__init__()
        self.self = _SelfAttention(h)
        self.output = _DenseLN(h, h, eps)


class _Layer(nn.Module):
    def __init__(self, h, inter, eps):
        super().__init__()
        self.attention = _Attention(h, eps)
        self.intermediate = _Dense(h, inter)
        self.output = _DenseLN(inter, h, eps)


class _Encoder(nn.Module):
    def __init__(self, h, inter, layers, eps):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(h, inter, eps) for _ in range(layers)])


class _Embeddings(nn.Module):
    def __init__(self, vocab, h, max_pos, eps):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, h, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, h)
        self.token_type_embeddings = nn.Embedding(2, h)
        self.LayerNorm = nn.LayerNorm(h, eps=eps)


class BertModelB200(nn.Module):
    """Same parameters / state_dict keys as ``transformers.BertModel(BertConfig(...))``."""

    def __init__(self, vocab_size, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, max_position_embeddings=512, layer_norm_eps=1e-12):
        super().__init__()
        self.config = types.SimpleNamespace(
            vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
            num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
            max_position_embeddings=max_position_embeddings, layer_norm_eps=layer_norm_eps,
            hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)      # BertConfig defaults (training only)
        self.embeddings = _Embeddings(vocab_size, hidden_size, max_position_embeddings, layer_norm_eps)
        self.encoder = _Encoder(hidden_size, intermediate_size, num_hidden_layers, layer_norm_eps)
        self.pooler = _Dense(hidden_size, hidden_size)       # computed-but-unused in the reference; kept for keys
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.normal_(p, std=0.02)
        with torch.no_grad():
            self.embeddings.word_embeddings.weight[0].zero_()
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    m.bias.zero_()
        self._packed = None
        self._packed_key = None

    @classmethod
    def from_hf(cls, hf_model):
        """Build from a ``transformers.BertModel`` (the object the reference passes to BioClinicalBERT_FT)."""
        c = hf_model.config
        m = cls(c.vocab_size, c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.intermediate_size,
                c.max_position_embeddings, c.layer_norm_eps)
        m.load_state_dict(hf_model.state_dict(), strict=True)
        return m

    # ------------------------------------------------------------------------------------------ packing
    def _pack(self):
        """bf16 copies of the GEMM weights ([Wq;Wk;Wv] concatenated), rebuilt when any parameter changes."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed_key == key:
            return self._packed
        layers = []
        with torch.no_grad():
            for l in self.encoder.layer:
                s = l.attention.self
                layers.append(dict(
                    wqkv=torch.cat([s.query.weight, s.key.weight, s.value.weight]).to(torch.bfloat16).contiguous(),
                    bqkv=torch.cat([s.query.bias, s.key.bias, s.value.bias]).float().contiguous(),
                    wo=l.attention.output.dense.weight.to(torch.bfloat16).contiguous(),
                    bo=l.attention.output.dense.bias.float().contiguous(),
                    ln1=(l.attention.output.LayerNorm.weight.float().contiguous(),
                         l.attention.output.LayerNorm.bias.float().contiguous()),
                    w1=l.intermediate.dense.weight.to(torch.bfloat16).contiguous(),
                    b1=l.intermediate.dense.bias.float().contiguous(),
                    w2=l.output.dense.weight.to(torch.bfloat16).contiguous(),
                    b2=l.output.dense.bias.float().contiguous(),
                    ln2=(l.output.LayerNorm.weight.float().contiguous(), l.output.LayerNorm.bias.float().contiguous()),
                ))
            e = self.embeddings
            emb = dict(word=e.word_embeddings.weight.float().contiguous(),
                       pos=e.position_embeddings.weight.float().contiguous(),
                       type0=e.token_type_embeddings.weight[0].float().contiguous(),
                       g=e.LayerNorm.weight.float().contiguous(), b=e.LayerNorm.bias.float().contiguous())
        self._packed, self._packed_key = dict(layers=layers, emb=emb), key
        return self._packed

    # ------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def encode_f32(self, input_ids, attention_mask=None):
        """Last hidden state as f32 [batch*seq, hidden], computed with an fp32 residual stream: GEMM operands are
        bf16 (tensor cores), but every residual add, LayerNorm input and LayerNorm output stays in fp32.  Meant for
        the small-batch towers (the demographic encoder runs M = batch rows) where the extra bytes are free and the
        bf16 rounding of the residual stream would otherwise dominate the error of the final logits."""
        if not input_ids.is_cuda:
            raise RuntimeError("BertModelB200 runs on a B200 only: move inputs to cuda (no CPU fallback)")
        c = self.config
        B, S = input_ids.shape
        pk = self._pack()
        eps = c.layer_norm_eps
        H, nh = c.hidden_size, c.num_attention_heads
        mask = (attention_mask != 0).to(torch.uint8).contiguous() if attention_mask is not None else None
        e = pk["emb"]
        x32 = torch.empty((B * S, H), device=input_ids.device, dtype=torch.float32)
        xb = ops.bert_embed(input_ids.to(torch.int64), e["word"], e["pos"], e["type0"], e["g"], e["b"], eps, S,
                            out_f32=x32)
        for l in pk["layers"]:
            if S == 1:
                ctx = ops.gemm_bias_act(xb, l["wqkv"][2 * H:], l["bqkv"][2 * H:])
            else:
                qkv = ops.gemm_bias_act(xb, l["wqkv"], l["bqkv"])
                ctx = ops.attn_fwd(qkv, B, S, nh, H // nh, key_mask=mask)
            t = ops.gemm_bias_act(ctx, l["wo"], l["bo"], residual=x32, out_dtype=torch.float32)
            xb, x32 = ops.layernorm(t, l["ln1"][0], l["ln1"][1], eps, want_f32=True)
            h = ops.gemm_bias_act(xb, l["w1"], l["b1"], act=ops.ACT_GELU_ERF)
            t = ops.gemm_bias_act(h, l["w2"], l["b2"], residual=x32, out_dtype=torch.float32)
            xb, x32 = ops.layernorm(t, l["ln2"][0], l["ln2"][1], eps, want_f32=True)
        return x32

    @torch.no_grad()
    def encode(self, input_ids, attention_mask=None, cls_only=False):
        """Last hidden state as bf16 [batch*seq, hidden] (row-major, sequence-major); with cls_only the hidden state
        of token 0 of every sequence, bf16 [batch, hidden] -- all the reference reads from the note encoder
        (`last_hidden_state[:, 0, :]`, 10_FAME.py:141).  In that mode the LAST layer computes keys / values for every
        token but the query, the attention, the attention output, both LayerNorms and the feed-forward block for the
        CLS rows only (the other rows of the last layer feed nothing): same numbers, 1/12 less work."""
        if not input_ids.is_cuda:
            raise RuntimeError("BertModelB200 runs on a B200 only: move inputs to cuda (no CPU fallback)")
        c = self.config
        B, S = input_ids.shape
        if S > c.max_position_embeddings:
            raise ValueError(f"sequence length {S} > max_position_embeddings {c.max_position_embeddings}")
        pk = self._pack()
        eps = c.layer_norm_eps
        H, nh = c.hidden_size, c.num_attention_heads
        mask = kv_len = None
        if attention_mask is not None:
            mask = (attention_mask != 0).to(torch.uint8).contiguous()
            if S > 128:
                kv_len = ops.mask_kv_len(mask)      # once per batch: the 12 layers skip all-padding key blocks
        e = pk["emb"]
        x = ops.bert_embed(input_ids.to(torch.int64), e["word"], e["pos"], e["type0"], e["g"], e["b"], eps, S)
        n_layers = len(pk["layers"])
        for li, l in enumerate(pk["layers"]):
            if cls_only and li == n_layers - 1 and S > 1 and H // nh == 64:
                return self._last_layer_cls(x, l, B, S, mask, eps)
            if S == 1:
                # one key per sequence: softmax == 1 exactly, so the context is the value projection itself and
                # Q / K never influence the output (the demographic encoder's shape, 10_FAME.py:199, 715-716)
                ctx = ops.gemm_bias_act(x, l["wqkv"][2 * H:], l["bqkv"][2 * H:])
            else:
                qkv = ops.gemm_bias_act(x, l["wqkv"], l["bqkv"])
                ctx = ops.attn_fwd(qkv, B, S, nh, H // nh, key_mask=mask, kv_len=kv_len)
            # the residual of the short-K attention-output projection is added by the LayerNorm kernel, not by the
            # GEMM epilogue (whose scattered residual loads were that GEMM's critical path); same rounding either way
            t = ops.gemm_bias_act(ctx, l["wo"], l["bo"])
            x = ops.layernorm(t, l["ln1"][0], l["ln1"][1], eps, out=t, residual=x)
            h = ops.gemm_bias_act(x, l["w1"], l["b1"], act=ops.ACT_GELU_ERF)
            t = ops.gemm_bias_act(h, l["w2"], l["b2"], residual=x)
            x = ops.layernorm(t, l["ln2"][0], l["ln2"][1], eps, out=t)
        return x.view(B, S, H)[:, 0, :] if cls_only else x

    def _last_layer_cls(self, x, l, B, S, mask, eps):
        """Last encoder layer for the CLS rows only: K / V projection of all tokens, then one query per sequence."""
        c = self.config
        H, nh = c.hidden_size, c.num_attention_heads
        x_cls = x.view(B, S, H)[:, 0, :]                                     # [B, H] view, row stride S * H
        kv = ops.gemm_bias_act(x, l["wqkv"][H:], l["bqkv"][H:])              # [B*S, 2H]: keys | values
        q = ops.gemm_bias_act(x_cls, l["wqkv"][:H], l["bqkv"][:H])           # [B, H]
        ctx = ops.attn_cls(q, kv, B, S, nh, H // nh, 0, H, key_mask=mask)
        t = ops.gemm_bias_act(ctx, l["wo"], l["bo"])
        y = ops.layernorm(t, l["ln1"][0], l["ln1"][1], eps, out=t, residual=x_cls)
        h = ops.gemm_bias_act(y, l["w1"], l["b1"], act=ops.ACT_GELU_ERF)
        t = ops.gemm_bias_act(h, l["w2"], l["b2"], residual=y)
        return ops.layernorm(t, l["ln2"][0], l["ln2"][1], eps, out=t)

    def forward(self, input_ids=None, attention_mask=None):
        """HF-style call: returns an object with ``last_hidden_state`` [batch, seq, hidden] (float32)."""
        B, S = input_ids.shape
        x = self.encode(input_ids, attention_mask)
        return types.SimpleNamespace(last_hidden_state=x.view(B, S, -1).float())

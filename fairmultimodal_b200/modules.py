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

    @classmethod
    def from_state_dict(cls, sd, prefix="BioBert.", **cfg):
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        vocab, hidden = sd["embeddings.word_embeddings.weight"].shape
        cfg.setdefault("num_hidden_layers", 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layer.")))
        bert = BertModelB200(vocab, hidden, **cfg)
        bert.load_state_dict(sd, strict=False)   # a pooler may be absent from hand-built dicts; it is unused
        return cls(bert)

    def encode_chunks(self, input_ids, attention_mask):
        """bf16 last hidden state [chunks*seq, hidden] (kept on device; CLS rows are every seq-th row)."""
        return self.BioBert.encode(input_ids, attention_mask)

    def forward(self, input_ids, attention_mask):
        C, S = input_ids.shape
        h = self.BioBert.encode(input_ids, attention_mask)
        return h.view(C, S, -1)[:, 0, :].float()


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

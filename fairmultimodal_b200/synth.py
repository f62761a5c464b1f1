"""Synthetic MIMIC-shaped cohort and deterministic random-init weights.

No real data or pretrained checkpoint is reachable (credentialed MIMIC-III, no network), so benchmarks and parity
tests use a seeded synthetic cohort with the real cohort's shape (33 721 patients, L = 542 lab tokens, label
prevalences from FinalCode/New/01_Data.log:55-57 of the reference) and weights drawn from a numpy Generator
keyed by (seed, parameter name), so that the build container, the GPU box and the golden-vector generator all
see bit-identical tensors without shipping a 400 MB state_dict.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np

LABEL_PREVALENCE = (0.101, 0.380, 0.900)           # mortality, LOS > 7 d, mechanical ventilation
AGE_P = (0.05, 0.15, 0.33, 0.45, 0.02)             # codes 0..4 (4 = "Other", never an EDDI group: 10_FAME.py:353)
GENDER_P = (0.44, 0.56)
ETH_P = (0.03, 0.10, 0.04, 0.13, 0.70)
INS_P = (0.03, 0.10, 0.55, 0.31, 0.01)             # code 5 never occurs -> exercises the empty-group skip
VOCAB = 28996                                      # Bio_ClinicalBERT (BERT-base cased)
CLS_ID, SEP_ID, PAD_ID = 101, 102, 0


def make_cohort(patients, lab_tokens=542, chunks="fixed4", seq_len=512, seed=1234, with_tokens=True):
    """Return a dict of numpy arrays: the 9 per-patient tensors of the reference TensorDataset
    (10_FAME.py:744-747) except the text embedding, plus the tokenised note chunks in CSR form."""
    rng = np.random.default_rng(seed)
    P = patients
    c = {}
    c["labels"] = (rng.random((P, 3)) < np.array(LABEL_PREVALENCE)).astype(np.float32)
    c["age_ids"] = rng.choice(5, P, p=AGE_P).astype(np.int64)
    c["gender_ids"] = rng.choice(2, P, p=GENDER_P).astype(np.int64)
    c["ethnicity_ids"] = rng.choice(5, P, p=ETH_P).astype(np.int64)
    c["insurance_ids"] = rng.choice(5, P, p=INS_P).astype(np.int64)
    lab = rng.standard_normal((P, lab_tokens)).astype(np.float32)
    const = rng.standard_normal(lab_tokens).astype(np.float32)
    lab = np.where(rng.random((P, lab_tokens)) < 0.6, const[None, :], lab)   # fillna(0) + z-score look-alike
    c["lab_features"] = lab.astype(np.float32)
    c["demo_dummy_ids"] = np.zeros((P, 1), dtype=np.int64)                    # 10_FAME.py:715
    c["demo_attn_mask"] = np.ones((P, 1), dtype=np.int64)                     # 10_FAME.py:716
    if chunks == "fixed4":
        n = np.full(P, 4, dtype=np.int64)
    elif chunks == "u1_16":
        n = rng.integers(1, 17, P)
    elif chunks == "u0_4":                                                    # includes note-less patients
        n = rng.integers(0, 5, P)
    else:
        n = np.full(P, int(chunks), dtype=np.int64)
    offsets = np.zeros(P + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(n)
    c["chunk_offsets"] = offsets
    C = int(offsets[-1])
    if with_tokens:
        ids = rng.integers(1000, VOCAB, (C, seq_len)).astype(np.int64)
        valid = np.full(C, seq_len, dtype=np.int64)
        last = offsets[1:][n > 0] - 1                                          # last chunk of each patient
        valid[last] = rng.integers(16, seq_len + 1, len(last))
        pos = np.arange(seq_len)[None, :]
        ids[:, 0] = CLS_ID
        ids[np.arange(C), valid - 1] = SEP_ID
        ids = np.where(pos < valid[:, None], ids, PAD_ID)
        c["input_ids"] = ids
        c["attention_mask"] = (pos < valid[:, None]).astype(np.int64)
        c["valid_len"] = valid
    return c


def pos_weight(labels):
    """N / (2 n_pos) per task (compute_class_weights, 10_FAME.py:48-52, 756-759)."""
    n_pos = np.maximum(labels.sum(axis=0), 1.0)
    return (labels.shape[0] / (2.0 * n_pos)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ weights
def bert_shapes(prefix, vocab, hidden=768, layers=12, inter=3072, max_pos=512, pooler=True):
    s = OrderedDict()
    e = prefix + "embeddings."
    s[e + "word_embeddings.weight"] = (vocab, hidden)
    s[e + "position_embeddings.weight"] = (max_pos, hidden)
    s[e + "token_type_embeddings.weight"] = (2, hidden)
    s[e + "LayerNorm.weight"] = (hidden,)
    s[e + "LayerNorm.bias"] = (hidden,)
    for i in range(layers):
        p = f"{prefix}encoder.layer.{i}."
        for n in ("query", "key", "value"):
            s[p + f"attention.self.{n}.weight"] = (hidden, hidden)
            s[p + f"attention.self.{n}.bias"] = (hidden,)
        s[p + "attention.output.dense.weight"] = (hidden, hidden)
        s[p + "attention.output.dense.bias"] = (hidden,)
        s[p + "attention.output.LayerNorm.weight"] = (hidden,)
        s[p + "attention.output.LayerNorm.bias"] = (hidden,)
        s[p + "intermediate.dense.weight"] = (inter, hidden)
        s[p + "intermediate.dense.bias"] = (inter,)
        s[p + "output.dense.weight"] = (hidden, inter)
        s[p + "output.dense.bias"] = (hidden,)
        s[p + "output.LayerNorm.weight"] = (hidden,)
        s[p + "output.LayerNorm.bias"] = (hidden,)
    if pooler:
        s[prefix + "pooler.dense.weight"] = (hidden, hidden)
        s[prefix + "pooler.dense.bias"] = (hidden,)
    return s


def fame_shapes(num_ages=5, num_genders=2, num_eth=5, num_ins=5, lab_tokens=542, hidden=768):
    """state_dict layout of MultimodalTransformer_EDDI_Sigmoid, in module order."""
    s = OrderedDict()
    s["sig_weights"] = (768,)
    s.update(bert_shapes("behrt_demo.bert.", num_ages + num_genders + num_eth + num_ins + 2, hidden))
    for n, k in (("age", num_ages), ("gender", num_genders), ("ethnicity", num_eth), ("insurance", num_ins)):
        s[f"behrt_demo.{n}_embedding.weight"] = (k, hidden)
    s["behrt_lab.pos_embedding"] = (lab_tokens, hidden)
    s["behrt_lab.token_embedding.weight"] = (hidden, 1)
    s["behrt_lab.token_embedding.bias"] = (hidden,)
    for i in range(2):
        p = f"behrt_lab.transformer_encoder.layers.{i}."
        s[p + "self_attn.in_proj_weight"] = (3 * hidden, hidden)
        s[p + "self_attn.in_proj_bias"] = (3 * hidden,)
        s[p + "self_attn.out_proj.weight"] = (hidden, hidden)
        s[p + "self_attn.out_proj.bias"] = (hidden,)
        s[p + "linear1.weight"] = (2048, hidden)
        s[p + "linear1.bias"] = (2048,)
        s[p + "linear2.weight"] = (hidden, 2048)
        s[p + "linear2.bias"] = (hidden,)
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = (hidden,)
            s[p + n + ".bias"] = (hidden,)
    for n in ("demo", "lab", "text"):
        s[f"{n}_projector.0.weight"] = (256, hidden)
        s[f"{n}_projector.0.bias"] = (256,)
    for n in ("demo", "lab", "text"):
        s[f"classifier_{n}.weight"] = (3, 256)
        s[f"classifier_{n}.bias"] = (3,)
    s["fusion_mlp.0.weight"] = (512, 768)
    s["fusion_mlp.0.bias"] = (512,)
    s["fusion_mlp.3.weight"] = (3, 512)
    s["fusion_mlp.3.bias"] = (3,)
    return s


def behrt_combined_shapes(lab_tokens=542, hidden=768):
    """state_dict layout of BEHRTModel_Combined (01_BEHRT.py:112-131): the lab tower under `lab_model.` + the head."""
    full = fame_shapes(lab_tokens=lab_tokens, hidden=hidden)
    out = OrderedDict()
    for k, shp in full.items():
        if k.startswith("behrt_lab."):
            out["lab_model." + k[len("behrt_lab."):]] = shp
    out["fusion_fc.weight"], out["fusion_fc.bias"] = (hidden, hidden), (hidden,)
    for h in ("classifier_mort", "classifier_los", "classifier_mech"):
        out[h + ".weight"], out[h + ".bias"] = (1, hidden), (1,)
    return out


def sigmoid_fusion_shapes(lab_tokens=542, hidden=768):
    """state_dict layout of 09_multimodal_sigmoid_fusion.py's MultimodalTransformer (09:162-195): the FAME towers under
    `BEHRT.` / `behrt_lab.` + projectors, three 256-wide sigmoid gates, aggregate_projector and classifier."""
    full = fame_shapes(lab_tokens=lab_tokens, hidden=hidden)
    out = OrderedDict()
    out["sig_weights_demo"] = out["sig_weights_lab"] = out["sig_weights_text"] = (256,)   # own parameters come first
    for k, shp in full.items():
        if k.startswith("behrt_demo."):
            out["BEHRT." + k[len("behrt_demo."):]] = shp
    for k, shp in full.items():
        if k.startswith("behrt_lab."):
            out[k] = shp
    for m in ("demo", "lab", "text"):
        out[f"{m}_projector.0.weight"], out[f"{m}_projector.0.bias"] = (256, hidden), (256,)
    out["aggregate_projector.0.weight"], out["aggregate_projector.0.bias"] = (512, 768), (512,)
    out["classifier.0.weight"], out["classifier.0.bias"] = (512, 512), (512,)
    out["classifier.3.weight"], out["classifier.3.bias"] = (3, 512), (3,)
    return out


def eddi_fusion_shapes(lab_tokens=542, hidden=768, demo_layers=6):
    """state_dict layout of 08_multimodal_eddi_fusion.py's MultimodalTransformer (08:314-346): a SIX-layer demographic
    BERT (08:264-267: 6 layers, 6 heads, 128 positions) + the lab tower + three projectors + nine scalar heads
    classifier_{demo,lab,text}_{mort,los,mv}."""
    full = fame_shapes(lab_tokens=lab_tokens, hidden=hidden)
    out = OrderedDict()
    for k, shp in full.items():
        if k.startswith("behrt_demo.bert.encoder.layer.") and int(k.split(".")[4]) >= demo_layers:
            continue
        if k.startswith(("behrt_demo.", "behrt_lab.")):
            out[k] = (128, hidden) if k.endswith("position_embeddings.weight") else shp   # max_position_embeddings=128
    for m in ("demo", "lab", "text"):
        out[f"{m}_projector.0.weight"], out[f"{m}_projector.0.bias"] = (256, hidden), (256,)
    for o in ("mort", "los", "mv"):
        for m in ("demo", "lab", "text"):
            out[f"classifier_{m}_{o}.weight"], out[f"classifier_{m}_{o}.bias"] = (1, 256), (1,)
    return out


def average_fusion_shapes(num_diseases=10, num_ages=5, num_segments=2, num_adm=4, num_disch=6, num_genders=2, num_eth=5,
                          num_ins=5, hidden=768):
    """state_dict layout of 07_multimodal_average_fusion.py's MultimodalTransformer (07:156-238): BEHRT = a 12-layer BERT
    + SEVEN embedding tables (age, segment, admission location, discharge location, gender, ethnicity, insurance),
    ts_linear / text_linear 768 -> 256 and a 512 -> 512 -> 3 classifier."""
    out = OrderedDict()
    out.update(bert_shapes("BEHRT.bert.", num_diseases + num_ages + num_segments + num_adm + num_disch + 2, hidden))
    for n, k in (("age", num_ages), ("segment", num_segments), ("admission_loc", num_adm), ("discharge_loc", num_disch),
                 ("gender", num_genders), ("ethnicity", num_eth), ("insurance", num_ins)):
        out[f"BEHRT.{n}_embedding.weight"] = (k, hidden)
    out["ts_linear.weight"], out["ts_linear.bias"] = (256, hidden), (256,)
    out["text_linear.weight"], out["text_linear.bias"] = (256, hidden), (256,)
    out["classifier.0.weight"], out["classifier.0.bias"] = (512, 512), (512,)
    out["classifier.3.weight"], out["classifier.3.bias"] = (3, 512), (3,)
    return out


def synth_tensor(name, shape, seed):
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    x = rng.standard_normal(shape, dtype=np.float32)
    leaf = name.rsplit(".", 2)
    if name.endswith(("LayerNorm.weight", "norm1.weight", "norm2.weight")):
        return 1.0 + 0.05 * x
    if name.endswith(("sig_weights", "pos_embedding")) or "sig_weights_" in name:
        return x                                         # nn.Parameter(torch.randn(...)), 10_FAME.py:213,252
    if name.endswith("token_embedding.weight"):
        return 0.5 * x                                   # Linear(1, 768): default init is U(-1, 1)
    if name.endswith(("_projector.0.weight", "fusion_mlp.0.weight", "fusion_mlp.3.weight")) or "classifier_" in name:
        return 0.04 * x
    if name.endswith(".bias") or leaf[-1] == "bias":
        return 0.02 * x
    x = 0.02 * x                                         # HF initializer_range
    if name.endswith("word_embeddings.weight"):
        x[0] = 0.0                                       # padding_idx row (BertConfig.pad_token_id = 0)
    return x


def synth_state_dict(shapes, seed=0):
    """name -> float32 numpy array, deterministic in (seed, name) and independent of dict order."""
    return OrderedDict((k, synth_tensor(k, shp, seed)) for k, shp in shapes.items())

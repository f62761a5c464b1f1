"""Note-encoder path (BioClinicalBERT_FT + chunk->patient pooling) on the B200 against the CPU oracle and against
the golden vectors produced by the unmodified reference."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# bf16 activations through 12 post-LN layers against an fp32 reference: hidden-state error is bounded relative to
# the largest reference magnitude (north_star: logits rel 1e-2 under bf16, measured after the fusion head).
REL_TOL_HIDDEN = 3e-2


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def test_matches_reference_golden(golden_dir):
    from fairmultimodal_b200 import modules, synth
    g = np.load(os.path.join(golden_dir, "notes.npz"))
    sd = {k: torch.from_numpy(v) for k, v in
          synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), int(g["wseed"])).items()}
    model = modules.BioClinicalBERT_FT.from_state_dict(sd).cuda()
    cls = model(torch.from_numpy(g["input_ids"]).cuda(), torch.from_numpy(g["attention_mask"]).cuda())
    ref = torch.from_numpy(g["cls"])
    err = (cls.cpu() - ref).abs().max().item()
    assert err <= REL_TOL_HIDDEN * ref.abs().max().item(), err
    rel_l2 = ((cls.cpu() - ref).norm() / ref.norm()).item()
    assert rel_l2 <= 2e-2, rel_l2
    pooled = modules.pool_chunks(torch.from_numpy(g["cls"]).cuda(), torch.from_numpy(g["offsets"]).cuda())
    np.testing.assert_array_equal(pooled.cpu().numpy(), g["pooled"])          # indexing + mean bit-exact
    pooled2 = modules.pool_chunks(cls, torch.from_numpy(g["offsets"]).cuda()).cpu().numpy()
    assert np.abs(pooled2 - g["pooled"]).max() <= REL_TOL_HIDDEN * np.abs(g["pooled"]).max()


def test_matches_oracle_small_model():
    """2-layer encoder, ragged valid lengths, several sequence lengths -- oracle computed live on the CPU."""
    from fairmultimodal_b200 import modules, synth
    from oracle import fame_oracle as O
    shapes = synth.bert_shapes("BioBert.", synth.VOCAB, layers=2)
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 3).items()}
    model = modules.BioClinicalBERT_FT.from_state_dict(sd, num_hidden_layers=2).cuda()
    for seq in (128, 200, 512):
        co = synth.make_cohort(3, lab_tokens=4, chunks=2, seq_len=seq, seed=seq)
        ids, mask = torch.from_numpy(co["input_ids"]), torch.from_numpy(co["attention_mask"])
        cls = model(ids.cuda(), mask.cuda()).cpu()
        with torch.no_grad():
            ref = O.bert_encode(sd, "BioBert.", ids, mask, num_layers=2)[:, 0, :]
        assert (cls - ref).abs().max().item() <= REL_TOL_HIDDEN * ref.abs().max().item()


def test_apply_on_patient_notes_host_logic():
    """Same patient order, column-major chunk order, empty-note handling as 10_FAME.py:144-173."""
    import pandas as pd
    from fairmultimodal_b200 import modules, synth
    shapes = synth.bert_shapes("BioBert.", synth.VOCAB, layers=1)
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(shapes, 5).items()}
    model = modules.BioClinicalBERT_FT.from_state_dict(sd, num_hidden_layers=1).cuda()
    co = synth.make_cohort(2, lab_tokens=4, chunks=2, seq_len=64, seed=4)
    ids, mask = co["input_ids"], co["attention_mask"]

    class Tok:
        def encode_plus(self, text, **kw):
            j = int(text.split("#")[1])
            return {"input_ids": torch.from_numpy(ids[j:j + 1]), "attention_mask": torch.from_numpy(mask[j:j + 1])}

    df = pd.DataFrame({"subject_id": [30, 10, 30, 20], "c0": ["n#0", "n#1", "n#2", None], "c1": [None, "  ", "n#3", None]})
    out = modules.apply_bioclinicalbert_on_patient_notes(df, ["c0", "c1"], Tok(), model, "cuda", max_length=64)
    assert out.shape == (3, 768)
    cls = model(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()).cpu().numpy()
    np.testing.assert_allclose(out[0], cls[[0, 2, 3]].mean(0), rtol=0, atol=1e-6)   # patient 30: c0 rows then c1 rows
    np.testing.assert_allclose(out[1], cls[1], rtol=0, atol=1e-6)                   # patient 10
    assert not out[2].any()                                                         # patient 20: no notes -> zeros


def test_full_size_properties():
    """BASELINE-size batch (256 chunks x 512 tokens): size-independent properties instead of an oracle run --
    determinism, permutation equivariance over chunks, and padding invariance of the CLS vector."""
    from fairmultimodal_b200 import modules, synth
    sd = {k: torch.from_numpy(v) for k, v in
          synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB), 7).items()}
    model = modules.BioClinicalBERT_FT.from_state_dict(sd).cuda()
    co = synth.make_cohort(64, lab_tokens=4, chunks="fixed4", seq_len=512, seed=1234)
    ids, mask = torch.from_numpy(co["input_ids"]).cuda(), torch.from_numpy(co["attention_mask"]).cuda()
    a = model(ids, mask)
    b = model(ids, mask)
    assert torch.equal(a, b)                                                   # deterministic
    perm = torch.randperm(256, device="cuda")
    c = model(ids[perm], mask[perm])
    assert torch.equal(c, a[perm])                                             # chunks are independent
    # changing the ids under the padding must not change CLS
    ids2 = torch.where(mask == 0, torch.full_like(ids, 1234), ids)
    d = model(ids2, mask)
    assert torch.equal(d, a)
    assert torch.isfinite(a).all()


def test_cls_only_last_layer_equals_full_encoder():
    """The hot path computes the last layer for the CLS rows only (the reference reads last_hidden_state[:, 0, :],
    10_FAME.py:141).  It must agree with the CLS rows of the all-token encoder: identical up to the last layer, and
    there the only arithmetic difference is the attention (fp32 probabilities on the one-query kernel, bf16 P on the
    tensor-core kernel)."""
    from fairmultimodal_b200 import modules, synth
    sd = {k: torch.from_numpy(v) for k, v in
          synth.synth_state_dict(synth.bert_shapes("BioBert.", synth.VOCAB, layers=3), 11).items()}
    model = modules.BioClinicalBERT_FT.from_state_dict(sd, num_hidden_layers=3).cuda()
    for seq in (512, 200):
        co = synth.make_cohort(5, lab_tokens=4, chunks=4, seq_len=seq, seed=seq + 1)
        ids, mask = torch.from_numpy(co["input_ids"]).cuda(), torch.from_numpy(co["attention_mask"]).cuda()
        full = model.encode_chunks(ids, mask).view(ids.shape[0], seq, -1)[:, 0, :].float()
        cls = model.encode_cls(ids, mask).float()
        assert cls.shape == full.shape
        assert (cls - full).abs().max().item() <= 1e-2 * full.abs().max().item()

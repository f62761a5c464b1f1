"""Disk format and host-side tensor preparation either side of the hot path: the synthetic-cohort CSV
writer and load_fame_cohort (restatement of 10_FAME.py:610-723) against the UNMODIFIED reference front half executed
here (when /root/reference is present) and against a committed fixture of its outputs."""
import os

import numpy as np
import pytest
import torch

from fairmultimodal_b200 import dataprep

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "dataprep.npz")
ARGS = dict(patients=64, lab_cols=12, chart_cols=6, max_chunks=4, seed=3)
T_KEYS = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "labels")


def test_writer_covers_the_awkward_cases(tmp_path):
    sp, up = dataprep.write_synthetic_csvs(str(tmp_path), **ARGS)
    import pandas as pd
    st, un = pd.read_csv(sp), pd.read_csv(up)
    assert {"subject_id", "hadm_id", "age", "GENDER", "ETHNICITY", "INSURANCE", *dataprep.OUTCOMES} <= set(st.columns)
    assert [c for c in un.columns if c.startswith("note_chunk_")] == [f"note_chunk_{i}" for i in range(1, 5)]
    assert set(st.subject_id) != set(un.subject_id)                      # one patient missing on each side
    assert st.filter(like="lab_t").isna().any().any()                    # NaN lab values
    assert un.filter(like="note_chunk_").isna().all(axis=1).any()        # note-less patients
    out = dataprep.load_fame_cohort(sp, up)
    n = len(out["df_filtered"])
    assert 0 < n < 63 and out["lab_features"].shape == (n, 19)           # 12 + 6 + the constant column
    assert torch.isfinite(out["lab_features"]).all()
    const = out["lab_feature_columns"].index("lab_t_const")
    assert out["lab_features"][:, const].abs().max() == 0                # zero-variance column -> zeros
    assert out["age_ids"].max() <= 4 and out["ethnicity_ids"].max() <= 4 and out["insurance_ids"].max() <= 5
    w = dataprep.compute_class_weights(out["df_filtered"], "short_term_mortality")
    pos = int(out["labels"][:, 0].sum())
    assert abs(w[1] - n / (2 * pos)) < 1e-12


def _compare(out, ref):
    for k in T_KEYS:
        assert np.array_equal(out[k].numpy(), np.asarray(ref[k])), k           # integer codes / labels: bit-exact
    assert list(out["note_columns"]) == [str(x) for x in ref["note_columns"]]
    assert list(out["lab_feature_columns"]) == [str(x) for x in ref["lab_feature_columns"]]
    assert np.array_equal(out["df_filtered"]["subject_id"].to_numpy(), np.asarray(ref["subject_id"]))
    assert np.array_equal(out["lab_features"].numpy(), np.asarray(ref["lab_features_t"]))  # same numpy calls: bit-exact


def test_loader_matches_committed_reference_outputs(tmp_path):
    sp, up = dataprep.write_synthetic_csvs(str(tmp_path), **ARGS)
    _compare(dataprep.load_fame_cohort(sp, up), np.load(GOLDEN, allow_pickle=False))


def test_loader_matches_unmodified_reference_front_half(tmp_path):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    from oracle import ref_frontend
    for seed, patients in ((3, 64), (11, 40)):
        d = tmp_path / f"c{seed}"
        sp, up = dataprep.write_synthetic_csvs(str(d), **{**ARGS, "seed": seed, "patients": patients})
        loc = ref_frontend.run_front_half(str(d))
        ref = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in loc.items() if k != "df_filtered"}
        ref["subject_id"] = loc["df_filtered"]["subject_id"].to_numpy()
        _compare(dataprep.load_fame_cohort(sp, up), ref)

"""Host side of the FAME pipeline, either side of the hot path: the on-disk cohort format that
10_FAME.py reads and the tensors its front half (run_experiment, 10_FAME.py:610-723) builds from it.

    write_synthetic_csvs   final_structured_common.csv / final_unstructured_common.csv in the column layout 10_FAME.py
                           consumes (the products of 00_data.py:385-386, 431-439, 494-536 after its renames): subject_id,
                           hadm_id, age, GENDER, ETHNICITY, INSURANCE, the three outcomes, lab_t* / chartevents_t* feature
                           columns; note_chunk_1..k columns of whitespace-token text, NaN where a patient has fewer chunks
    load_fame_cohort       restatement of 10_FAME.py:610-723: merge, note filter, age buckets, ethnicity / insurance
                           mapping, category codes, lab feature selection, fillna(0) + z-score, the per-patient tensors

Pure pandas / numpy (the reference does this on the host too); no device code.  tests/test_dataprep_cpu.py checks it
against the UNMODIFIED reference front half executed in the build container (its locals are harvested at the first
network call) and against a committed golden fixture.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

ETHNICITIES = ("WHITE", "BLACK", "HISPANIC", "ASIAN", "UNKNOWN/NOT SPECIFIED", "white ")   # raw spellings, incl. unmapped
INSURANCES = ("Medicare", "Private", "Medicaid", "Government", "Self Pay", "self pay")
OUTCOMES = ("short_term_mortality", "los_binary", "mechanical_ventilation")


def write_synthetic_csvs(out_dir, patients=64, lab_cols=12, chart_cols=6, max_chunks=4, words_per_chunk=40, seed=0,
                         missing_frac=0.3):
    """Write the two CSVs; returns (structured_path, unstructured_path).  Includes the awkward cases the reference's
    front half handles: patients without any note (dropped by its filter), a patient present in only one file (dropped
    by the inner merge), NaN lab values (fillna(0)), a constant lab column (std 0), ages outside every bucket and
    non-numeric ages ("Other"), unmapped / differently-cased ethnicity and insurance strings."""
    rng = np.random.default_rng(seed)
    P = patients
    sid = np.arange(1000, 1000 + P)
    hadm = sid * 10 + 7
    age = rng.integers(15, 95, P).astype(object)
    age[rng.random(P) < 0.05] = 300                       # MIMIC's shifted ages of > 89-year-olds -> "Other"
    if P > 3:
        age[3] = "unknown"
    st = pd.DataFrame({"subject_id": sid, "hadm_id": hadm, "age": age,
                       "GENDER": rng.choice(["M", "F"], P), "ETHNICITY": rng.choice(ETHNICITIES, P, p=(.55, .15, .08, .07, .1, .05)),
                       "INSURANCE": rng.choice(INSURANCES, P, p=(.5, .3, .1, .04, .03, .03))})
    for i, prev in zip(OUTCOMES, (0.15, 0.4, 0.8)):
        st[i] = (rng.random(P) < prev).astype(int)
    for j in range(lab_cols):
        v = rng.standard_normal(P) * (1 + j) + j
        v[rng.random(P) < missing_frac] = np.nan
        st[f"lab_t{j}"] = v
    for j in range(chart_cols):
        st[f"chartevents_t{j}"] = rng.standard_normal(P) * 3
    st["lab_t_const"] = 2.5                               # zero variance: (x - mean) / (0 + 1e-6) = 0
    notes = {}
    n_chunks = rng.integers(0, max_chunks + 1, P)         # 0 = no notes at all
    for c in range(max_chunks):
        col = []
        for p in range(P):
            if c < n_chunks[p]:
                col.append(" ".join(f"w{int(t)}" for t in rng.integers(0, 5000, int(rng.integers(3, words_per_chunk + 1)))))
            else:
                col.append(np.nan)
        notes[f"note_chunk_{c + 1}"] = col
    un = pd.DataFrame({"subject_id": sid, "hadm_id": hadm, **notes})
    if P > 5:
        un.loc[5, "note_chunk_1"] = "   "                 # whitespace only: not a valid note (10_FAME.py:632)
    for i in OUTCOMES:                                    # duplicated in the unstructured file; dropped by the reader
        un[i] = st[i]
    un["age"], un["GENDER"], un["ETHNICITY"], un["INSURANCE"] = st["age"], st["GENDER"], st["ETHNICITY"], st["INSURANCE"]
    if P > 8:
        st = st.drop(index=7).reset_index(drop=True)      # in the unstructured file only
        un = un.drop(index=8).reset_index(drop=True)      # in the structured file only
    os.makedirs(out_dir, exist_ok=True)
    sp, up = os.path.join(out_dir, "final_structured_common.csv"), os.path.join(out_dir, "final_unstructured_common.csv")
    st.to_csv(sp, index=False)
    un.to_csv(up, index=False)
    return sp, up


def _age_bucket(age):                                     # 10_FAME.py:644-658
    try:
        age = float(age)
    except Exception:
        return "Other"
    if 15 <= age <= 29:
        return "15-29"
    if 30 <= age <= 49:
        return "30-49"
    if 50 <= age <= 69:
        return "50-69"
    if 70 <= age <= 89:
        return "70-89"
    return "Other"


def _map_ethnicity(e):                                    # 10_FAME.py:662-670
    try:
        return {0: "White", 1: "Black", 2: "Hispanic", 3: "Asian"}.get(int(e), "Other")
    except Exception:
        e = str(e).strip().title()
        return {"White": "White", "Black": "Black", "Asian": "Asian", "Hispanic": "Hispanic"}.get(e, "Other")


def _map_insurance(i):                                    # 10_FAME.py:677-686
    try:
        return {0: "Government", 1: "Medicare", 2: "Medicaid", 3: "Private", 4: "Self Pay"}.get(int(i), "Other")
    except Exception:
        i = str(i).strip().title()
        return {"Government": "Government", "Medicare": "Medicare", "Medicaid": "Medicaid", "Private": "Private",
                "Self Pay": "Self Pay"}.get(i, "Other")


def load_fame_cohort(structured_csv="final_structured_common.csv", unstructured_csv="final_unstructured_common.csv"):
    """10_FAME.py:610-723.  Returns a dict: df_filtered, note_columns, lab_feature_columns and the tensors
    demo_dummy_ids, demo_attn_mask, age_ids, gender_ids, ethnicity_ids, insurance_ids (int64), lab_features,
    labels (float32), in the row order of df_filtered (= the order apply_bioclinicalbert_on_patient_notes keeps)."""
    structured = pd.read_csv(structured_csv, low_memory=False)
    unstructured = pd.read_csv(unstructured_csv, low_memory=False)
    unstructured = unstructured.drop(columns=list(OUTCOMES) + ["age", "GENDER", "ETHNICITY", "INSURANCE"], errors="ignore")
    merged = pd.merge(structured, unstructured, on=["subject_id", "hadm_id"], how="inner", suffixes=("_struct", "_unstruct"))
    if merged.empty:
        raise ValueError("Merged DataFrame is empty. Check your merge keys.")
    for c in OUTCOMES:
        merged[c] = merged[c].astype(int)
    note_columns = [c for c in merged.columns if c.startswith("note_")]
    notes = merged[note_columns]
    valid = np.zeros(len(merged), dtype=bool)
    for c in note_columns:                               # any non-null string with a non-blank character (10_FAME.py:630-634)
        col = notes[c]
        valid |= col.map(lambda v: isinstance(v, str) and bool(v.strip())).to_numpy(dtype=bool)
    df = merged[valid].copy()
    if "age" not in df.columns:
        if "Age" in df.columns:
            df.rename(columns={"Age": "age"}, inplace=True)
        else:
            df["age"] = 0
    df["age"] = df["age"].apply(_age_bucket).astype("category").cat.codes
    if "ETHNICITY" in df.columns:
        df["ETHNICITY"] = df["ETHNICITY"].apply(_map_ethnicity).astype("category").cat.codes
    else:
        df["ETHNICITY"] = 0
    if "INSURANCE" in df.columns:
        df["INSURANCE"] = df["INSURANCE"].apply(_map_insurance).astype("category").cat.codes
    else:
        df["INSURANCE"] = 0
    gender_col = "GENDER" if ("GENDER" in df.columns and df["GENDER"].dtype == object) else \
        ("GENDERS" if "GENDERS" in df.columns else "GENDER")
    df[gender_col] = df[gender_col].astype("category").cat.codes
    exclude = {"subject_id", "ROW_ID", "hadm_id", "ICUSTAY_ID", *OUTCOMES, "age", "GENDER", "GENDERS", "ETHNICITY", "INSURANCE"}
    lab_cols = [c for c in df.columns if c not in exclude and not c.startswith("note_")
                and pd.api.types.is_numeric_dtype(df[c])]
    df[lab_cols] = df[lab_cols].fillna(0)
    lab = df[lab_cols].values.astype(np.float32)
    lab = (lab - np.mean(lab, axis=0)) / (np.std(lab, axis=0) + 1e-6)
    n = len(df)
    return {
        "df_filtered": df, "note_columns": note_columns, "lab_feature_columns": lab_cols,
        "demo_dummy_ids": torch.zeros((n, 1), dtype=torch.long), "demo_attn_mask": torch.ones((n, 1), dtype=torch.long),
        "age_ids": torch.tensor(df["age"].values, dtype=torch.long),
        "gender_ids": torch.tensor(df["GENDER"].values, dtype=torch.long),
        "ethnicity_ids": torch.tensor(df["ETHNICITY"].values, dtype=torch.long),
        "insurance_ids": torch.tensor(df["INSURANCE"].values, dtype=torch.long),
        "lab_features": torch.tensor(lab, dtype=torch.float32),
        "labels": torch.tensor(df[list(OUTCOMES)].values.astype(np.float32), dtype=torch.float32),
    }


def compute_class_weights(df, label_column):
    """10_FAME.py:48-52: N / (count_c * n_classes) per class, indexed by class value ([1] = the positive weight)."""
    counts = df[label_column].value_counts().sort_index()
    return len(df) / (counts * len(counts))

"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; argument structs match the header; state_dict layouts match the reference's."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from fairmultimodal_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "fame_b200.h")).read()
    declared = set(re.findall(r"\b(fame_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fame_b200.h but not exported"
    assert declared == set(_lib.OP_TABLE) | set(_lib.PLAIN_SYMBOLS) | set(_lib.FLAT_OPS)


def test_struct_sizes_match_header(lib):
    # compile a tiny C program against the header and compare sizeof() with the ctypes mirrors
    import subprocess, tempfile
    names = dict(_lib.STRUCT_NAMES)
    header = open(os.path.join(ROOT, "include", "fame_b200.h")).read()
    assert set(re.findall(r"\}\s*(fame_[a-z0-9_]+_args)\s*;", header)) == set(names)
    src = '#include <stdio.h>\n#include "fame_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert ctypes.sizeof(names[n]) == int(sz), n


def test_no_device_is_an_error_not_a_fallback(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.fame_device_check() != 0
    assert b"" != lib.fame_strerror(-1)
    from fairmultimodal_b200 import ops
    with pytest.raises(_lib.FameError):
        ops.layernorm(torch.zeros(4, 768, dtype=torch.bfloat16), torch.ones(768), torch.zeros(768), 1e-5)


def test_note_encoder_state_dict_layout():
    from fairmultimodal_b200.bert import BertModelB200
    m = BertModelB200(64, num_hidden_layers=2)
    shapes = synth.bert_shapes("", 64, layers=2)
    sd = m.state_dict()
    assert set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)


def test_hf_bert_round_trip():
    from transformers import BertConfig, BertModel
    from fairmultimodal_b200.bert import BertModelB200
    hf = BertModel(BertConfig(vocab_size=50, num_hidden_layers=1))
    m = BertModelB200.from_hf(hf)
    hf.load_state_dict(m.state_dict(), strict=True)          # loads both ways


def test_synth_cohort_shapes():
    co = synth.make_cohort(10, lab_tokens=12, chunks="u1_16", seq_len=64, seed=0)
    C = int(co["chunk_offsets"][-1])
    assert co["input_ids"].shape == (C, 64) and (co["input_ids"][:, 0] == synth.CLS_ID).all()
    v = co["valid_len"]
    assert (co["input_ids"][range(C), v - 1] == synth.SEP_ID).all()
    assert ((co["attention_mask"].sum(1)) == v).all()
    assert co["insurance_ids"].max() <= 4 and co["age_ids"].max() <= 4


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs without a GPU and prints ONE JSON line with the keys the driver reads: the
    reference's CPU path of the training step (headline) with the note encoder as a sub-object, and `--config 2` the
    note encoder alone."""
    import json
    import subprocess
    import sys

    def run(*extra):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                              *extra], capture_output=True, text=True, timeout=900, check=True).stdout
        lines = [l for l in out.strip().splitlines() if l.startswith("{")]
        assert len(lines) == 1
        return json.loads(lines[0])

    d = run()
    assert d["impl"] == "reference" and d["unit"] == "patients/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["n_gpus"] == 1 and d["config"]["workload"].startswith("fame_train_step_32patients")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "patients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    n = d["note_encoder"]
    assert n["unit"] == "chunks/s" and n["value"] > 0 and n["cpu_baseline"]["kind"] in ("reference", "port")
    d = run("--config", "2", "--ref-chunks-per-step", "1")
    assert d["impl"] == "reference" and d["unit"] == "chunks/s" and d["value"] > 0
    assert d["config"]["workload"].startswith("note_encoder_fwd")
    assert d["e2e"] == {"value": d["value"], "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

"""TEST INFRASTRUCTURE ONLY: run the UNMODIFIED front half of the reference's run_experiment (10_FAME.py:606-723) on two
CSV files and harvest the tensors it builds.  The function reads its inputs from the working directory and then calls
the HF hub; the hub call (AutoTokenizer.from_pretrained, 10_FAME.py:726) is replaced by a sentinel exception and the
locals of the run_experiment frame are read from the traceback -- everything before that line ran exactly as written.
Build container only (needs /root/reference)."""
import contextlib
import io
import os

from oracle import ref_loader

KEYS = ("demo_dummy_ids", "demo_attn_mask", "age_ids", "gender_ids", "ethnicity_ids", "insurance_ids", "lab_features_t",
        "labels", "note_columns", "lab_feature_columns", "df_filtered")


class _Stop(Exception):
    pass


def run_front_half(csv_dir):
    ref = ref_loader.load()

    class _Tok:
        @staticmethod
        def from_pretrained(*a, **k):
            raise _Stop()

    saved, cwd = ref.AutoTokenizer, os.getcwd()
    ref.AutoTokenizer = _Tok
    try:
        os.chdir(csv_dir)
        with contextlib.redirect_stdout(io.StringIO()):
            ref.run_experiment({"batch_size": 8})
    except _Stop as e:
        tb = e.__traceback__
        while tb is not None and tb.tb_frame.f_code.co_name != "run_experiment":
            tb = tb.tb_next
        loc = tb.tb_frame.f_locals
        return {k: loc[k] for k in KEYS}
    finally:
        os.chdir(cwd)
        ref.AutoTokenizer = saved
    raise RuntimeError("the reference front half did not reach the tokenizer call")

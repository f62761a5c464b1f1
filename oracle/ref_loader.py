"""Load the UNMODIFIED reference script (FinalCode/New/Final/10_FAME.py) as a module.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container, not the GPU box).
Used by oracle/make_golden.py to generate the golden vectors under tests/golden/ and by the CPU tests that
pin oracle/fame_oracle.py against the real reference.  Nothing in the product package imports this.

The script imports three third-party modules that are absent here and unused by the hot-path code
(10_FAME.py:19-21); empty stand-ins are registered so the import succeeds.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("FAME_REFERENCE_ROOT", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "FinalCode", "New", "Final", "10_FAME.py")


def available() -> bool:
    return os.path.exists(REF_FILE)


def load():
    if "fame_ref" in sys.modules:
        return sys.modules["fame_ref"]
    if not available():
        raise FileNotFoundError(REF_FILE)
    for name in ("iterstrat", "iterstrat.ml_stratifiers", "matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["iterstrat.ml_stratifiers"].MultilabelStratifiedShuffleSplit = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("fame_ref", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fame_ref"] = mod
    spec.loader.exec_module(mod)
    return mod

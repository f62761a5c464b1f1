"""ctypes binding of ``libfame_b200.so`` (C ABI declared in ``include/fame_b200.h``).

There is deliberately no fallback: if the shared library is missing, or the device is not a B200
(compute capability 10.x), every op raises.  The argument structs below mirror the header field for field.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfame_b200.so")

FAME_OK = 0
ACT_NONE, ACT_GELU_ERF, ACT_RELU = 0, 1, 2
DT_BF16, DT_F32 = 0, 1


class FameError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("y", C.c_void_p), ("ldy", C.c_int64),
        ("y_dtype", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("act", C.c_int32),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("y", C.c_void_p), ("ldy", C.c_int64),
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("eps", C.c_float),
    ]


class BertEmbedArgs(C.Structure):
    _fields_ = [
        ("ids", C.c_void_p), ("word", C.c_void_p), ("pos", C.c_void_p), ("type0", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("y", C.c_void_p), ("err_flag", C.c_void_p),
        ("tokens", C.c_int32), ("seq_len", C.c_int32), ("hidden", C.c_int32), ("vocab", C.c_int32),
        ("eps", C.c_float),
    ]


class AttnFwdArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ld_qkv", C.c_int64),
        ("key_mask", C.c_void_p),
        ("ctx", C.c_void_p), ("ld_ctx", C.c_int64),
        ("batch", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float), ("algo", C.c_int32),
    ]


class SegmentMeanArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64), ("x_dtype", C.c_int32),
        ("offsets", C.c_void_p), ("out", C.c_void_p),
        ("patients", C.c_int32), ("cols", C.c_int32), ("mode", C.c_int32),
    ]


# name -> args struct for every `int fame_<op>(const args*, void* ws, size_t ws_bytes, stream)` entry point
OP_TABLE = {
    "fame_gemm_bias_act": GemmArgs,
    "fame_layernorm": LayerNormArgs,
    "fame_bert_embed": BertEmbedArgs,
    "fame_attn_fwd": AttnFwdArgs,
    "fame_segment_mean": SegmentMeanArgs,
}
PLAIN_SYMBOLS = ["fame_strerror", "fame_last_cuda_error", "fame_abi_version", "fame_device_check", "fame_sm_count"]

_lib = None


def load() -> C.CDLL:
    """Load the library once; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FameError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fairmultimodal_b200 has no CPU or eager fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.fame_strerror.restype = C.c_char_p
    lib.fame_strerror.argtypes = [C.c_int]
    for name in ("fame_last_cuda_error", "fame_abi_version", "fame_device_check", "fame_sm_count"):
        getattr(lib, name).restype = C.c_int
        getattr(lib, name).argtypes = []
    for name, struct in OP_TABLE.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(struct), C.c_void_p, C.c_size_t, C.c_void_p]
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != FAME_OK:
        lib = load()
        msg = lib.fame_strerror(code).decode()
        extra = f" (cuda error {lib.fame_last_cuda_error()})" if code == -6 else ""
        raise FameError(f"{what}: {msg}{extra}")


def call(name: str, args, stream: int, workspace: int = 0, workspace_bytes: int = 0) -> None:
    lib = load()
    check(getattr(lib, name)(C.byref(args), workspace, workspace_bytes, stream), name)

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


class DropoutCfg(C.Structure):
    _fields_ = [("step", C.c_void_p), ("seed", C.c_uint32), ("thresh16", C.c_uint32), ("group_shift", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64), ("residual_dtype", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int64),
        ("y_dtype", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("act", C.c_int32), ("drop", DropoutCfg),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64), ("x_dtype", C.c_int32),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("y", C.c_void_p), ("y_f32", C.c_void_p), ("ldy", C.c_int64), ("stats", C.c_void_p),
        ("rows", C.c_int32), ("cols", C.c_int32),
        ("eps", C.c_float), ("residual", C.c_void_p), ("ldr", C.c_int64),
    ]


class BertEmbedArgs(C.Structure):
    _fields_ = [
        ("ids", C.c_void_p), ("word", C.c_void_p), ("pos", C.c_void_p), ("type0", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("y", C.c_void_p), ("y_f32", C.c_void_p), ("sum_out", C.c_void_p),
        ("stats", C.c_void_p), ("err_flag", C.c_void_p),
        ("tokens", C.c_int32), ("seq_len", C.c_int32), ("hidden", C.c_int32), ("vocab", C.c_int32),
        ("eps", C.c_float),
    ]


class AttnFwdArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ld_qkv", C.c_int64),
        ("key_mask", C.c_void_p),
        ("ctx", C.c_void_p), ("ld_ctx", C.c_int64),
        ("batch", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float), ("algo", C.c_int32), ("lse", C.c_void_p), ("kv_len", C.c_void_p),
        ("drop", DropoutCfg),
    ]


class SegmentMeanArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64), ("x_dtype", C.c_int32),
        ("offsets", C.c_void_p), ("out", C.c_void_p),
        ("patients", C.c_int32), ("cols", C.c_int32), ("mode", C.c_int32),
    ]


class LabEmbedArgs(C.Structure):
    _fields_ = [
        ("lab", C.c_void_p), ("w_tok", C.c_void_p), ("b_tok", C.c_void_p), ("pos", C.c_void_p), ("y", C.c_void_p),
        ("batch", C.c_int32), ("L", C.c_int32), ("hidden", C.c_int32),
    ]


class SeqMeanArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out", C.c_void_p), ("batch", C.c_int32), ("L", C.c_int32), ("cols", C.c_int32)]


class DemoAddArgs(C.Structure):
    _fields_ = [
        ("cls", C.c_void_p), ("ld_cls", C.c_int64), ("cls_dtype", C.c_int32),
        ("ids", C.c_void_p * 4), ("table", C.c_void_p * 4), ("n_rows", C.c_int32 * 4),
        ("out", C.c_void_p), ("batch", C.c_int32), ("hidden", C.c_int32),
    ]


class EmbedMeanArgs(C.Structure):
    _fields_ = [
        ("cls", C.c_void_p), ("ld_cls", C.c_int64), ("cls_dtype", C.c_int32), ("n_tables", C.c_int32),
        ("ids", C.c_void_p * 8), ("table", C.c_void_p * 8), ("dtable", C.c_void_p * 8), ("n_rows", C.c_int32 * 8),
        ("out", C.c_void_p), ("dout", C.c_void_p), ("batch", C.c_int32), ("hidden", C.c_int32),
    ]


class FusionFwdArgs(C.Structure):
    _fields_ = [
        ("emb", C.c_void_p * 3), ("wp_t", C.c_void_p), ("bp", C.c_void_p), ("w_mod", C.c_float * 3),
        ("sig_w", C.c_void_p), ("w3_t", C.c_void_p), ("b3", C.c_void_p), ("w4", C.c_void_p), ("b4", C.c_void_p),
        ("wc", C.c_void_p), ("bc", C.c_void_p), ("proj", C.c_void_p), ("gated", C.c_void_p),
        ("pre_relu", C.c_void_p), ("logits", C.c_void_p), ("mod_logits", C.c_void_p), ("sig_out", C.c_void_p),
        ("B", C.c_int32), ("w_mod_dev", C.c_void_p),
    ]


class LossStatsArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("labels", C.c_void_p), ("attr", C.c_void_p * 3), ("pos_weight", C.c_void_p),
        ("stats", C.c_void_p), ("B", C.c_int32),
    ]


class LossFwdBwdArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("labels", C.c_void_p), ("attr", C.c_void_p * 3), ("pos_weight", C.c_void_p),
        ("stats", C.c_void_p), ("sig_w", C.c_void_p), ("n_sig", C.c_int32),
        ("lambda_edd", C.c_float), ("lambda_l1", C.c_float),
        ("dlogits", C.c_void_p), ("loss_out", C.c_void_p), ("B", C.c_int32),
    ]


class EvalCountsArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("ld", C.c_int64), ("labels", C.c_void_p), ("attr", C.c_void_p * 3),
        ("thr", C.c_double * 3), ("sweep", C.c_void_p), ("out", C.c_void_p),
        ("N", C.c_int32), ("logits_are_probs", C.c_int32),
    ]


class RankCountsArgs(C.Structure):
    _fields_ = [
        ("scores", C.c_void_p), ("y", C.c_void_p), ("N", C.c_int32), ("i0", C.c_int32), ("i1", C.c_int32),
        ("auroc2", C.c_void_p), ("ap_sum", C.c_void_p), ("npos_nneg", C.c_void_p),
    ]


class SigmoidProbsArgs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("ld", C.c_int64), ("labels", C.c_void_p), ("probs", C.c_void_p),
        ("y8", C.c_void_p), ("N", C.c_int32),
    ]


class GemmOperand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("stride_b0", C.c_int64), ("stride_b1", C.c_int64),
                ("mn_major", C.c_int32)]


class GemmExArgs(C.Structure):
    _fields_ = [
        ("a", GemmOperand), ("b", GemmOperand), ("bias", C.c_void_p),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64), ("aux_stride_b0", C.c_int64), ("aux_stride_b1", C.c_int64),
        ("aux_mode", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int64), ("y_stride_b0", C.c_int64), ("y_stride_b1", C.c_int64),
        ("y_dtype", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("nb0", C.c_int32), ("nb1", C.c_int32),
        ("act", C.c_int32), ("alpha", C.c_float), ("n_valid", C.c_int32), ("split_k", C.c_int32),
        ("drop", DropoutCfg), ("pre_act", C.c_void_p), ("ld_pre", C.c_int64),
    ]


AUX_NONE, AUX_ADD_BF16, AUX_ADD_F32, AUX_RELU_MASK_BF16, AUX_GELU_BWD_BF16 = 0, 1, 2, 3, 4
LOSS_STATS_LEN = 104
EVAL_COUNTS_LEN = 914

_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
# flat-argument entry points (backward pass / optimizer): name -> argument types, the stream is appended
FLAT_OPS = {
    "fame_mask_kv_len": [_P, _I32, _I32, _P],
    "fame_attn_cls": [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _I64, _I32, _I32, _I32, _I32, _F],
    "fame_layernorm_bwd": [_P, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P],
    "fame_dropout_apply": [_P, _I32, _I64, _I32, _I32, _P],
    "fame_focal_loss_fwd_bwd": [_P, _P, _P, _F, _F, _I32, _P, _P, _P],
    "fame_relu_fwd": [_P, _I64],
    "fame_relu_bwd": [_P, _P, _I64],
    "fame_gelu_fwd": [_P, _P, _I64],
    "fame_gelu_bwd": [_P, _P, _P, _I64],
    "fame_colsum": [_P, _I32, _I64, _I32, _I32, _P],
    "fame_seq_mean_bwd": [_P, _P, _I32, _I32, _I32],
    "fame_lab_embed_bwd": [_P, _P, _P, _P, _P, _I32, _I32, _I32],
    "fame_attn_bwd_softmax": [_P, _P, _P, _P, _I64, _I32, _I32, _F],
    "fame_bert_embed_bwd": [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32],
    "fame_demo_add_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32],
    "fame_sgemm_small": [_P, _I64, _I64, _P, _I64, _I64, _P, _I64, _I32, _I32, _I32, _F, _I32],
    "fame_fusion_bwd_hidden": [_P, _P, _P, _P, _I32],
    "fame_fusion_bwd_gate": [_P, _P, _P, _F, _F, _F, _F, _P, _P, _I32, _P],
    "fame_grad_sumsq": [_P, _I64, _P],
    "fame_clip_adamw": [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _F, _I32, _P, _P, _P, _P],
    "fame_decay_only": [_P, _I64, _F, _F, _P, _P],
    "fame_cast_bf16": [_P, _P, _I64],
    "fame_transpose_bf16_table": [_P, _I32, _I32],
    "fame_wgrad_small": [_P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _P],
    "fame_attn_delta": [_P, _P, _I64, _P, _I32, _I32, _I32, _I32],
    "fame_attn_bwd_pds": [_P, _I64, _P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _F, _P],
    "fame_attn_bwd_fused": [_P, _I64, _P, _I64, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _F, _P],
}

# name -> args struct for every `int fame_<op>(const args*, void* ws, size_t ws_bytes, stream)` entry point
OP_TABLE = {
    "fame_gemm_bias_act": GemmArgs,
    "fame_layernorm": LayerNormArgs,
    "fame_bert_embed": BertEmbedArgs,
    "fame_attn_fwd": AttnFwdArgs,
    "fame_segment_mean": SegmentMeanArgs,
    "fame_lab_embed": LabEmbedArgs,
    "fame_seq_mean": SeqMeanArgs,
    "fame_demo_add": DemoAddArgs,
    "fame_embed_mean_add": EmbedMeanArgs,
    "fame_embed_mean_add_bwd": EmbedMeanArgs,
    "fame_fusion_fwd": FusionFwdArgs,
    "fame_loss_stats": LossStatsArgs,
    "fame_loss_fwd_bwd": LossFwdBwdArgs,
    "fame_eval_counts": EvalCountsArgs,
    "fame_rank_counts": RankCountsArgs,
    "fame_sigmoid_probs": SigmoidProbsArgs,
    "fame_gemm_ex": GemmExArgs,
}
PLAIN_SYMBOLS = ["fame_strerror", "fame_last_cuda_error", "fame_abi_version", "fame_device_check", "fame_sm_count",
                 "fame_set_sm_budget",
                 "fame_rank_counts_workspace_bytes", "fame_fusion_fwd_workspace_bytes"]
# C struct name -> ctypes mirror (tests compare sizeof() of both)
STRUCT_NAMES = {
    "fame_gemm_args": GemmArgs, "fame_layernorm_args": LayerNormArgs, "fame_bert_embed_args": BertEmbedArgs,
    "fame_attn_fwd_args": AttnFwdArgs, "fame_segment_mean_args": SegmentMeanArgs,
    "fame_lab_embed_args": LabEmbedArgs, "fame_seq_mean_args": SeqMeanArgs, "fame_demo_add_args": DemoAddArgs,
    "fame_embed_mean_args": EmbedMeanArgs,
    "fame_fusion_fwd_args": FusionFwdArgs, "fame_loss_stats_args": LossStatsArgs,
    "fame_loss_fwd_bwd_args": LossFwdBwdArgs, "fame_eval_counts_args": EvalCountsArgs,
    "fame_rank_counts_args": RankCountsArgs, "fame_sigmoid_probs_args": SigmoidProbsArgs,
    "fame_gemm_ex_args": GemmExArgs,
}

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
    lib.fame_set_sm_budget.restype = C.c_int
    lib.fame_set_sm_budget.argtypes = [C.c_int]
    lib.fame_rank_counts_workspace_bytes.restype = C.c_size_t
    lib.fame_rank_counts_workspace_bytes.argtypes = [C.c_int32]
    lib.fame_fusion_fwd_workspace_bytes.restype = C.c_size_t
    lib.fame_fusion_fwd_workspace_bytes.argtypes = [C.c_int32]
    for name, struct in OP_TABLE.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(struct), C.c_void_p, C.c_size_t, C.c_void_p]
    for name, argtypes in FLAT_OPS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = list(argtypes) + [C.c_void_p]
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


def call_flat(name: str, stream: int, *args) -> None:
    lib = load()
    check(getattr(lib, name)(*args, stream), name)

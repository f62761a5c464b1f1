/*
 * fame_b200.h -- C ABI of libfame_b200.so, the sm_100a (B200) kernel library behind the FAME hot path.
 *
 * The reference (AI-for-Health-Data/FairMultimodal, FinalCode/New/Final/10_FAME.py) has no FFI: its boundary is
 * the Python symbol surface of 10_FAME.py.  Each entry point below names the reference code it replaces
 * (file:line, relative to the reference root; "HF" = transformers/models/bert/modeling_bert.py 5.5.0, the
 * third-party module the reference calls at 10_FAME.py:140,188,199).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the field says "host"; the caller owns all buffers;
 *   - every call is asynchronous on `stream`, performs no allocation and no host synchronisation, and is
 *     CUDA-graph capturable;
 *   - return 0 on success, a negative FAME_ERR_* otherwise; nothing throws or aborts;
 *   - there is NO fallback: on a device that is not compute capability 10.x every call returns FAME_ERR_ARCH.
 */
#ifndef FAME_B200_H_
#define FAME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* fame_stream_t; /* == cudaStream_t */

enum {
    FAME_OK = 0,
    FAME_ERR_ARCH = -1,      /* device is not sm_100 */
    FAME_ERR_SHAPE = -2,     /* unsupported / inconsistent dimensions */
    FAME_ERR_ALIGN = -3,     /* pointer or leading dimension not 16-byte aligned */
    FAME_ERR_WORKSPACE = -4, /* workspace too small */
    FAME_ERR_NULLPTR = -5,   /* required pointer is NULL */
    FAME_ERR_CUDA = -6       /* a CUDA runtime / driver call failed (see fame_last_cuda_error) */
};

enum { FAME_ACT_NONE = 0, FAME_ACT_GELU_ERF = 1, FAME_ACT_RELU = 2 };
enum { FAME_DT_BF16 = 0, FAME_DT_F32 = 1 };

const char* fame_strerror(int code);
int fame_last_cuda_error(void); /* cudaError_t of the last FAME_ERR_CUDA on this thread */
int fame_abi_version(void);
int fame_device_check(void); /* FAME_OK iff the current device is compute capability 10.x */
int fame_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * K1  fame_gemm_bias_act:  Y[M,N] = act(X[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N])
 * X, W, residual: bf16 row-major; bias: f32; Y: bf16 or f32.  tcgen05 + TMEM + TMA.
 * Replaces nn.Linear: HF:179-181 (Q/K/V), HF:295 (attention output dense), HF:340 (intermediate dense + GELU),
 * HF:353 (output dense); nn.MultiheadAttention in/out projections and linear1/linear2 (10_FAME.py:214).
 * Requirements: K % 8 == 0, N % 8 == 0, ldx/ldw/ldr/ldy % 8 == 0, 16-byte aligned pointers. */
typedef struct {
    const void* x;
    int64_t ldx;
    const void* w;
    int64_t ldw;
    const float* bias; /* may be NULL */
    const void* residual; /* may be NULL */
    int64_t ldr;
    void* y;
    int64_t ldy;
    int32_t y_dtype; /* FAME_DT_* */
    int32_t M, N, K;
    int32_t act; /* FAME_ACT_* */
} fame_gemm_args;
int fame_gemm_bias_act(const fame_gemm_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3  fame_layernorm:  Y[r,:] = (X[r,:] - mean) * rsqrt(var + eps) * gamma + beta     (biased variance)
 * X bf16 [rows, cols] (residual already added by the producing GEMM), Y bf16.  cols % 8 == 0, cols <= 1024.
 * Replaces BertSelfOutput / BertOutput LayerNorm (HF:294-298, 352-356; eps 1e-12) and norm1 / norm2 of
 * nn.TransformerEncoderLayer (10_FAME.py:214; eps 1e-5). */
typedef struct {
    const void* x;
    int64_t ldx;
    const float* gamma;
    const float* beta;
    void* y;
    int64_t ldy;
    int32_t rows, cols;
    float eps;
} fame_layernorm_args;
int fame_layernorm(const fame_layernorm_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K4  fame_bert_embed:  Y[t,:] = LN(word[ids[t]] + type[0] + pos[t % seq_len])  -> bf16
 * Replaces BertEmbeddings.forward (HF:102-112).  ids int64 [tokens]; tables f32; hidden % 128 == 0, <= 1024.
 * Ids outside [0, vocab) set *err_flag (device int32, may be NULL) and are clamped. */
typedef struct {
    const int64_t* ids;
    const float* word; /* [vocab, hidden] */
    const float* pos;  /* [max_pos, hidden] */
    const float* type0; /* [hidden] (token_type row 0) */
    const float* gamma;
    const float* beta;
    void* y; /* bf16 [tokens, hidden] */
    int32_t* err_flag;
    int32_t tokens, seq_len, hidden, vocab;
    float eps;
} fame_bert_embed_args;
int fame_bert_embed(const fame_bert_embed_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2  fame_attn_fwd:  ctx = softmax(Q K^T * scale + key_mask) V   per (sequence, head); tcgen05 + TMEM.
 * qkv: bf16 [batch*seq, 3*heads*head_dim] packed [Q | K | V] (the output of the fused QKV GEMM);
 * key_mask: uint8 [batch, seq], 1 = attend, 0 = masked (additive -inf), may be NULL;
 * ctx: bf16 [batch*seq, heads*head_dim].  head_dim 64 or 96, any seq.  Masked keys get probability exactly 0; a row
 * whose keys are ALL masked yields zeros (the reference path never produces one: [CLS] is always attended).
 * Replaces the sdpa call in BertSelfAttention (HF:192-206) incl. the additive mask built at HF:709-713, and
 * F.multi_head_attention_forward's attention core inside nn.TransformerEncoderLayer (10_FAME.py:214,222). */
typedef struct {
    const void* qkv;
    int64_t ld_qkv;
    const uint8_t* key_mask;
    void* ctx;
    int64_t ld_ctx;
    int32_t batch, seq, heads, head_dim;
    float scale;
    int32_t algo; /* 0 = auto; 1 = full-row TMEM kernel (head_dim 64, seq <= 512); 2 = streaming flash kernel */
} fame_attn_fwd_args;
int fame_attn_fwd(const fame_attn_fwd_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K5  fame_segment_mean:  out[p,:] = mean_{c in [offsets[p], offsets[p+1])} x[c*ldx : c*ldx+cols]; zeros if empty.
 * The CLS gather is folded in through ldx (= seq_len*hidden when x is the encoder's last hidden state).
 * Replaces `outputs.last_hidden_state[:, 0, :]` (10_FAME.py:141) + np.vstack/np.mean(axis=0) and the
 * zero row for note-less patients (10_FAME.py:153-154,169-172).  offsets int32 [patients+1], non-decreasing.
 * cols % 8 == 0. */
typedef struct {
    const void* x;
    int64_t ldx;
    int32_t x_dtype; /* FAME_DT_* */
    const int32_t* offsets;
    float* out; /* [patients, cols] */
    int32_t patients, cols;
    int32_t mode; /* 0 = mean (aggregation="mean"), 1 = max (the reference's other aggregation branch) */
} fame_segment_mean_args;
int fame_segment_mean(const fame_segment_mean_args* a, void* workspace, size_t workspace_bytes,
                      fame_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FAME_B200_H_ */

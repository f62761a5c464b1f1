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
/* SMs the persistent tensor-core kernels (GEMM, attention) launched AFTER this call may occupy; 0 = all.  The
 * data-parallel training step leaves a few SMs to the latency-bound demographic-tower stream and to the NCCL
 * all-reduce kernels that run concurrently with the lab tower.  Host-side state, not a device setting. */
int fame_set_sm_budget(int sms);

/* ------------------------------------------------------------------------------------------------------------
 * Dropout of the training step (train() mode of the reference: HF hidden_dropout_prob / attention_probs_dropout_prob,
 * HF:111,205,297,355; nn.TransformerEncoderLayer dropout / dropout1 / dropout2 / self_attn.dropout, 10_FAME.py:214;
 * fusion_mlp[2], 10_FAME.py:255).  Masks are never stored: every kernel that applies a site's mask in the forward
 * pass or needs it again in the backward pass evaluates the same counter-based hash of (seed, *step, row, column),
 * fairmultimodal_b200/csrc/dropout.cuh.  Kept values are scaled by 65536 / (65536 - thresh16).  torch's Philox
 * stream is not reproducible by another implementation; the distribution (independent Bernoulli keeps) is. */
typedef struct {
    const int32_t* step;  /* device memory: step counter mixed into the seed at run time, so that replays of a captured
                             CUDA graph draw fresh masks; may be NULL (= 0) */
    uint32_t seed;        /* seed of this dropout site */
    uint32_t thresh16;    /* round(p * 65536); 0 = no dropout (all other fields ignored) */
    int32_t group_shift;  /* 2^group_shift consecutive columns share one draw (0 = per element; 6 = one draw per
                             64-wide attention head: dropout of a length-1 softmax, HF:205) */
} fame_dropout_cfg;

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
    int32_t residual_dtype; /* FAME_DT_BF16 (default) or FAME_DT_F32 (fp32 residual stream of the small-batch towers) */
    void* y;
    int64_t ldy;
    int32_t y_dtype; /* FAME_DT_* */
    int32_t M, N, K;
    int32_t act; /* FAME_ACT_* */
    fame_dropout_cfg drop; /* dropout of act(X W^T + bias) BEFORE the residual add; thresh16 = 0: none */
} fame_gemm_args;
int fame_gemm_bias_act(const fame_gemm_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* K1 (general)  fame_gemm_ex:  C[b0][b1][m,n] = epi( alpha * sum_k A(m,k) B(n,k) ), same tcgen05 kernel.
 * An operand is K-major (stored [rows = m|n, cols = k], mn_major = 0) or MN-major (stored [rows = k, cols = m|n],
 * mn_major = 1); this covers the three products of a linear layer without any transposed copies:
 *     forward  Y  = X W^T   : A = X (K-major),  B = W (K-major)
 *     dgrad    dX = dY W    : A = dY (K-major), B = W (MN-major)          [autograd of the nn.Linear calls above]
 *     wgrad    dW = dY^T X  : A = dY (MN-major), B = X (MN-major), f32 output
 * and, with the two batch levels (b0 = sequence, b1 = head), the five products of the attention backward.
 * aux: FAME_AUX_ADD_* adds a residual, FAME_AUX_RELU_MASK_BF16 zeroes the result where aux <= 0 (ReLU backward).
 * Strides and leading dimensions are in elements and must be multiples of 8; N % 8 == 0. */
enum { FAME_AUX_NONE = 0, FAME_AUX_ADD_BF16 = 1, FAME_AUX_ADD_F32 = 2, FAME_AUX_RELU_MASK_BF16 = 3,
       FAME_AUX_GELU_BWD_BF16 = 4 /* result *= gelu'(aux), aux = saved bf16 pre-activation; M <= 32 path only */ };
typedef struct {
    const void* ptr; /* bf16 */
    int64_t ld;
    int64_t stride_b0, stride_b1;
    int32_t mn_major;
} fame_gemm_operand;
typedef struct {
    fame_gemm_operand a, b;
    const float* bias; /* [N] or NULL */
    const void* aux;
    int64_t ld_aux, aux_stride_b0, aux_stride_b1;
    int32_t aux_mode;
    void* y;
    int64_t ldy, y_stride_b0, y_stride_b1;
    int32_t y_dtype;
    int32_t M, N, K, nb0, nb1;
    int32_t act;
    float alpha;
    int32_t n_valid; /* 0 = N; otherwise B has only n_valid (<= N) rows and output columns beyond it are zeros */
    int32_t split_k; /* 0 = Y is written.  != 0 = accumulate mode: Y (f32, initialised by the caller, e.g. the zeroed
                        gradient buffer) += result through float4 atomics, with the contraction cut into n slices
                        (n > 0) or into as many as fill the SMs (-1).  Used by the weight-gradient products, whose
                        contraction runs over all tokens while the output has only a few tiles. */
    fame_dropout_cfg drop; /* dropout of the result (after alpha / bias / act, before an additive aux); K-major A,
                              unbatched, split_k = 0 only.  thresh16 = 0: none */
    void* pre_act;         /* optional bf16 [M, ld_pre]: alpha * A B^T + bias before the activation (what the GELU backward
                              needs); M <= 32 (weight-streaming path) only, NULL otherwise */
    int64_t ld_pre;
} fame_gemm_ex_args;
int fame_gemm_ex(const fame_gemm_ex_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3  fame_layernorm:  Y[r,:] = (X[r,:] - mean) * rsqrt(var + eps) * gamma + beta     (biased variance)
 * X bf16 [rows, cols] (residual already added by the producing GEMM), Y bf16.  cols % 8 == 0, cols <= 1024.
 * Replaces BertSelfOutput / BertOutput LayerNorm (HF:294-298, 352-356; eps 1e-12) and norm1 / norm2 of
 * nn.TransformerEncoderLayer (10_FAME.py:214; eps 1e-5). */
typedef struct {
    const void* x;
    int64_t ldx;
    int32_t x_dtype; /* FAME_DT_BF16 or FAME_DT_F32 */
    const float* gamma;
    const float* beta;
    void* y;      /* bf16 output, may be NULL */
    float* y_f32; /* f32 output, may be NULL (at least one output is required); same ldy */
    int64_t ldy;
    float* stats; /* [rows][2] = {mean, rstd} saved for the backward pass, may be NULL */
    int32_t rows, cols;
    float eps;
    const void* residual; /* optional bf16 [rows, cols] (row stride ldr): Y = LN(bf16(X + residual)) -- the residual add
                             of BertSelfOutput / BertOutput when it does not ride in the producing GEMM's epilogue;
                             bf16 X and bf16 Y only; may be NULL */
    int64_t ldr;
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
    float* y_f32; /* optional f32 copy of the same rows, may be NULL */
    float* sum_out; /* optional f32 pre-LayerNorm sum (saved for the backward pass), may be NULL */
    float* stats;   /* optional [tokens][2] = {mean, rstd}, may be NULL */
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
    int32_t algo; /* 0: the persistent two-tile kernel in its default shape: head_dim 64 -> 64-key blocks, two CTAs per SM
                     (four query tiles / sixteen softmax warps per SM); head_dim 96 -> 128-key blocks, one CTA per SM.
                     3: 128-key blocks, one CTA per SM (any head_dim); 4: as 3 with the two softmax warpgroups taking
                     turns on the exponential phase; 5: 64-key blocks (head_dim 64).  3-5 exist for A/B measurements;
                     all variants produce the same result up to the order of the online-softmax blocks.
                     Other values: FAME_ERR_SHAPE */
    float* lse;   /* optional f32 [batch, heads, seq]: log2-domain log-sum-exp of each row of scaled scores, saved for
                     fame_attn_bwd_pds (P = 2^(s * scale * log2 e - lse)); may be NULL */
    const int32_t* kv_len; /* optional int32 [batch] from fame_mask_kv_len: 1 + index of the last attended key of each
                     sequence.  Key blocks (128 keys) that lie entirely beyond it hold only masked keys (probability
                     exactly 0) and are not loaded, multiplied or exponentiated; results are identical with and
                     without it.  may be NULL */
    fame_dropout_cfg drop; /* dropout of the attention probabilities (training): mask row = (sequence, head, query),
                     column = key; thresh16 = 0: none */
} fame_attn_fwd_args;
int fame_attn_fwd(const fame_attn_fwd_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* fame_mask_kv_len: |kv_len[b]| = 1 + max{k : key_mask[b, k] != 0} (0 when the row is all zero); the value is >= 0
 * when the row is a prefix mask (ones, then zeros: a padded tokenizer batch) and negative when the attended keys have
 * holes.  fame_attn_fwd skips key blocks beyond |kv_len| and, for prefix rows, derives key validity from the length
 * without reading the mask bytes.  key_mask uint8 [batch, seq] as handed to fame_attn_fwd (the reference's
 * attention_mask, HF:709-713), computed once per batch and shared by the 12 layers. */
int fame_mask_kv_len(const uint8_t* key_mask, int32_t batch, int32_t seq, int32_t* kv_len, fame_stream_t stream);

/* fame_attn_cls: attention for ONE query row per sequence (head_dim 64) -- the last encoder layer of the note
 * encoder, whose output the reference reads at token 0 only (`last_hidden_state[:, 0, :]`, 10_FAME.py:141; attention
 * per HF modeling_bert.py:192-206 with the key-padding mask of HF:709-713).  q bf16 [batch, heads * 64] (row stride
 * ld_q): the query projection of the CLS rows; kv bf16 [batch * seq, >= max(k_col0, v_col0) + heads * 64] (row stride
 * ld_kv): key / value projections of ALL tokens, head h at columns k_col0 + 64 h / v_col0 + 64 h; key_mask uint8
 * [batch, seq] or NULL; ctx bf16 [batch, heads * 64] (row stride ld_ctx).  A sequence whose keys are all masked gets a
 * zero row, like fame_attn_fwd.  HBM-bound streaming kernel (reads K and V once), no tensor cores. */
int fame_attn_cls(const void* q, int64_t ld_q, const void* kv, int64_t ld_kv, int32_t k_col0, int32_t v_col0,
                  const uint8_t* key_mask, void* ctx, int64_t ld_ctx, int32_t batch, int32_t seq, int32_t heads,
                  int32_t head_dim, float scale, fame_stream_t stream);

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

/* ------------------------------------------------------------------------------------------------------------
 * K4b fame_lab_embed:  Y[b*L + l, :] = lab[b,l] * w_tok + b_tok + pos[l, :]  -> bf16 [batch*L, hidden]
 * Replaces token_embedding + pos_embedding of BEHRTModel_Lab.forward (10_FAME.py:218-220). hidden % 8 == 0. */
typedef struct {
    const float* lab;   /* [batch, L] */
    const float* w_tok; /* [hidden] (Linear(1, hidden).weight[:, 0]) */
    const float* b_tok; /* [hidden] */
    const float* pos;   /* [L, hidden] */
    void* y;            /* bf16 [batch*L, hidden] */
    int32_t batch, L, hidden;
} fame_lab_embed_args;
int fame_lab_embed(const fame_lab_embed_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* K6 fame_seq_mean:  out[b, :] = mean_l x[b*L + l, :]   (x.mean(dim=1), 10_FAME.py:224).  x bf16, out f32. */
typedef struct {
    const void* x; /* bf16 [batch*L, cols] contiguous */
    float* out;    /* [batch, cols] */
    int32_t batch, L, cols;
} fame_seq_mean_args;
int fame_seq_mean(const fame_seq_mean_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* K4c fame_demo_add:  out[b,:] = cls[b*ld_cls : +hidden] + (E_age[clamp(age_b)] + E_gender[..] + E_eth[..] + E_ins[..]) / 4
 * (clamp + 4 lookups + average of BEHRTModel_Demo.forward, 10_FAME.py:195-206).  cls bf16, tables f32, out f32. */
typedef struct {
    const void* cls;
    int64_t ld_cls;
    int32_t cls_dtype;       /* FAME_DT_BF16 or FAME_DT_F32 */
    const int64_t* ids[4];   /* age, gender, ethnicity, insurance  [batch] */
    const float* table[4];   /* [n_rows[k], hidden] */
    int32_t n_rows[4];
    float* out;              /* [batch, hidden] */
    int32_t batch, hidden;
} fame_demo_add_args;
int fame_demo_add(const fame_demo_add_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* fame_embed_mean_add / fame_embed_mean_add_bwd: the same with n_tables (1..8) code tables -- the structured encoder of
 * the average-fusion ablation adds the mean of SEVEN clamped embedding rows (age, segment, admission location,
 * discharge location, gender, ethnicity, insurance) to the CLS state (07_multimodal_average_fusion.py:183-203):
 *   forward   out[b] = cls[b] + (sum_k table_k[clamp(ids_k[b], 0, n_rows_k - 1)]) / n_tables
 *   backward  dtable_k[clamp(ids_k[b])] += dout[b] / n_tables   (f32 atomics into caller-zeroed gradient tables)
 * The forward reads cls / table / out; the backward reads dout / dtable.  Unused slots are ignored. */
typedef struct {
    const void* cls;         /* forward: [batch, hidden] bf16 or f32, row stride ld_cls */
    int64_t ld_cls;
    int32_t cls_dtype;       /* FAME_DT_BF16 or FAME_DT_F32 */
    int32_t n_tables;        /* 1..8 */
    const int64_t* ids[8];   /* [batch] each */
    const float* table[8];   /* forward: [n_rows[k], hidden] */
    float* dtable[8];        /* backward: [n_rows[k], hidden] */
    int32_t n_rows[8];
    float* out;              /* forward: [batch, hidden] */
    const float* dout;       /* backward: [batch, hidden] */
    int32_t batch, hidden;
} fame_embed_mean_args;
int fame_embed_mean_add(const fame_embed_mean_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);
int fame_embed_mean_add_bwd(const fame_embed_mean_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K7 fame_fusion_fwd: EDDI-weighted, sigmoid-gated modality fusion and heads (10_FAME.py:276-308), fp32.
 * Weights wp_t / w3_t are the TRANSPOSED projector / fusion_mlp.0 matrices ([in, out]); hidden sizes are the
 * reference's fixed 768 -> 3 x 256 -> 768 -> 512 -> 3.  Optional outputs may be NULL. */
typedef struct {
    const float* emb[3];  /* demo, lab, text embeddings [B,768] */
    const float* wp_t;    /* [3][768][256] */
    const float* bp;      /* [3][256] */
    float w_mod[3];       /* (w_demo, w_lab, w_text) */
    const float* sig_w;   /* [768] */
    const float* w3_t;    /* [768][512] */
    const float* b3;      /* [512] */
    const float* w4;      /* [3][512] */
    const float* b4;      /* [3] */
    const float* wc;      /* [3][3][256] classifier_{demo,lab,text}.weight */
    const float* bc;      /* [3][3] */
    float* proj;          /* [B,768] relu(projector) outputs, unweighted */
    float* gated;         /* [B,768] "gated_vector" */
    float* pre_relu;      /* [B,512] "fusion_pre_relu" */
    float* logits;        /* [B,3]   "fused_logits" (required) */
    float* mod_logits;    /* [3][B][3] "modality_logits" */
    float* sig_out;       /* [768]   "sigmoid_weights" */
    int32_t B;
    const float* w_mod_dev; /* optional f32 [3] in DEVICE memory: when non-NULL the kernels read the modality weights
                             from it at run time and ignore w_mod -- a captured CUDA graph of the training step then
                             follows update_dynamic_weights_all_tasks (10_FAME.py:805-830) without re-capture */
} fame_fusion_fwd_args;
int fame_fusion_fwd(const fame_fusion_fwd_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);
/* Workspace needed when proj / gated / pre_relu are not all requested (small batches run as three launches whose
 * intermediates must live somewhere); 0 for batches that take the single fused kernel. */
size_t fame_fusion_fwd_workspace_bytes(int32_t B);

/* ------------------------------------------------------------------------------------------------------------
 * K8 joint loss of train_step (10_FAME.py:420-444), two passes so that data-parallel ranks can SUM-all-reduce the
 * statistics in between (every rank then evaluates the GLOBAL-batch loss and the gradient of its own patients):
 *   fame_loss_stats   : per-rank statistics -> stats[FAME_LOSS_STATS_LEN] (int64; counts and fixed-point sums)
 *   fame_loss_fwd_bwd : stats -> loss_out = {total, bce, leddi, l1} and dlogits = d total / d fused_logits
 * stats layout: [0..2] sum|p-y| (2^24 fixed point); [3..5] BCE sums (2^24); [6..77] subgroup error sums
 * [outcome][attr][code 0..7] (2^24); [78..101] subgroup counts [attr][code]; [102] patients; [103] bad-code flag.
 * Every patient's term is rounded to fixed point BEFORE it is added, so all 104 values are exact integer sums:
 * shards of a batch (data-parallel ranks) add up bit for bit to the statistics of the whole batch.
 * Subgroup membership = the int64 code itself (0..7), groups "present" = count > 0 (torch.unique, 10_FAME.py:432). */
#define FAME_LOSS_STATS_LEN 104
typedef struct {
    const float* logits; /* [B,3] */
    const float* labels; /* [B,3] */
    const int64_t* attr[3]; /* age_ids, ethnicity_ids, insurance_ids [B] */
    const float* pos_weight; /* [3] */
    int64_t* stats; /* [FAME_LOSS_STATS_LEN]; ACCUMULATED into (caller zeroes it) */
    int32_t B;
} fame_loss_stats_args;
int fame_loss_stats(const fame_loss_stats_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

typedef struct {
    const float* logits;
    const float* labels;
    const int64_t* attr[3];
    const float* pos_weight;
    const int64_t* stats; /* global statistics */
    const float* sig_w;   /* [n_sig] for the L1 term, may be NULL */
    int32_t n_sig;
    float lambda_edd, lambda_l1;
    float* dlogits;  /* [B,3], may be NULL (loss only) */
    float* loss_out; /* [4] */
    int32_t B;       /* local patients */
} fame_loss_fwd_bwd_args;
int fame_loss_fwd_bwd(const fame_loss_fwd_bwd_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K10 fame_eval_counts: one pass over predictions -> integer counts (uint64, accumulated; caller zeroes):
 *   out[((o*3 + a)*8 + code)*4 + {TP,FN,FP,TN}]  per outcome o, attribute a (age, ethnicity, insurance), code 0..7
 *   out[288 + o*4 + {TP,FN,FP,TN}]               per outcome
 *   out[300 + (o*2 + label)*102 + k]             F1-sweep histogram: k = #sweep thresholds strictly below p
 *   out[912] = patients, out[913] = bad-code flag
 * prediction = (double)float32(sigmoid(logit)) > thr[o]  (strict, float64 compare: 10_FAME.py:55, 476, 518).
 * Replaces compute_eddi / calculate_tpr_and_fpr / confusion_matrix / the f1 sweep (10_FAME.py:54-97, 470-481, 514-540). */
#define FAME_EVAL_COUNTS_LEN 914
typedef struct {
    const float* logits; /* [N, ld] */
    int64_t ld;
    const float* labels; /* [N,3] */
    const int64_t* attr[3];
    double thr[3];
    const double* sweep; /* device [101] or NULL */
    uint64_t* out;
    int32_t N;
    int32_t logits_are_probs;
} fame_eval_counts_args;
int fame_eval_counts(const fame_eval_counts_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* fame_rank_counts: exact tie-aware rank statistics of one outcome for samples [i0, i1) against all N samples.
 *   auroc2 += sum_{i positive} (#neg with s >= s_i) + (#neg with s > s_i);   AUROC = 1 - auroc2 / (2 Npos Nneg)
 *   ap_sum += sum_{i positive} #pos(s >= s_i) / #all(s >= s_i);              AP    = ap_sum / Npos
 *   npos_nneg[0..1] += positives / negatives among [i0, i1)
 * Equal to sklearn roc_auc_score / average_precision_score on the same float32 scores (10_FAME.py:520-526).
 * workspace: fame_rank_counts_workspace_bytes(i1 - i0). */
typedef struct {
    const float* scores;  /* [N] float32 probabilities */
    const uint8_t* y;     /* [N] 0/1 */
    int32_t N, i0, i1;
    uint64_t* auroc2;
    double* ap_sum;
    uint64_t* npos_nneg;
} fame_rank_counts_args;
size_t fame_rank_counts_workspace_bytes(int32_t n_i);
int fame_rank_counts(const fame_rank_counts_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* fame_sigmoid_probs: p = float32 sigmoid (correctly rounded) of logits[:, outcome]; y8 = (labels[:, outcome] != 0). */
typedef struct {
    const float* logits;
    int64_t ld;
    const float* labels; /* [N,3] or NULL */
    float* probs;        /* [3][N] outcome-major */
    uint8_t* y8;         /* [3][N] or NULL */
    int32_t N;
} fame_sigmoid_probs_args;
int fame_sigmoid_probs(const fame_sigmoid_probs_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream);

/* ============================================================================================================
 * Backward pass and optimizer of train_step (10_FAME.py:444-447: total_loss.backward(), clip_grad_norm_(1.0),
 * AdamW.step()).  The reference gets these from torch.autograd / torch.optim; here they are explicit kernels.
 * Tensor-core products (dgrad / wgrad / attention backward) go through fame_gemm_ex; the entry points below take
 * flat argument lists (device pointers, sizes, scalars) and are asynchronous on `stream` like everything else.
 * ============================================================================================================ */

/* LayerNorm backward from the saved {mean, rstd}: dx (bf16 and/or f32), dgamma/dbeta ACCUMULATED with f32 atomics.
 * dx_drop (bf16, optional): dx with the per-element dropout mask `drop` re-applied -- the gradient of the linear layer
 * inside  t = residual + dropout(linear(.))  (the residual branch takes the unmasked dx); NULL when the site has no
 * dropout. */
int fame_layernorm_bwd(const void* x, int32_t x_dtype, const void* dy, int32_t dy_dtype, const float* stats,
                       const float* gamma, void* dx_bf16, float* dx_f32, float* dgamma, float* dbeta, int32_t rows,
                       int32_t cols, void* dx_drop, const fame_dropout_cfg* drop,
                       const void* residual /* bf16 [rows, cols] or NULL: the LayerNorm input was x + residual */, fame_stream_t stream);
/* In-place dropout of x [rows, cols] (bf16 or f32, row stride ld): the sites with no producing GEMM epilogue
 * (BertEmbeddings dropout HF:111, fusion_mlp[2] 10_FAME.py:255) and their gradients. */
int fame_dropout_apply(void* x, int32_t x_dtype, int64_t ld, int32_t rows, int32_t cols, const fame_dropout_cfg* drop,
                       fame_stream_t stream);
/* erf-GELU forward on a saved bf16 pre-activation and its backward dpre = dh * gelu'(pre); n % 8 == 0. */
int fame_gelu_fwd(const void* pre, void* h, int64_t n, fame_stream_t stream);
int fame_gelu_bwd(const void* pre, const void* dh, void* dpre, int64_t n, fame_stream_t stream);
/* out[c] += sum_r x[r, c]  (bias gradients). */
int fame_colsum(const void* x, int32_t x_dtype, int64_t ld, int32_t rows, int32_t cols, float* out, fame_stream_t stream);
/* Backward of x.mean(dim=1) (10_FAME.py:224): dx[b*L + l, :] = dout[b, :] / L  (bf16). */
int fame_seq_mean_bwd(const float* dout, void* dx, int32_t batch, int32_t L, int32_t cols, fame_stream_t stream);
/* Backward of the lab token embedding (10_FAME.py:218-220): dpos [L, hidden] written (fixed summation order), dw / dbias
 * [hidden] accumulated with one f32 atomic per column and CTA.  dx bf16 [batch * L, hidden], 16-byte aligned;
 * hidden % 8 == 0, hidden <= 3072. */
int fame_lab_embed_bwd(const void* dx, const float* lab, float* dpos, float* dw, float* dbias, int32_t batch, int32_t L,
                       int32_t hidden, fame_stream_t stream);
/* Attention backward, softmax part: from f32 scores S = Q K^T and dP = dO V^T (rows = batch*heads*seq, leading
 * dimension ld >= seq) produce P = softmax(scale S) and dS = scale P (dP - rowsum(P dP)) as bf16 [rows, ld]. */
int fame_attn_bwd_softmax(const float* s, const float* dp, void* p, void* ds, int64_t rows, int32_t seq, int32_t ld,
                          float scale, fame_stream_t stream);
/* Scatter the gradient of the BERT embedding sum into word / position / token-type tables (padding row skipped). */
int fame_bert_embed_bwd(const float* d_sum, const int64_t* ids, float* dword, float* dpos, float* dtype0, int32_t tokens,
                        int32_t seq, int32_t hidden, int32_t vocab, int32_t pad_idx, fame_stream_t stream);
/* dE_k[clamp(ids_k[b])] += dout[b] / 4 for the four demographic tables (10_FAME.py:201-205). */
int fame_demo_add_bwd(const float* dout, const int64_t* ids0, const int64_t* ids1, const int64_t* ids2, const int64_t* ids3,
                      float* t0, float* t1, float* t2, float* t3, int32_t n0, int32_t n1, int32_t n2, int32_t n3,
                      int32_t batch, int32_t hidden, fame_stream_t stream);
/* Small fp32 product with arbitrary strides, C = alpha * A B (+ C): the fusion head's backward matrices. */
int fame_sgemm_small(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbk, int64_t sbn, float* c,
                     int64_t ldc, int32_t M, int32_t N, int32_t K, float alpha, int32_t accumulate, fame_stream_t stream);
/* Fusion head backward, elementwise stages (10_FAME.py:287-296): hidden-layer ReLU mask; gate / sig_weights / L1. */
int fame_fusion_bwd_hidden(const float* dlogits, const float* w4, const float* pre, float* dhid, int32_t B,
                           fame_stream_t stream);
/* w_dev: optional f32 [3] in device memory overriding (w0, w1, w2), as fame_fusion_fwd_args.w_mod_dev. */
int fame_fusion_bwd_gate(const float* dgated, const float* proj, const float* sig_w, float w0, float w1, float w2,
                         float lambda_l1, float* dproj, float* dsig, int32_t B, const float* w_dev,
                         fame_stream_t stream);
/* K9: *out += sum g^2 (float64); then clip_grad_norm_(max_norm) + one AdamW step over flat f32 buffers. */
int fame_grad_sumsq(const float* g, int64_t n, double* out, fame_stream_t stream);
int fame_clip_adamw(float* p, const float* g, float* m, float* v, int64_t n, const double* sumsq, float max_norm, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int32_t step, float* grad_norm_out,
                    const int32_t* step_dev, const float* hyper_dev, void* p_bf16, fame_stream_t stream);
/* fame_decay_only: the AdamW step of a parameter range whose gradient and Adam moments are identically zero (the
 * query / key projection weights of the length-1 demographic BERT, 10_FAME.py:199): p <- p (1 - lr wd),
 * bf16 shadow refreshed.  Equal to fame_clip_adamw on that range (update term 0 / (0 + eps)), at 10 instead of 32 bytes
 * per parameter.  hyper_dev {lr, weight_decay} overrides the scalars when not NULL. */
int fame_decay_only(float* p, int64_t n, float lr, float weight_decay, const float* hyper_dev, void* p_bf16,
                    fame_stream_t stream);
/* step_dev / hyper_dev = {lr, weight_decay} (device, may be NULL) override the by-value arguments at run time so that a
 * captured CUDA graph follows the step count and learning-rate schedule; p_bf16 (may be NULL) receives a bf16 copy of
 * the updated parameters for the tensor-core GEMMs. */
int fame_cast_bf16(const float* x, void* y, int64_t n, fame_stream_t stream);
/* Transposed bf16 shadows of a table of weight matrices in ONE launch (data-gradient products of the <= 32-row
 * demographic tower read W^T rows).  table: device array of n_entries records
 *   { const bf16* src [rows, cols]; bf16* dst [cols, rows]; int32 rows, cols, tile0, tiles_x; }   (32 bytes each)
 * tile0 = index of the record's first 64x64 tile within the launch, tiles_x = ceil(cols / 64). */
/* Weight gradient of a layer with at most 32 rows (demographic tower): dW[N,K] (+)= dY[M,N]^T . X[M,K]; bf16 operands,
 * f32 result; accumulate = 0 overwrites dW.  dbias (optional, [N]): the bias gradient sum_m dY[m, n] from the same
 * staged tile, written (or added when accumulate) by the first k-block.  Bound by the gradient write, CUDA-core FMAs
 * (skinny_gemm.cuh). */
int fame_wgrad_small(const void* dy, int64_t ld_dy, const void* x, int64_t ld_x, float* out, int64_t ld_out, int32_t M,
                     int32_t N, int32_t K, int32_t accumulate, float* dbias, fame_stream_t stream);
/* Attention backward, first half (autograd of the sdpa call inside nn.MultiheadAttention, 10_FAME.py:214,445):
 * P = softmax(Q K^T scale) and dS = scale P (dO V^T - delta) per (sequence, head), both bf16 [batch, heads, seq, ldp]
 * (columns >= seq zero), from the packed qkv tensor, dO = dctx, the forward's lse and delta = rowsum(dO * O)
 * (fame_attn_delta).  Both score products stay in TMEM.  The three products dV = P^T dO, dK = dS^T Q, dQ = dS K follow
 * through fame_gemm_ex.  head_dim 64 or 96; ldp % 8 == 0, ldp >= seq.
 * fame_attn_delta: delta f32 [batch, heads, seq] from dctx and ctx (bf16 [batch * seq, ld], ONE row stride for both; any
 * even head_dim -- heads that divide 32 with head_dim * heads % 256 == 0 take the one-warp-per-token kernel). */
int fame_attn_delta(const void* dctx, const void* ctx, int64_t ld, float* delta, int32_t batch, int32_t seq,
                    int32_t heads, int32_t head_dim, fame_stream_t stream);
int fame_attn_bwd_pds(const void* qkv, int64_t ld_qkv, const void* dctx, int64_t ld_dctx, const float* lse,
                      const float* delta, void* p, void* ds, int64_t ldp, int32_t batch, int32_t seq, int32_t heads,
                      int32_t head_dim, float scale, const fame_dropout_cfg* drop /* the forward's, or NULL */,
                      fame_stream_t stream);
/* fame_attn_bwd_fused: the whole attention backward of one layer in two launches, P and dS never written to memory
 * (10_FAME.py:212-215 under total_loss.backward(), 10_FAME.py:445): dqkv [batch * seq, ld_dqkv] bf16 receives dQ | dK | dV in
 * the layout of the packed qkv tensor.  Pass 1 (key tiles stationary) recomputes S^T = K Q^T and dP^T = V dO^T per 64-query
 * half block, turns them into bf16 P^T / dS^T inside TMEM and accumulates dV += P^T dO, dK += dS^T Q; pass 2 (query tiles
 * stationary) does the same for dQ += dS K.  lse from fame_attn_fwd, delta from fame_attn_delta; drop = the forward's
 * dropout configuration or NULL.  head_dim 64 or 96. */
int fame_attn_bwd_fused(const void* qkv, int64_t ld_qkv, const void* dctx, int64_t ld_dctx, const float* lse,
                        const float* delta, void* dqkv, int64_t ld_dqkv, int32_t batch, int32_t seq, int32_t heads,
                        int32_t head_dim, float scale, const fame_dropout_cfg* drop /* the forward's, or NULL */,
                        fame_stream_t stream);
/* Text-only baseline (02_BioClinicalBERT.py): FocalLoss(gamma, alpha, pos_weight_i) summed over the three
 * outcomes, each a batch mean (18-38, 143-147): *loss_out += loss (float64, caller zeroes it); dlogits [batch, 3] =
 * d loss / d logits (may be NULL).  fame_relu_fwd / fame_relu_bwd: ReLU of the classifier's 256-wide hidden layer, in
 * place, and its backward mask (dh = pre > 0 ? dh : 0).  batch_total_dev (optional, device int64 [1]): the GLOBAL
 * batch size when `batch` is one data-parallel rank's shard -- the mean then runs over it, so the SUM over ranks of
 * loss / gradients equals the single-process result on the concatenated batch. */
int fame_focal_loss_fwd_bwd(const float* logits, const float* labels, const float* pos_weight, float gamma, float alpha,
                            int32_t batch, double* loss_out, float* dlogits, const int64_t* batch_total_dev,
                            fame_stream_t stream);
int fame_relu_fwd(float* x, int64_t n, fame_stream_t stream);
int fame_relu_bwd(float* dh, const float* pre, int64_t n, fame_stream_t stream);
int fame_transpose_bf16_table(const void* table, int32_t n_entries, int32_t total_tiles, fame_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FAME_B200_H_ */

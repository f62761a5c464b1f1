// Host side of libfame_b200.so: argument validation, TMA descriptor encoding, kernel launches.
// See include/fame_b200.h for the contract of each entry point.
#include "../../include/fame_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "attn_cls.cuh"
#include "attn_pair_sm100.cuh"
#include "attn_bwd_sm100.cuh"
#include "attn_bwd_fused_sm100.cuh"
#include "backward.cuh"
#include "embed_mean.cuh"
#include "gemm_sm100.cuh"
#include "heads.cuh"
#include "metrics.cuh"
#include "rowwise.cuh"
#include "skinny_gemm.cuh"

#ifndef FAME_USE_SKINNY_GEMM
#define FAME_USE_SKINNY_GEMM 1
#endif

namespace {

// SMs the persistent tensor-core kernels (GEMM, attention forward / backward) may occupy; 0 = all of them.
static int g_sm_budget = 0;
static inline int persistent_sms(int sm_count) {
    return (g_sm_budget > 0 && g_sm_budget < sm_count) ? g_sm_budget : sm_count;
}


thread_local int g_last_cuda_error = 0;

int cuda_fail(cudaError_t e) {
    g_last_cuda_error = static_cast<int>(e);
    return FAME_ERR_CUDA;
}

struct DeviceInfo {
    int checked = 0;
    int ok = 0;
    int sm_count = 0;
};
DeviceInfo g_dev[64];

int device_info(DeviceInfo** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    if (dev < 0 || dev >= 64) return FAME_ERR_ARCH;
    DeviceInfo& d = g_dev[dev];
    if (!d.checked) {
        int major = 0, sms = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) return cuda_fail(e);
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return cuda_fail(e);
        d.ok = (major == 10);
        d.sm_count = sms;
        d.checked = 1;
    }
    *out = &d;
    return d.ok ? FAME_OK : FAME_ERR_ARCH;
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encode() {
    if (g_encode != nullptr) return FAME_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return cuda_fail(e);
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) return cuda_fail(cudaErrorNotSupported);
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return FAME_OK;
}

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = box_rows x 64 columns, 128B swizzle.
int encode_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    int rc = get_encode();
    if (rc != FAME_OK) return rc;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = 100000 + static_cast<int>(r);
        return FAME_ERR_CUDA;
    }
    return FAME_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FAME_OK : cuda_fail(e);
}

#include "gemm_host.inc"

template <int D, bool kDrop, int KB>
static int launch_pair_t(const fame_attn_fwd_args* a, const CUtensorMap& tq, const CUtensorMap& tkv, int sm_count,
                         fame_stream_t stream) {
    using Cfg = fame::ApCfg<D, KB>;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fame::attn_fwd_pair_kernel<D, kDrop, KB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return cuda_fail(e);
        attr_set[dev] = true;
    }
    fame::FaParams p;
    p.key_mask = a->key_mask;
    p.ctx = reinterpret_cast<__nv_bfloat16*>(a->ctx);
    p.ld_ctx = a->ld_ctx;
    p.batch = a->batch;
    p.seq = a->seq;
    p.heads = a->heads;
    p.q_col0 = 0;
    p.k_col0 = a->heads * D;
    p.v_col0 = 2 * a->heads * D;
    p.scale_log2e = a->scale * 1.4426950408889634f;
    p.lse = a->lse;
    p.kv_len = a->kv_len;
    p.drop = drop_cfg(a->drop);
    p.pingpong = a->algo == 4 ? 1 : 0;      // algo 4: the warpgroups take turns on the exponential phase (A/B only)
    const int qpairs = (a->seq + 255) / 256;
    const long long items = (long long)a->batch * a->heads * qpairs;
    if (items > 0x7fffffffll) return FAME_ERR_SHAPE;
    const int slots = persistent_sms(sm_count) * Cfg::kCtasPerSm;
    const int grid = items < slots ? (int)items : slots;
    fame::attn_fwd_pair_kernel<D, kDrop, KB><<<grid, fame::kApThreads, Cfg::kSmemBytes, stream>>>(tq, tkv, p, (int)items,
                                                                                              qpairs);
    return launch_status();
}

template <int D, int KB>
static int launch_pair(const fame_attn_fwd_args* a, const CUtensorMap& tq, const CUtensorMap& tkv, int sm_count,
                       fame_stream_t stream) {
    if (a->drop.thresh16 != 0) {
        if (a->drop.thresh16 >= 65536u || a->drop.group_shift != 0) return FAME_ERR_SHAPE;
        return launch_pair_t<D, true, KB>(a, tq, tkv, sm_count, stream);
    }
    return launch_pair_t<D, false, KB>(a, tq, tkv, sm_count, stream);
}


template <int D, bool kDrop>
static int launch_attn_bwd_pds(const CUtensorMap& tq, const CUtensorMap& tdo, const fame::AbParams& p, int sm_count,
                               fame_stream_t stream) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fame::attn_bwd_pds_kernel<D, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             fame::AbCfg<D>::kSmemBytes);
        if (e != cudaSuccess) return cuda_fail(e);
        attr_set[dev] = true;
    }
    const int qtiles = (p.seq + 127) / 128;
    const long long items = (long long)p.batch * p.heads * qtiles;
    if (items > 0x7fffffffll) return FAME_ERR_SHAPE;
    const int sms = persistent_sms(sm_count);
    const int grid = items < sms ? (int)items : sms;
    fame::attn_bwd_pds_kernel<D, kDrop><<<grid, fame::kAbThreads, fame::AbCfg<D>::kSmemBytes, stream>>>(tq, tdo, p, (int)items,
                                                                                                qtiles);
    return launch_status();
}

template <int D, bool kKV, bool kDrop>
static int launch_attn_bwd_fused(const CUtensorMap& tq128, const CUtensorMap& td128, const CUtensorMap& tq64,
                                 const CUtensorMap& td64, const fame::AfParams& p, int sm_count, fame_stream_t stream) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fame::attn_bwd_fused_kernel<D, kKV, kDrop>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, fame::AfCfg<D>::kSmemBytes);
        if (e != cudaSuccess) return cuda_fail(e);
        attr_set[dev] = true;
    }
    const int rtiles = (p.seq + 127) / 128;
    const long long items = (long long)p.batch * p.heads * rtiles;
    if (items > 0x7fffffffll) return FAME_ERR_SHAPE;
    const int sms = persistent_sms(sm_count);
    const int grid = items < sms ? (int)items : sms;
    fame::attn_bwd_fused_kernel<D, kKV, kDrop><<<grid, fame::kAfThreads, fame::AfCfg<D>::kSmemBytes, stream>>>(
        tq128, td128, tq64, td64, p, (int)items, rtiles);
    return launch_status();
}

template <int D, bool kDrop>
static int launch_attn_bwd_fused_both(const CUtensorMap& tq128, const CUtensorMap& td128, const CUtensorMap& tq64,
                                      const CUtensorMap& td64, const fame::AfParams& p, int sm_count, fame_stream_t stream) {
    int rc = launch_attn_bwd_fused<D, true, kDrop>(tq128, td128, tq64, td64, p, sm_count, stream);     // dK, dV
    if (rc != FAME_OK) return rc;
    return launch_attn_bwd_fused<D, false, kDrop>(tq128, td128, tq64, td64, p, sm_count, stream);      // dQ
}


}  // namespace

extern "C" {

const char* fame_strerror(int code) {
    switch (code) {
        case FAME_OK: return "ok";
        case FAME_ERR_ARCH: return "device is not sm_100 (B200); this library has no fallback path";
        case FAME_ERR_SHAPE: return "unsupported or inconsistent shape";
        case FAME_ERR_ALIGN: return "pointer or leading dimension is not 16-byte aligned";
        case FAME_ERR_WORKSPACE: return "workspace too small";
        case FAME_ERR_NULLPTR: return "required pointer is NULL";
        case FAME_ERR_CUDA: return "CUDA call failed (see fame_last_cuda_error)";
        default: return "unknown fame error";
    }
}

int fame_last_cuda_error(void) { return g_last_cuda_error; }
int fame_abi_version(void) { return 1; }

int fame_device_check(void) {
    DeviceInfo* d = nullptr;
    return device_info(&d);
}

int fame_set_sm_budget(int sms) {
    if (sms < 0) return FAME_ERR_SHAPE;
    g_sm_budget = sms;
    return FAME_OK;
}

int fame_sm_count(void) {
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    return rc == FAME_OK ? d->sm_count : rc;
}

// ------------------------------------------------------------------------------------------------ K1
int fame_gemm_bias_act(const fame_gemm_args* a, void* /*workspace*/, size_t /*workspace_bytes*/,
                       fame_stream_t stream) {
    if (a == nullptr || a->x == nullptr || a->w == nullptr || a->y == nullptr) return FAME_ERR_NULLPTR;
    if (a->M < 0 || a->N <= 0 || a->K <= 0) return FAME_ERR_SHAPE;
    if ((a->K & 7) || (a->N & 7)) return FAME_ERR_SHAPE;
    if (a->act < FAME_ACT_NONE || a->act > FAME_ACT_RELU) return FAME_ERR_SHAPE;
    if (a->y_dtype != FAME_DT_BF16 && a->y_dtype != FAME_DT_F32) return FAME_ERR_SHAPE;
    if (a->ldx < a->K || a->ldw < a->K || a->ldy < a->N) return FAME_ERR_SHAPE;
    if ((a->ldx & 7) || (a->ldw & 7) || (a->ldy & 7)) return FAME_ERR_ALIGN;
    if (!aligned16(a->x) || !aligned16(a->w) || !aligned16(a->y)) return FAME_ERR_ALIGN;
    if (a->bias != nullptr && !aligned16(a->bias)) return FAME_ERR_ALIGN;
    if (a->residual != nullptr && (!aligned16(a->residual) || (a->ldr & 7) || a->ldr < a->N)) return FAME_ERR_ALIGN;
    if (a->residual_dtype != FAME_DT_BF16 && a->residual_dtype != FAME_DT_F32) return FAME_ERR_SHAPE;
    fame_gemm_ex_args e = {};
    e.a.ptr = a->x; e.a.ld = a->ldx;
    e.b.ptr = a->w; e.b.ld = a->ldw;
    e.bias = a->bias;
    if (a->residual != nullptr) {
        e.aux = a->residual;
        e.ld_aux = a->ldr;
        e.aux_mode = a->residual_dtype == FAME_DT_F32 ? FAME_AUX_ADD_F32 : FAME_AUX_ADD_BF16;
    }
    e.y = a->y; e.y_dtype = a->y_dtype; e.ldy = a->ldy;
    e.M = a->M; e.N = a->N; e.K = a->K;
    e.nb0 = e.nb1 = 1;
    e.act = a->act;
    e.alpha = 1.0f;
    e.drop = a->drop;
    return gemm_ex(&e, stream);
}

int fame_gemm_ex(const fame_gemm_ex_args* a, void* /*workspace*/, size_t /*workspace_bytes*/, fame_stream_t stream) {
    return gemm_ex(a, stream);
}

// ------------------------------------------------------------------------------------------------ K3
int fame_layernorm(const fame_layernorm_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->x == nullptr || a->gamma == nullptr || a->beta == nullptr) return FAME_ERR_NULLPTR;
    if (a->y == nullptr && a->y_f32 == nullptr) return FAME_ERR_NULLPTR;
    if (a->rows < 0 || a->cols <= 0 || (a->cols & 7) || a->cols > 1024) return FAME_ERR_SHAPE;
    if (a->x_dtype != FAME_DT_BF16 && a->x_dtype != FAME_DT_F32) return FAME_ERR_SHAPE;
    if ((a->ldx & 7) || (a->ldy & 7) || a->ldx < a->cols || a->ldy < a->cols) return FAME_ERR_ALIGN;
    if (!aligned16(a->x) || !aligned16(a->gamma) || !aligned16(a->beta)) return FAME_ERR_ALIGN;
    if ((a->y != nullptr && !aligned16(a->y)) || (a->y_f32 != nullptr && !aligned16(a->y_f32))) return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->rows == 0) return FAME_OK;
    if (a->residual != nullptr) {
        if (a->x_dtype != FAME_DT_BF16 || a->y == nullptr || a->y_f32 != nullptr) return FAME_ERR_SHAPE;
        if (!aligned16(a->residual) || (a->ldr & 7) || a->ldr < a->cols) return FAME_ERR_ALIGN;
    }
    if (a->x_dtype == FAME_DT_BF16 && a->y != nullptr && a->y_f32 == nullptr) {
        // bf16 -> bf16: two rows per warp, packed registers (rowwise.cuh)
        const int per_block = 2 * fame::kLn2WarpsPerBlock;
        const int grid2 = (a->rows + per_block - 1) / per_block;
        const __nv_bfloat16* xr = reinterpret_cast<const __nv_bfloat16*>(a->x);
        const __nv_bfloat16* rr = reinterpret_cast<const __nv_bfloat16*>(a->residual);
        __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(a->y);
        float2* st = reinterpret_cast<float2*>(a->stats);
        const int th = fame::kLn2WarpsPerBlock * 32;
#define FAME_LN2(CH, RES) \
    fame::layernorm_bf16_rows2_kernel<CH, RES><<<grid2, th, 0, stream>>>(xr, a->ldx, a->gamma, a->beta, yr, a->ldy, st, \
                                                                         a->rows, a->cols, a->eps, rr, a->ldr)
        if (a->cols <= 768) {
            if (rr != nullptr) FAME_LN2(3, true); else FAME_LN2(3, false);
        } else {
            if (rr != nullptr) FAME_LN2(4, true); else FAME_LN2(4, false);
        }
#undef FAME_LN2
        return launch_status();
    }
    const int grid = (a->rows + fame::kLnWarpsPerBlock - 1) / fame::kLnWarpsPerBlock;
    if (a->x_dtype == FAME_DT_F32)
        fame::layernorm_kernel<true><<<grid, fame::kLnWarpsPerBlock * 32, 0, stream>>>(
            a->x, a->ldx, a->gamma, a->beta, reinterpret_cast<__nv_bfloat16*>(a->y), a->y_f32, a->ldy,
            reinterpret_cast<float2*>(a->stats), a->rows, a->cols, a->eps);
    else
        fame::layernorm_kernel<false><<<grid, fame::kLnWarpsPerBlock * 32, 0, stream>>>(
            a->x, a->ldx, a->gamma, a->beta, reinterpret_cast<__nv_bfloat16*>(a->y), a->y_f32, a->ldy,
            reinterpret_cast<float2*>(a->stats), a->rows, a->cols, a->eps);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------ K4
int fame_bert_embed(const fame_bert_embed_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->ids == nullptr || a->word == nullptr || a->pos == nullptr || a->type0 == nullptr ||
        a->gamma == nullptr || a->beta == nullptr || a->y == nullptr)
        return FAME_ERR_NULLPTR;
    if (a->tokens < 0 || a->seq_len <= 0 || a->vocab <= 0 || a->hidden <= 0 || (a->hidden & 127) || a->hidden > 1024)
        return FAME_ERR_SHAPE;
    if (!aligned16(a->word) || !aligned16(a->pos) || !aligned16(a->type0) || !aligned16(a->gamma) ||
        !aligned16(a->beta) || !aligned16(a->y))
        return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->tokens == 0) return FAME_OK;
    const int grid = (a->tokens + fame::kLnWarpsPerBlock - 1) / fame::kLnWarpsPerBlock;
    fame::bert_embed_ln_kernel<<<grid, fame::kLnWarpsPerBlock * 32, 0, stream>>>(
        reinterpret_cast<const long long*>(a->ids), a->word, a->pos, a->type0, a->gamma, a->beta,
        reinterpret_cast<__nv_bfloat16*>(a->y), a->y_f32, a->sum_out, reinterpret_cast<float2*>(a->stats), a->err_flag,
        a->tokens, a->seq_len, a->hidden, a->vocab, a->eps);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------ K2
int fame_attn_fwd(const fame_attn_fwd_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->qkv == nullptr || a->ctx == nullptr) return FAME_ERR_NULLPTR;
    if ((a->head_dim != 64 && a->head_dim != 96) || a->seq <= 0 || a->heads <= 0 || a->batch < 0)
        return FAME_ERR_SHAPE;
    if (a->algo != 0 && (a->algo < 3 || a->algo > 5)) return FAME_ERR_SHAPE;   // see fame_attn_fwd_args.algo
    const int64_t width = 3ll * a->heads * a->head_dim;
    if (a->ld_qkv < width || a->ld_ctx < (int64_t)a->heads * a->head_dim) return FAME_ERR_SHAPE;
    if ((a->ld_qkv & 7) || (a->ld_ctx & 7) || !aligned16(a->qkv) || !aligned16(a->ctx)) return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->batch == 0) return FAME_OK;
    if (a->batch > 65535 || a->heads > 65535) return FAME_ERR_SHAPE;

    // head_dim 64: 64-key blocks, two CTAs per SM (algo 0 / 5); algo 3 / 4 force the 128-key, one-CTA-per-SM variant
    const bool kb64 = a->head_dim == 64 && (a->algo == 0 || a->algo == 5);
    CUtensorMap tq, tkv;
    rc = encode_bf16_2d(&tq, a->qkv, (uint64_t)a->batch * a->seq, (uint64_t)width, (uint64_t)a->ld_qkv, 128);
    if (rc != FAME_OK) return rc;
    rc = encode_bf16_2d(&tkv, a->qkv, (uint64_t)a->batch * a->seq, (uint64_t)width, (uint64_t)a->ld_qkv, kb64 ? 64 : 128);
    if (rc != FAME_OK) return rc;
    if (a->head_dim == 96) return launch_pair<96, 128>(a, tq, tkv, d->sm_count, stream);
    return kb64 ? launch_pair<64, 64>(a, tq, tkv, d->sm_count, stream) : launch_pair<64, 128>(a, tq, tkv, d->sm_count, stream);
}

int fame_mask_kv_len(const uint8_t* key_mask, int32_t batch, int32_t seq, int32_t* kv_len, fame_stream_t stream) {
    if (key_mask == nullptr || kv_len == nullptr) return FAME_ERR_NULLPTR;
    if (batch < 0 || seq <= 0) return FAME_ERR_SHAPE;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (batch == 0) return FAME_OK;
    fame::mask_kv_len_kernel<<<(batch + 7) / 8, 256, 0, stream>>>(key_mask, batch, seq, kv_len);
    return launch_status();
}

int fame_attn_cls(const void* q, int64_t ld_q, const void* kv, int64_t ld_kv, int32_t k_col0, int32_t v_col0,
                  const uint8_t* key_mask, void* ctx, int64_t ld_ctx, int32_t batch, int32_t seq, int32_t heads,
                  int32_t head_dim, float scale, fame_stream_t stream) {
    if (q == nullptr || kv == nullptr || ctx == nullptr) return FAME_ERR_NULLPTR;
    if (head_dim != fame::kAcD || seq <= 0 || heads <= 0 || batch < 0 || k_col0 < 0 || v_col0 < 0) return FAME_ERR_SHAPE;
    if (seq > 12000) return FAME_ERR_SHAPE;                        // the probability row lives in shared memory
    const int64_t width = (int64_t)heads * head_dim;
    if (ld_q < width || ld_ctx < width || ld_kv < (k_col0 > v_col0 ? k_col0 : v_col0) + width) return FAME_ERR_SHAPE;
    if ((ld_q & 7) || (ld_kv & 7) || (ld_ctx & 7) || (k_col0 & 7) || (v_col0 & 7)) return FAME_ERR_ALIGN;
    if (!aligned16(q) || !aligned16(kv) || !aligned16(ctx)) return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (batch == 0) return FAME_OK;
    if (batch > 65535) return FAME_ERR_SHAPE;
    fame::AttnClsParams p;
    p.q = reinterpret_cast<const __nv_bfloat16*>(q);
    p.kv = reinterpret_cast<const __nv_bfloat16*>(kv);
    p.key_mask = key_mask;
    p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
    p.ld_q = ld_q; p.ld_kv = ld_kv; p.ld_ctx = ld_ctx;
    p.k_col0 = k_col0; p.v_col0 = v_col0; p.seq = seq; p.heads = heads;
    p.scale_log2e = scale * 1.4426950408889634f;
    fame::attn_cls_kernel<<<dim3(heads, batch), fame::kAcThreads, (size_t)seq * sizeof(float), stream>>>(p);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------ K5
int fame_segment_mean(const fame_segment_mean_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->offsets == nullptr || a->out == nullptr) return FAME_ERR_NULLPTR;
    if (a->patients < 0 || a->cols <= 0 || (a->cols & 7)) return FAME_ERR_SHAPE;
    if (a->x_dtype != FAME_DT_BF16 && a->x_dtype != FAME_DT_F32) return FAME_ERR_SHAPE;
    const int64_t ld_align = a->x_dtype == FAME_DT_BF16 ? 7 : 3;
    if ((a->ldx & ld_align) || !aligned16(a->out) || (a->x != nullptr && !aligned16(a->x))) return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->patients == 0) return FAME_OK;
    if (a->mode != 0 && a->mode != 1) return FAME_ERR_SHAPE;
    const bool bf = a->x_dtype == FAME_DT_BF16, mx = a->mode == 1;
#define FAME_SEG(B, X) \
    fame::segment_reduce_kernel<B, X><<<a->patients, 128, 0, stream>>>(a->x, a->ldx, a->offsets, a->out, a->patients, a->cols)
    if (bf && !mx) FAME_SEG(true, false);
    else if (bf && mx) FAME_SEG(true, true);
    else if (!bf && !mx) FAME_SEG(false, false);
    else FAME_SEG(false, true);
#undef FAME_SEG
    return launch_status();
}

// ------------------------------------------------------------------------------------------------ K4b / K6 / K4c
int fame_lab_embed(const fame_lab_embed_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->lab == nullptr || a->w_tok == nullptr || a->b_tok == nullptr || a->pos == nullptr ||
        a->y == nullptr)
        return FAME_ERR_NULLPTR;
    if (a->batch < 0 || a->L <= 0 || a->hidden <= 0 || (a->hidden & 7)) return FAME_ERR_SHAPE;
    if (!aligned16(a->w_tok) || !aligned16(a->b_tok) || !aligned16(a->pos) || !aligned16(a->y)) return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    const long long tokens = (long long)a->batch * a->L;
    if (tokens == 0) return FAME_OK;
    if (tokens > 0x7fffffffll) return FAME_ERR_SHAPE;
    const int grid = (int)((tokens + fame::kLnWarpsPerBlock - 1) / fame::kLnWarpsPerBlock);
    fame::lab_embed_kernel<<<grid, fame::kLnWarpsPerBlock * 32, 0, stream>>>(
        a->lab, a->w_tok, a->b_tok, a->pos, reinterpret_cast<__nv_bfloat16*>(a->y), (int)tokens, a->L, a->hidden);
    return launch_status();
}

int fame_seq_mean(const fame_seq_mean_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->x == nullptr || a->out == nullptr) return FAME_ERR_NULLPTR;
    if (a->batch < 0 || a->L <= 0 || a->cols <= 0 || (a->cols & 7)) return FAME_ERR_SHAPE;
    if (!aligned16(a->x) || !aligned16(a->out)) return FAME_ERR_ALIGN;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->batch == 0) return FAME_OK;
    if (a->cols > 1024) return FAME_ERR_SHAPE;
    // enough blocks to cover the SMs; the row splits of one sequence form a thread-block cluster (<= 8, portable) and
    // are combined in rank order through distributed shared memory (deterministic, no atomics, no memset)
    int splits = 1;
    while (a->batch * splits < d->sm_count && splits < 8 && splits * 8 <= a->L) splits *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a->batch, splits);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 1;
    attr.val.clusterDim.y = splits;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fame::seq_mean_kernel, reinterpret_cast<const __nv_bfloat16*>(a->x), a->out,
                                       (int)a->L, (int)a->cols, splits);
    if (e != cudaSuccess) return cuda_fail(e);
    return launch_status();
}

int fame_demo_add(const fame_demo_add_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->cls == nullptr || a->out == nullptr) return FAME_ERR_NULLPTR;
    for (int k = 0; k < 4; ++k)
        if (a->ids[k] == nullptr || a->table[k] == nullptr || a->n_rows[k] <= 0) return FAME_ERR_NULLPTR;
    if (a->batch < 0 || a->hidden <= 0) return FAME_ERR_SHAPE;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->batch == 0) return FAME_OK;
    if (a->cls_dtype != FAME_DT_BF16 && a->cls_dtype != FAME_DT_F32) return FAME_ERR_SHAPE;
#define FAME_DEMO_ADD(F)                                                                                              \
    fame::demo_add_kernel<F><<<a->batch, 128, 0, stream>>>(                                                           \
        a->cls, a->ld_cls, reinterpret_cast<const long long*>(a->ids[0]), reinterpret_cast<const long long*>(a->ids[1]), \
        reinterpret_cast<const long long*>(a->ids[2]), reinterpret_cast<const long long*>(a->ids[3]), a->table[0],    \
        a->table[1], a->table[2], a->table[3], a->n_rows[0], a->n_rows[1], a->n_rows[2], a->n_rows[3], a->out, a->hidden)
    if (a->cls_dtype == FAME_DT_F32) FAME_DEMO_ADD(true);
    else FAME_DEMO_ADD(false);
#undef FAME_DEMO_ADD
    return launch_status();
}

static int embed_mean_launch(const fame_embed_mean_args* a, bool backward, fame_stream_t stream) {
    if (a == nullptr) return FAME_ERR_NULLPTR;
    if (a->n_tables < 1 || a->n_tables > fame::kEmMaxTables || a->batch < 0 || a->hidden <= 0) return FAME_ERR_SHAPE;
    if (backward ? a->dout == nullptr : (a->cls == nullptr || a->out == nullptr)) return FAME_ERR_NULLPTR;
    if (!backward && a->cls_dtype != FAME_DT_BF16 && a->cls_dtype != FAME_DT_F32) return FAME_ERR_SHAPE;
    fame::EmbedMeanParams p = {};
    for (int k = 0; k < a->n_tables; ++k) {
        if (a->ids[k] == nullptr || a->n_rows[k] <= 0) return FAME_ERR_NULLPTR;
        if (backward ? a->dtable[k] == nullptr : a->table[k] == nullptr) return FAME_ERR_NULLPTR;
        p.ids[k] = reinterpret_cast<const long long*>(a->ids[k]);
        p.table[k] = a->table[k];
        p.dtable[k] = a->dtable[k];
        p.rows[k] = a->n_rows[k];
    }
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->batch == 0) return FAME_OK;
    p.cls = a->cls; p.ld_cls = a->ld_cls; p.cls_f32 = a->cls_dtype == FAME_DT_F32; p.n_tables = a->n_tables;
    p.out = a->out; p.dout = a->dout; p.hidden = a->hidden;
    if (backward) fame::embed_mean_add_bwd_kernel<<<a->batch, 128, 0, stream>>>(p);
    else fame::embed_mean_add_kernel<<<a->batch, 128, 0, stream>>>(p);
    return launch_status();
}
int fame_embed_mean_add(const fame_embed_mean_args* a, void*, size_t, fame_stream_t stream) {
    return embed_mean_launch(a, false, stream);
}
int fame_embed_mean_add_bwd(const fame_embed_mean_args* a, void*, size_t, fame_stream_t stream) {
    return embed_mean_launch(a, true, stream);
}

// ------------------------------------------------------------------------------------------------ K7
// few patients: 3 launches whose CTAs own output columns (heads.cuh); needs room for proj / gated / pre_relu when the
// caller does not ask for them
static const int kFusionSmallMaxB = 1024;
size_t fame_fusion_fwd_workspace_bytes(int32_t B) {
    return (B > 0 && B <= kFusionSmallMaxB) ? (size_t)B * (768 + 768 + 512) * sizeof(float) : 0;
}

int fame_fusion_fwd(const fame_fusion_fwd_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream) {
    if (a == nullptr || a->logits == nullptr || a->wp_t == nullptr || a->bp == nullptr || a->sig_w == nullptr ||
        a->w3_t == nullptr || a->b3 == nullptr || a->w4 == nullptr || a->b4 == nullptr)
        return FAME_ERR_NULLPTR;
    for (int m = 0; m < 3; ++m)
        if (a->emb[m] == nullptr || !aligned16(a->emb[m])) return a->emb[m] == nullptr ? FAME_ERR_NULLPTR : FAME_ERR_ALIGN;
    if (a->mod_logits != nullptr && (a->wc == nullptr || a->bc == nullptr)) return FAME_ERR_NULLPTR;
    if (a->B < 0) return FAME_ERR_SHAPE;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->B == 0) return FAME_OK;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fame::fusion_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             fame::kFuSmemBytes);
        if (e != cudaSuccess) return cuda_fail(e);
        attr_set[dev] = true;
    }
    fame::FusionParams p;
    for (int m = 0; m < 3; ++m) { p.emb[m] = a->emb[m]; p.w_mod[m] = a->w_mod[m]; }
    p.wp_t = a->wp_t; p.bp = a->bp; p.sig_w = a->sig_w; p.w3_t = a->w3_t; p.b3 = a->b3; p.w4 = a->w4; p.b4 = a->b4;
    p.wc = a->wc; p.bc = a->bc; p.proj = a->proj; p.gated = a->gated; p.pre_relu = a->pre_relu;
    p.logits = a->logits; p.mod_logits = a->mod_logits; p.sig_out = a->sig_out; p.B = a->B;
    p.w_mod_dev = a->w_mod_dev;
    const int grid = (a->B + fame::kFuRows - 1) / fame::kFuRows;
    if (a->B <= kFusionSmallMaxB) {
        float* ws = reinterpret_cast<float*>(workspace);
        const bool need_ws = a->proj == nullptr || a->gated == nullptr || a->pre_relu == nullptr;
        if (need_ws && (ws == nullptr || workspace_bytes < fame_fusion_fwd_workspace_bytes(a->B))) return FAME_ERR_WORKSPACE;
        if (need_ws && !aligned16(ws)) return FAME_ERR_ALIGN;
        float* proj = a->proj != nullptr ? a->proj : ws;
        float* gated = a->gated != nullptr ? a->gated : ws + (size_t)a->B * 768;
        float* pre = a->pre_relu != nullptr ? a->pre_relu : ws + (size_t)a->B * 1536;
        if (!aligned16(proj) || !aligned16(gated) || !aligned16(pre)) return FAME_ERR_ALIGN;
        fame::fusion_small_proj_kernel<<<dim3(24, grid), 256, 0, stream>>>(p, proj, gated);
        fame::fusion_small_hidden_kernel<<<dim3(16, grid), 256, 0, stream>>>(p, gated, pre);
        fame::fusion_small_logits_kernel<<<grid, 256, 0, stream>>>(p, proj, pre);
        return launch_status();
    }
    fame::fusion_fwd_kernel<<<grid, fame::kFuThreads, fame::kFuSmemBytes, stream>>>(p);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------ K8
int fame_loss_stats(const fame_loss_stats_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->logits == nullptr || a->labels == nullptr || a->pos_weight == nullptr ||
        a->stats == nullptr || a->attr[0] == nullptr || a->attr[1] == nullptr || a->attr[2] == nullptr)
        return FAME_ERR_NULLPTR;
    if (a->B < 0) return FAME_ERR_SHAPE;
    static_assert(FAME_LOSS_STATS_LEN == fame::kLossStatsLen, "header / kernel stats layout mismatch");
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->B == 0) return FAME_OK;
    fame::LossStatsParams p;
    p.logits = a->logits; p.labels = a->labels; p.pos_weight = a->pos_weight;
    for (int k = 0; k < 3; ++k) p.attr[k] = reinterpret_cast<const long long*>(a->attr[k]);
    p.stats = reinterpret_cast<long long*>(a->stats);
    p.B = a->B;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fame::loss_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             fame::kLsSmemBytes);
        if (e != cudaSuccess) return cuda_fail(e);
        attr_set[dev] = true;
    }
    const int per_block = fame::kLsThreads * fame::kPatPerThread;   // patients per block per trip
    int grid = (a->B + per_block - 1) / per_block;
    if (grid > 4 * d->sm_count) grid = 4 * d->sm_count;             // 4 resident CTAs per SM (50 KB of bins each)
    fame::loss_stats_kernel<<<grid, fame::kLsThreads, fame::kLsSmemBytes, stream>>>(p);
    return launch_status();
}

int fame_loss_fwd_bwd(const fame_loss_fwd_bwd_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->stats == nullptr || a->pos_weight == nullptr) return FAME_ERR_NULLPTR;
    if (a->dlogits != nullptr && (a->logits == nullptr || a->labels == nullptr || a->attr[0] == nullptr ||
                                  a->attr[1] == nullptr || a->attr[2] == nullptr))
        return FAME_ERR_NULLPTR;
    if (a->B < 0 || a->n_sig < 0) return FAME_ERR_SHAPE;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    fame::LossGradParams p;
    p.logits = a->logits; p.labels = a->labels; p.pos_weight = a->pos_weight;
    for (int k = 0; k < 3; ++k) p.attr[k] = reinterpret_cast<const long long*>(a->attr[k]);
    p.stats = reinterpret_cast<const long long*>(a->stats);
    p.sig_w = a->sig_w; p.n_sig = a->n_sig; p.lambda_edd = a->lambda_edd; p.lambda_l1 = a->lambda_l1;
    p.dlogits = a->dlogits; p.loss_out = a->loss_out; p.B = a->B;
    int grid = (a->B + 1023) / 1024;     // 256 threads x 4 patients per trip
    if (grid < 1) grid = 1;
    if (grid > 4 * d->sm_count) grid = 4 * d->sm_count;
    fame::loss_fwd_bwd_kernel<<<grid, 256, 0, stream>>>(p);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------ K10
int fame_eval_counts(const fame_eval_counts_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->logits == nullptr || a->labels == nullptr || a->out == nullptr || a->attr[0] == nullptr ||
        a->attr[1] == nullptr || a->attr[2] == nullptr)
        return FAME_ERR_NULLPTR;
    if (a->N < 0 || a->ld < 3) return FAME_ERR_SHAPE;
    static_assert(FAME_EVAL_COUNTS_LEN == fame::kEvLen, "header / kernel count layout mismatch");
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->N == 0) return FAME_OK;
    fame::EvalCountsParams p;
    p.logits = a->logits; p.ld = a->ld; p.labels = a->labels;
    for (int k = 0; k < 3; ++k) { p.attr[k] = reinterpret_cast<const long long*>(a->attr[k]); p.thr[k] = a->thr[k]; }
    p.sweep = a->sweep;
    p.out = reinterpret_cast<unsigned long long*>(a->out);
    p.N = a->N;
    p.logits_are_probs = a->logits_are_probs;
    int grid = (a->N + 1023) / 1024;     // 256 threads x 4 patients per trip
    if (grid > 6 * d->sm_count) grid = 6 * d->sm_count;   // 70 registers: 3 resident CTAs per SM, 2 waves
    fame::eval_counts_kernel<<<grid, 256, 0, stream>>>(p);
    return launch_status();
}

static inline size_t rank_npad(int32_t n) {
    size_t p = fame::kSortTile;
    while (p < (size_t)n) p <<= 1;
    return p;
}
// brute-force sub-range: per-block float64 partials; full range: padded key array + prefix-positives array
size_t fame_rank_counts_workspace_bytes(int32_t n_i) {
    const size_t brute = sizeof(double) * (size_t)((n_i + 255) / 256 + 1);
    const size_t sorted = sizeof(uint32_t) * (rank_npad(n_i) + (size_t)n_i);
    return brute > sorted ? brute : sorted;
}

int fame_rank_counts(const fame_rank_counts_args* a, void* workspace, size_t workspace_bytes, fame_stream_t stream) {
    if (a == nullptr || a->scores == nullptr || a->y == nullptr || a->auroc2 == nullptr || a->ap_sum == nullptr ||
        a->npos_nneg == nullptr)
        return FAME_ERR_NULLPTR;
    if (a->N < 0 || a->i0 < 0 || a->i1 < a->i0 || a->i1 > a->N) return FAME_ERR_SHAPE;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    const int n_i = a->i1 - a->i0;
    if (n_i == 0) return FAME_OK;
    if (workspace == nullptr) return FAME_ERR_NULLPTR;
    if (workspace_bytes < fame_rank_counts_workspace_bytes(n_i)) return FAME_ERR_WORKSPACE;
    if (a->i0 == 0 && a->i1 == a->N) {
        // full range: sort + tie-run scan, O(N log N)
        const size_t npad = rank_npad(a->N);
        if (npad > (size_t)1 << 30) return FAME_ERR_SHAPE;
        uint32_t* keys = reinterpret_cast<uint32_t*>(workspace);
        uint32_t* cpos = keys + npad;
        fame::rank_keys_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, stream>>>(a->scores, a->y, a->N, (int)npad, keys);
        const unsigned tiles = (unsigned)(npad / fame::kSortTile);
        fame::bitonic_smem_kernel<<<tiles, 1024, 0, stream>>>(keys, 2, fame::kSortTile, fame::kSortTile / 2);
        for (size_t k = 2 * (size_t)fame::kSortTile; k <= npad; k <<= 1) {
            for (size_t j = k >> 1; j >= (size_t)fame::kSortTile; j >>= 1)
                fame::bitonic_global_kernel<<<(unsigned)((npad / 2 + 255) / 256), 256, 0, stream>>>(keys, (int)npad, (int)k,
                                                                                               (int)j);
            fame::bitonic_smem_kernel<<<tiles, 1024, 0, stream>>>(keys, (int)k, (int)k, fame::kSortTile / 2);
        }
        fame::rank_scan_kernel<<<1, 1024, 0, stream>>>(keys, a->N, cpos, reinterpret_cast<unsigned long long*>(a->auroc2),
                                                       a->ap_sum, reinterpret_cast<unsigned long long*>(a->npos_nneg));
        return launch_status();
    }
    // a sub-range of i against all j (a caller that shards i across ranks): exact O(N^2 / ranks) compare
    fame::RankParams p;
    p.scores = a->scores; p.y = a->y; p.N = a->N; p.i0 = a->i0; p.i1 = a->i1;
    p.auroc2 = reinterpret_cast<unsigned long long*>(a->auroc2);
    p.ap_partial = reinterpret_cast<double*>(workspace);
    p.npos_nneg = reinterpret_cast<unsigned long long*>(a->npos_nneg);
    const int grid = (n_i + 255) / 256;
    fame::rank_counts_kernel<<<grid, 256, 0, stream>>>(p);
    fame::sum_partials_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<const double*>(workspace), grid, a->ap_sum);
    return launch_status();
}

#include <math.h>
#include "train_abi.inc"

int fame_sigmoid_probs(const fame_sigmoid_probs_args* a, void*, size_t, fame_stream_t stream) {
    if (a == nullptr || a->logits == nullptr || a->probs == nullptr) return FAME_ERR_NULLPTR;
    if (a->y8 != nullptr && a->labels == nullptr) return FAME_ERR_NULLPTR;
    if (a->N < 0 || a->ld < 3) return FAME_ERR_SHAPE;
    DeviceInfo* d = nullptr;
    int rc = device_info(&d);
    if (rc != FAME_OK) return rc;
    if (a->N == 0) return FAME_OK;
    fame::sigmoid_probs_kernel<<<(a->N + 255) / 256, 256, 0, stream>>>(a->logits, a->ld, a->labels, a->probs, a->y8, a->N);
    return launch_status();
}

}  // extern "C"

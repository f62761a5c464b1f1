// HBM-bound row kernels: LayerNorm (K3), BERT embedding gather + LN (K4), CLS-gather + segmented
// chunk->patient mean (K5).  One warp per row, 16-byte vector loads, warp-shuffle reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fame {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- patient-record loader shared by the loss-statistics and evaluation-count kernels -------------------------------
// One thread takes kPatPerThread = 4 CONSECUTIVE patients per trip: 12 floats of logits and of labels (3 x 16-byte
// loads each when the rows are dense: row stride 3 floats) and 4 int64 codes per sensitive attribute (2 x 16-byte
// loads each).  12 independent 16-byte loads per thread keep ~49 KB per SM in flight at 256 threads per SM, which is
// what HBM3e needs (bandwidth x latency / 148 SMs ~ 44 KB); one patient per trip left these kernels latency bound
// at 15-24 % of the copy bandwidth.
constexpr int kPatPerThread = 4;

struct PatientQuad {
    float z[kPatPerThread][3];
    float y[kPatPerThread][3];
    long long code[3][kPatPerThread];
    int n;   // patients valid in this quad (0..4)
};

__device__ __forceinline__ void load_patient_quad(PatientQuad& q, const float* __restrict__ logits, long long ld,
                                                  const float* __restrict__ labels,
                                                  const long long* const (&attr)[3], long long b0, long long N,
                                                  bool vec_ok) {
    q.n = (int)(N - b0 < kPatPerThread ? (N - b0 < 0 ? 0 : N - b0) : kPatPerThread);
    if (q.n == kPatPerThread && vec_ok) {
        const float4* zl = reinterpret_cast<const float4*>(logits + 3 * b0);
        const float4* yl = reinterpret_cast<const float4*>(labels + 3 * b0);
        float4 zv[3], yv[3];
        longlong2 cv[3][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            zv[k] = __ldg(zl + k);
            yv[k] = __ldg(yl + k);
            cv[k][0] = __ldg(reinterpret_cast<const longlong2*>(attr[k] + b0));
            cv[k][1] = __ldg(reinterpret_cast<const longlong2*>(attr[k] + b0) + 1);
        }
        const float* zf = reinterpret_cast<const float*>(zv);
        const float* yf = reinterpret_cast<const float*>(yv);
#pragma unroll
        for (int u = 0; u < kPatPerThread; ++u)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                q.z[u][i] = zf[3 * u + i];
                q.y[u][i] = yf[3 * u + i];
            }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            q.code[k][0] = cv[k][0].x; q.code[k][1] = cv[k][0].y;
            q.code[k][2] = cv[k][1].x; q.code[k][3] = cv[k][1].y;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kPatPerThread; ++u) {
            const bool live = u < q.n;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                q.z[u][i] = live ? __ldg(logits + (b0 + u) * ld + i) : 0.f;
                q.y[u][i] = live ? __ldg(labels + 3 * (b0 + u) + i) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) q.code[k][u] = live ? __ldg(attr[k] + b0 + u) : 0;
        }
    }
}

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 v = __bfloat1622float2(h[t]);
        f[2 * t] = v.x;
        f[2 * t + 1] = v.y;
    }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float* f) {
    uint4 o;
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]);
    __nv_bfloat162 d = __floats2bfloat162_rn(f[6], f[7]);
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    o.z = *reinterpret_cast<uint32_t*>(&c);
    o.w = *reinterpret_cast<uint32_t*>(&d);
    return o;
}

// ---------------------------------------------------------------------------------------- K3 LayerNorm
// cols <= 1024, cols % 8 == 0.  Lane l owns the 8-element chunks l, l+32, l+64, l+96.
// Input bf16 or f32; outputs bf16 and/or f32 (either may be null); optional per-row {mean, rstd} for the backward.
constexpr int kLnWarpsPerBlock = 8;
constexpr int kLnMaxChunks = 4;

template <bool kInF32>
__device__ __forceinline__ void ln_load8(const void* __restrict__ xr, int ch, float* f) {
    if (kInF32) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xr) + 2 * ch);
        const float4 b = __ldg(reinterpret_cast<const float4*>(xr) + 2 * ch + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(xr) + ch), f);
    }
}

template <bool kInF32>
__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
layernorm_kernel(const void* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32,
                 long long ldy, float2* __restrict__ stats, int rows, int cols, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kLnWarpsPerBlock + warp;
    if (row >= rows) return;
    const int nchunks = cols >> 3;
    const void* xr = kInF32 ? static_cast<const void*>(reinterpret_cast<const float*>(x) + (long long)row * ldx)
                            : static_cast<const void*>(reinterpret_cast<const __nv_bfloat16*>(x) + (long long)row * ldx);
    float v[kLnMaxChunks][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxChunks; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
            ln_load8<kInF32>(xr, ch, v[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        }
    }
    const float mean = warp_sum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxChunks; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                q += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
    if (stats != nullptr && lane == 0) stats[row] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < kLnMaxChunks; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch);
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * ch);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * ch + 1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
            if (y_bf16 != nullptr) reinterpret_cast<uint4*>(y_bf16 + (long long)row * ldy)[ch] = float_to_bf16x8(o);
            if (y_f32 != nullptr) {
                float4* yo = reinterpret_cast<float4*>(y_f32 + (long long)row * ldy) + 2 * ch;
                yo[0] = make_float4(o[0], o[1], o[2], o[3]);
                yo[1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

// bf16 -> bf16 fast path (the 24 LayerNorms of a note-encoder step: 201 MB in, 201 MB out each, HBM bound).
// A warp owns TWO rows at a time and keeps them as packed bf16 (12 + 12 registers at 768 columns) instead of 32 + 32
// floats: ~40 registers per thread => 48 resident warps per SM x 2 rows x 1.5 KB = 147 KB of loads in flight per SM
// (the one-row kernel above had 32 warps x 1.5 KB = 49 KB, below what 6.5 TB/s x ~1.5 us of loaded latency needs),
// and the two rows' shuffle reductions interleave.  Two-pass variance on the register copy, as above.
constexpr int kLn2WarpsPerBlock = 8;

// volatile: each pass re-expands the packed row copy; without it the compiler keeps all 48 expanded floats alive
// across the reductions and spills them
__device__ __forceinline__ float bf16_lo(uint32_t w) {
    float f;
    asm volatile("shl.b32 %0, %1, 16;" : "=f"(f) : "r"(w));
    return f;
}
__device__ __forceinline__ float bf16_hi(uint32_t w) {
    float f;
    asm volatile("and.b32 %0, %1, 0xffff0000;" : "=f"(f) : "r"(w));
    return f;
}
__device__ __forceinline__ float sum8(const uint4& u) {
    return ((bf16_lo(u.x) + bf16_hi(u.x)) + (bf16_lo(u.y) + bf16_hi(u.y))) +
           ((bf16_lo(u.z) + bf16_hi(u.z)) + (bf16_lo(u.w) + bf16_hi(u.w)));
}
__device__ __forceinline__ float sq(float d) { return d * d; }
__device__ __forceinline__ float sqdev8(const uint4& u, float m) {
    return ((sq(bf16_lo(u.x) - m) + sq(bf16_hi(u.x) - m)) + (sq(bf16_lo(u.y) - m) + sq(bf16_hi(u.y) - m))) +
           ((sq(bf16_lo(u.z) - m) + sq(bf16_hi(u.z) - m)) + (sq(bf16_lo(u.w) - m) + sq(bf16_hi(u.w) - m)));
}
__device__ __forceinline__ uint32_t ln_out2(uint32_t w, float m, float rs, float g0, float g1, float b0, float b1) {
    return pack_bf16x2((bf16_lo(w) - m) * rs * g0 + b0, (bf16_hi(w) - m) * rs * g1 + b1);
}

// bf16(a + b) of two packed rows chunks, summed in f32 (same rounding as the GEMM epilogue's residual add)
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
    return pack_bf16x2(bf16_lo(a) + bf16_lo(b), bf16_hi(a) + bf16_hi(b));
}
__device__ __forceinline__ uint4 add_bf16x8(const uint4& a, const uint4& b) {
    return make_uint4(add_bf16x2(a.x, b.x), add_bf16x2(a.y, b.y), add_bf16x2(a.z, b.z), add_bf16x2(a.w, b.w));
}

// kRes: y = LN(x + residual).  The residual add of  LN(dense(.) + x)  (HF:297, 355) normally rides in the producing
// GEMM's epilogue; for the attention-output projection (K = N = 768: 6144 clk of MMAs per tile) the epilogue's
// scattered 16-byte residual loads were its critical path (tensor pipe 33 % active, stall reason long_scoreboard),
// while this streaming kernel absorbs the same 201 MB at HBM speed.
template <int CH, bool kRes = false>   // chunks of 8 columns per lane: cols <= CH * 256
__global__ void __launch_bounds__(kLn2WarpsPerBlock * 32, kRes ? 4 : 5)
layernorm_bf16_rows2_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                            const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, long long ldy,
                            float2* __restrict__ stats, int rows, int cols, float eps,
                            const __nv_bfloat16* __restrict__ res = nullptr, long long ldr = 0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * kLn2WarpsPerBlock + warp) * 2;
    if (row0 >= rows) return;
    const bool two = row0 + 1 < rows;
    const int nchunks = cols >> 3;
    const uint4* xa = reinterpret_cast<const uint4*>(x + (long long)row0 * ldx);
    const uint4* xb = reinterpret_cast<const uint4*>(x + (long long)(row0 + (two ? 1 : 0)) * ldx);
    uint4 ra[CH], rb[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
            ra[i] = ld_nc_na(xa + ch);
            rb[i] = ld_nc_na(xb + ch);
        } else {
            ra[i] = make_uint4(0, 0, 0, 0);
            rb[i] = make_uint4(0, 0, 0, 0);
        }
    }
    if (kRes) {
        const uint4* qa = reinterpret_cast<const uint4*>(res + (long long)row0 * ldr);
        const uint4* qb = reinterpret_cast<const uint4*>(res + (long long)(row0 + (two ? 1 : 0)) * ldr);
        uint4 sa4[CH], sb4[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int ch = lane + 32 * i;
            if (ch < nchunks) {
                sa4[i] = ld_nc_na(qa + ch);
                sb4[i] = ld_nc_na(qb + ch);
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (lane + 32 * i < nchunks) {
                ra[i] = add_bf16x8(ra[i], sa4[i]);
                rb[i] = add_bf16x8(rb[i], sb4[i]);
            }
        }
    }
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        sa += sum8(ra[i]);
        sb += sum8(rb[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    const float inv_n = 1.0f / (float)cols;
    const float ma = sa * inv_n, mb = sb * inv_n;
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        if (lane + 32 * i < nchunks) {      // the zero padding of absent chunks must not enter the variance
            qa += sqdev8(ra[i], ma);
            qb += sqdev8(rb[i], mb);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qa += __shfl_xor_sync(0xffffffffu, qa, o);
        qb += __shfl_xor_sync(0xffffffffu, qb, o);
    }
    const float rsa = rsqrtf(qa * inv_n + eps), rsb = rsqrtf(qb * inv_n + eps);
    if (stats != nullptr && lane == 0) {
        stats[row0] = make_float2(ma, rsa);
        if (two) stats[row0 + 1] = make_float2(mb, rsb);
    }
    uint4* ya = reinterpret_cast<uint4*>(y + (long long)row0 * ldy);
    uint4* yb = reinterpret_cast<uint4*>(y + (long long)(row0 + 1) * ldy);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
            const float4 g0 = ld_nc_f4_pinned(reinterpret_cast<const float4*>(gamma) + 2 * ch);
            const float4 b0 = ld_nc_f4_pinned(reinterpret_cast<const float4*>(beta) + 2 * ch);
            uint4 oa, ob;
            oa.x = ln_out2(ra[i].x, ma, rsa, g0.x, g0.y, b0.x, b0.y);
            oa.y = ln_out2(ra[i].y, ma, rsa, g0.z, g0.w, b0.z, b0.w);
            ob.x = ln_out2(rb[i].x, mb, rsb, g0.x, g0.y, b0.x, b0.y);
            ob.y = ln_out2(rb[i].y, mb, rsb, g0.z, g0.w, b0.z, b0.w);
            const float4 g1 = ld_nc_f4_pinned(reinterpret_cast<const float4*>(gamma) + 2 * ch + 1);
            const float4 b1 = ld_nc_f4_pinned(reinterpret_cast<const float4*>(beta) + 2 * ch + 1);
            oa.z = ln_out2(ra[i].z, ma, rsa, g1.x, g1.y, b1.x, b1.y);
            oa.w = ln_out2(ra[i].w, ma, rsa, g1.z, g1.w, b1.z, b1.w);
            ob.z = ln_out2(rb[i].z, mb, rsb, g1.x, g1.y, b1.x, b1.y);
            ob.w = ln_out2(rb[i].w, mb, rsb, g1.z, g1.w, b1.z, b1.w);
            ya[ch] = oa;
            if (two) yb[ch] = ob;
        }
    }
}

// |kv_len[b]| = 1 + index of the last non-zero byte of key_mask[b, :]  (0 for an all-zero row); warp per sequence.
// The sign says whether the mask is a PREFIX mask (every key before the last attended one is attended too -- what a
// padded tokenizer batch produces): kv_len >= 0 then, and the attention kernel derives key validity from the length
// alone; a mask with holes gets -last and the kernel reads the mask bytes.
__global__ void __launch_bounds__(256)
mask_kv_len_kernel(const uint8_t* __restrict__ key_mask, int batch, int seq, int* __restrict__ kv_len) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= batch) return;
    int last = 0, ones = 0;
    for (int k = lane; k < seq; k += 32)
        if (key_mask[(long long)b * seq + k] != 0) {
            last = k + 1;
            ++ones;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        ones += __shfl_xor_sync(0xffffffffu, ones, o);
    }
    if (lane == 0) kv_len[b] = ones == last ? last : -last;
}

// ---------------------------------------------------------------------------------------- K4 BERT embedding
// hidden % 128 == 0, hidden <= 1024: lane l owns float4 chunks l, l+32, ... (hidden/128 of them).
constexpr int kEmbMaxChunks = 8;

__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
bert_embed_ln_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                     const float* __restrict__ pos, const float* __restrict__ type0,
                     const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                     float* __restrict__ y_f32, float* __restrict__ sum_out, float2* __restrict__ stats,
                     int* __restrict__ err_flag, int tokens, int seq_len, int hidden, int vocab, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kLnWarpsPerBlock + warp;
    if (t >= tokens) return;
    long long id = ids[t];
    if (id < 0 || id >= vocab) {
        if (err_flag != nullptr && lane == 0) atomicExch(err_flag, 1);
        id = id < 0 ? 0 : vocab - 1;
    }
    const int nchunks = hidden >> 7;  // float4 chunks per lane
    const float4* wr = reinterpret_cast<const float4*>(word + id * hidden);
    const float4* pr = reinterpret_cast<const float4*>(pos + (long long)(t % seq_len) * hidden);
    const float4* tr = reinterpret_cast<const float4*>(type0);
    float4 v[kEmbMaxChunks];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kEmbMaxChunks; ++i) {
        if (i < nchunks) {
            const int ch = lane + 32 * i;
            const float4 w = __ldg(wr + ch), ty = __ldg(tr + ch), po = __ldg(pr + ch);
            // same association as HF: (inputs_embeds + token_type_embeddings) + position_embeddings
            v[i].x = (w.x + ty.x) + po.x;
            v[i].y = (w.y + ty.y) + po.y;
            v[i].z = (w.z + ty.z) + po.z;
            v[i].w = (w.w + ty.w) + po.w;
            s += v[i].x + v[i].y + v[i].z + v[i].w;
            if (sum_out != nullptr) reinterpret_cast<float4*>(sum_out + (long long)t * hidden)[ch] = v[i];
        }
    }
    const float mean = warp_sum(s) / (float)hidden;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kEmbMaxChunks; ++i) {
        if (i < nchunks) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += a * a + b * b + c * c + d * d;
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)hidden + eps);
    if (stats != nullptr && lane == 0) stats[t] = make_float2(mean, rstd);
    __nv_bfloat16* yr = y + (long long)t * hidden;
#pragma unroll
    for (int i = 0; i < kEmbMaxChunks; ++i) {
        if (i < nchunks) {
            const int ch = lane + 32 * i;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + ch);
            const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
            const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
            if (y_f32 != nullptr)
                reinterpret_cast<float4*>(y_f32 + (long long)t * hidden)[ch] = make_float4(o0, o1, o2, o3);
            __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1);
            __nv_bfloat162 hi = __floats2bfloat162_rn(o2, o3);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(yr)[ch] = o;
        }
    }
}

// ---------------------------------------------------------------------------------------- K5 segment mean / max
// One block per patient; thread t owns columns [8t, 8t+8).  Rows (chunks) are streamed 4 at a time.
// kMax = false: arithmetic mean (aggregation="mean", 10_FAME.py:171); true: column-wise max (the "max" branch).
template <bool kBf16>
__device__ __forceinline__ void load_row8(const void* __restrict__ x, long long off, float* f) {
    if (kBf16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + off));
        bf16x8_to_float(v, f);
    } else {
        const float* xr = reinterpret_cast<const float*>(x) + off;
        const float4 a = __ldg(reinterpret_cast<const float4*>(xr));
        const float4 b = __ldg(reinterpret_cast<const float4*>(xr) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
}

template <bool kBf16, bool kMax>
__global__ void __launch_bounds__(128)
segment_reduce_kernel(const void* __restrict__ x, long long ldx, const int* __restrict__ offsets,
                      float* __restrict__ out, int patients, int cols) {
    const int p = blockIdx.x;
    if (p >= patients) return;
    const int beg = __ldg(offsets + p), end = __ldg(offsets + p + 1);
    const int n = end - beg;
    for (int c0 = threadIdx.x * 8; c0 < cols; c0 += blockDim.x * 8) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = (kMax && n > 0) ? -INFINITY : 0.f;
        int r = beg;
        for (; r + 4 <= end; r += 4) {
            float f[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) load_row8<kBf16>(x, (long long)(r + u) * ldx + c0, f[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = kMax ? fmaxf(acc[j], f[u][j]) : acc[j] + f[u][j];
        }
        for (; r < end; ++r) {
            float f[8];
            load_row8<kBf16>(x, (long long)r * ldx + c0, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = kMax ? fmaxf(acc[j], f[j]) : acc[j] + f[j];
        }
        // mean: sum-then-divide with rows added in order == np.mean(axis=0) on f32 rows, bit for bit; n == 0 -> zeros
        const float den = (!kMax && n > 0) ? (float)n : 1.0f;
        float* o = out + (long long)p * cols + c0;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0] / den, acc[1] / den, acc[2] / den, acc[3] / den);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] / den, acc[5] / den, acc[6] / den, acc[7] / den);
    }
}

// ---------------------------------------------------------------------------------------- K4b lab embedding
// x[b*L + l, :] = lab[b, l] * w_tok + b_tok + pos[l, :]   (BEHRTModel_Lab.forward, 10_FAME.py:218-220) -> bf16.
// One warp per token; hidden % 256 == 0 (lane owns 8-column chunks lane, lane+32, ...), hidden <= 1024.
__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
lab_embed_kernel(const float* __restrict__ lab, const float* __restrict__ w_tok, const float* __restrict__ b_tok,
                 const float* __restrict__ pos, __nv_bfloat16* __restrict__ y, int tokens, int L, int hidden) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kLnWarpsPerBlock + warp;
    if (t >= tokens) return;
    const float v = __ldg(lab + t);
    const float* pr = pos + (long long)(t % L) * hidden;
    __nv_bfloat16* yr = y + (long long)t * hidden;
    for (int ch = lane; ch < (hidden >> 3); ch += 32) {
        float o[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(w_tok) + 2 * ch + h);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b_tok) + 2 * ch + h);
            const float4 pp = __ldg(reinterpret_cast<const float4*>(pr) + 2 * ch + h);
            // nn.Linear(1, H): x * w + b in one rounding per op as torch does (mul, add bias, add pos)
            o[4 * h + 0] = (v * w.x + bb.x) + pp.x;
            o[4 * h + 1] = (v * w.y + bb.y) + pp.y;
            o[4 * h + 2] = (v * w.z + bb.z) + pp.z;
            o[4 * h + 3] = (v * w.w + bb.w) + pp.w;
        }
        reinterpret_cast<uint4*>(yr)[ch] = float_to_bf16x8(o);
    }
}

// ---------------------------------------------------------------------------------------- K6 sequence mean
// out[b, :] = mean_l x[b*L + l, :]  (BEHRTModel_Lab.forward, 10_FAME.py:224).  bf16 in, f32 out.
// grid = (batch, splits), launched as thread-block clusters (1, splits, 1): block (b, s) sums rows l = s, s + splits,
// ...; the partial rows meet in the distributed shared memory of the cluster and rank 0 adds them IN RANK ORDER, so
// the result does not depend on block timing.  (Float atomics here made the whole training step irreproducible: the
// 1e-7 jitter of this mean reaches the loss gradient and the demographic tower -- a 12-layer post-LN BERT fed a
// constant token -- amplifies it to 1e-3 in its weight gradients.)  cols <= 1024, cols % 8 == 0.
__device__ __forceinline__ float ld_shared_cluster_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(128)
seq_mean_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int L, int cols, int splits) {
    __shared__ float part[1024];
    const int b = blockIdx.x, sp = blockIdx.y;
    const __nv_bfloat16* xb = x + (long long)b * L * cols;
    const float inv = 1.0f / (float)L;
    const int c0 = threadIdx.x * 8;           // 128 threads x 8 columns cover cols <= 1024
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (c0 < cols) {
        int l = sp;
        for (; l + 7 * splits < L; l += 8 * splits) {      // 8 independent 16-byte loads in flight per thread
            uint4 raw[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                raw[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(l + u * splits) * cols + c0));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float f[8];
                bf16x8_to_float(raw[u], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        }
        for (; l < L; l += splits) {
            float f[8];
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(xb + (long long)l * cols + c0)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
    if (splits == 1) {
        if (c0 < cols) {
            float* o = out + (long long)b * cols + c0;
            *reinterpret_cast<float4*>(o) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
            *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[c0 + j] = acc[j];
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (sp == 0 && c0 < cols) {
        float tot[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tot[j] = acc[j];
        for (int r = 1; r < splits; ++r) {
            uint32_t remote;
            const uint32_t local = static_cast<uint32_t>(__cvta_generic_to_shared(&part[c0]));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
#pragma unroll
            for (int j = 0; j < 8; ++j) tot[j] += ld_shared_cluster_f32(remote + 4 * j);
        }
        float* o = out + (long long)b * cols + c0;
        *reinterpret_cast<float4*>(o) = make_float4(tot[0] * inv, tot[1] * inv, tot[2] * inv, tot[3] * inv);
        *reinterpret_cast<float4*>(o + 4) = make_float4(tot[4] * inv, tot[5] * inv, tot[6] * inv, tot[7] * inv);
    }
    // nobody leaves while rank 0 may still read its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------- K4c demographic add
// out[b, :] = cls[b*ld_cls : , :] + (E_age[clamp(age[b])] + E_gen[..] + E_eth[..] + E_ins[..]) / 4
// (BEHRTModel_Demo.forward, 10_FAME.py:195-206).  cls bf16 (CLS row of the demo BERT), tables f32, out f32.
template <bool kClsF32>
__global__ void __launch_bounds__(128)
demo_add_kernel(const void* __restrict__ cls, long long ld_cls, const long long* __restrict__ age,
                const long long* __restrict__ gen, const long long* __restrict__ eth, const long long* __restrict__ ins,
                const float* __restrict__ e_age, const float* __restrict__ e_gen, const float* __restrict__ e_eth,
                const float* __restrict__ e_ins, int n_age, int n_gen, int n_eth, int n_ins, float* __restrict__ out,
                int hidden) {
    const int b = blockIdx.x;
    auto clampi = [](long long v, int n) { return (int)(v < 0 ? 0 : (v > n - 1 ? n - 1 : v)); };
    const float* ra = e_age + (long long)clampi(age[b], n_age) * hidden;
    const float* rg = e_gen + (long long)clampi(gen[b], n_gen) * hidden;
    const float* re = e_eth + (long long)clampi(eth[b], n_eth) * hidden;
    const float* ri = e_ins + (long long)clampi(ins[b], n_ins) * hidden;
    for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
        const float extra = (((ra[c] + rg[c]) + re[c]) + ri[c]) / 4.0f;
        const float cv = kClsF32 ? reinterpret_cast<const float*>(cls)[(long long)b * ld_cls + c]
                                 : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(cls)[(long long)b * ld_cls + c]);
        out[(long long)b * hidden + c] = cv + extra;
    }
}

}  // namespace fame

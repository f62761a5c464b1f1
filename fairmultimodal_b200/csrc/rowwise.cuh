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
constexpr int kLnWarpsPerBlock = 8;
constexpr int kLnMaxChunks = 4;

__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
layernorm_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, long long ldy, int rows,
                      int cols, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kLnWarpsPerBlock + warp;
    if (row >= rows) return;
    const int nchunks = cols >> 3;
    const __nv_bfloat16* xr = x + (long long)row * ldx;
    float v[kLnMaxChunks][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxChunks; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr) + ch);
            bf16x8_to_float(u, v[i]);
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
    __nv_bfloat16* yr = y + (long long)row * ldy;
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
            reinterpret_cast<uint4*>(yr)[ch] = float_to_bf16x8(o);
        }
    }
}

// ---------------------------------------------------------------------------------------- K4 BERT embedding
// hidden % 128 == 0, hidden <= 1024: lane l owns float4 chunks l, l+32, ... (hidden/128 of them).
constexpr int kEmbMaxChunks = 8;

__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
bert_embed_ln_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                     const float* __restrict__ pos, const float* __restrict__ type0,
                     const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
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
    __nv_bfloat16* yr = y + (long long)t * hidden;
#pragma unroll
    for (int i = 0; i < kEmbMaxChunks; ++i) {
        if (i < nchunks) {
            const int ch = lane + 32 * i;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + ch);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + ch);
            __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x,
                                                      (v[i].y - mean) * rstd * g.y + b.y);
            __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + b.z,
                                                      (v[i].w - mean) * rstd * g.w + b.w);
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

}  // namespace fame

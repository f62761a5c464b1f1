// K2c: attention for ONE query row per sequence (the [CLS] token), head_dim 64.
//
// The reference's note encoder consumes only last_hidden_state[:, 0, :] (10_FAME.py:141), so in the LAST encoder
// layer every query but the first of each chunk is dead work: the layer needs K and V of all 512 tokens but the
// scores, the softmax, P.V and everything behind the attention only for the CLS row (HF modeling_bert.py:192-206
// evaluated for one query).  Per (chunk, head) that is a 1 x S x 64 product pair: 0.26 MFLOP against 128 KB of K / V
// -- an HBM-bound streaming kernel, not tensor-core work (256 chunks: 403 MB of K / V).
//
//   grid = (heads, batch), 128 threads.
//   phase 1: thread t scores keys t, t + 128, ...: one 128-byte K row each, dot with q kept in registers
//   phase 2: block max / sum (softmax over the attended keys; masked keys get probability exactly 0)
//   phase 3: warp w accumulates P.V over keys w, w + 4, ...: each V row is one coalesced 128-byte warp load,
//            lane l owns dims 2l, 2l + 1; the four partial sums meet in shared memory.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sm100_ptx.cuh"
#include "rowwise.cuh"

namespace fame {

constexpr int kAcThreads = 128;
constexpr int kAcD = 64;

struct AttnClsParams {
    const __nv_bfloat16* q;    // [batch, heads * 64]            (row stride ld_q)
    const __nv_bfloat16* kv;   // [batch * seq, ...]: K of head h at column k_col0 + 64 h, V at v_col0 + 64 h
    const uint8_t* key_mask;   // [batch, seq] (1 = attend) or nullptr
    __nv_bfloat16* ctx;        // [batch, heads * 64]            (row stride ld_ctx)
    long long ld_q, ld_kv, ld_ctx;
    int k_col0, v_col0, seq, heads;
    float scale_log2e;
};

__global__ void __launch_bounds__(kAcThreads)
attn_cls_kernel(const AttnClsParams p) {
    extern __shared__ float ac_smem[];
    float* prob = ac_smem;                    // [seq]
    __shared__ float red[4];
    __shared__ float part[4][kAcD];
    const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int S = p.seq;
    // q (scaled) in registers: 8 x uint4 of bf16
    float q[kAcD];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(p.q + (long long)b * p.ld_q + h * kAcD);
#pragma unroll
        for (int c = 0; c < kAcD / 8; ++c) {
            const uint4 u = __ldg(qp + c);
            bf16x8_to_float(u, q + 8 * c);
        }
#pragma unroll
        for (int d = 0; d < kAcD; ++d) q[d] *= p.scale_log2e;
    }
    const __nv_bfloat16* kbase = p.kv + (long long)b * S * p.ld_kv + p.k_col0 + h * kAcD;
    const __nv_bfloat16* vbase = p.kv + (long long)b * S * p.ld_kv + p.v_col0 + h * kAcD;
    const uint8_t* mask = p.key_mask != nullptr ? p.key_mask + (long long)b * S : nullptr;
    float mx = -INFINITY;
    for (int k = t; k < S; k += kAcThreads) {
        float s = -INFINITY;
        if (mask == nullptr || mask[k] != 0) {
            const uint4* kp = reinterpret_cast<const uint4*>(kbase + (long long)k * p.ld_kv);
            uint4 u[kAcD / 8];
#pragma unroll
            for (int c = 0; c < kAcD / 8; ++c) u[c] = __ldg(kp + c);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < kAcD / 8; ++c) {
                float f[8];
                bf16x8_to_float(u[c], f);
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    a0 = fmaf(q[8 * c + i], f[i], a0);
                    a1 = fmaf(q[8 * c + i + 1], f[i + 1], a1);
                }
            }
            s = a0 + a1;
        }
        prob[k] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    const float m_safe = mx == -INFINITY ? 0.f : mx;
    float sum = 0.f;
    for (int k = t; k < S; k += kAcThreads) {
        const float e = exp2f(prob[k] - m_safe);      // masked: exp2(-inf) = 0
        prob[k] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncthreads();                                   // red[] reads above are done, prob[] writes visible below
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float l = red[0] + red[1] + red[2] + red[3];
    const float inv = l > 0.f ? 1.0f / l : 0.f;        // all keys masked: zero row, as the full kernel produces
    // phase 3: P.V
    float acc0 = 0.f, acc1 = 0.f;
    constexpr int U = 8;
    int k = warp;
    for (; k + 4 * (U - 1) < S; k += 4 * U) {
        __nv_bfloat162 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            v[u] = __ldg(reinterpret_cast<const __nv_bfloat162*>(vbase + (long long)(k + 4 * u) * p.ld_kv) + lane);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float pk = prob[k + 4 * u];
            const float2 f = __bfloat1622float2(v[u]);
            acc0 = fmaf(pk, f.x, acc0);
            acc1 = fmaf(pk, f.y, acc1);
        }
    }
    for (; k < S; k += 4) {
        const float pk = prob[k];
        const float2 f = __bfloat1622float2(__ldg(reinterpret_cast<const __nv_bfloat162*>(vbase + (long long)k * p.ld_kv) + lane));
        acc0 = fmaf(pk, f.x, acc0);
        acc1 = fmaf(pk, f.y, acc1);
    }
    part[warp][2 * lane] = acc0;
    part[warp][2 * lane + 1] = acc1;
    __syncthreads();
    if (t < kAcD / 2) {
        const float o0 = (part[0][2 * t] + part[1][2 * t] + part[2][2 * t] + part[3][2 * t]) * inv;
        const float o1 = (part[0][2 * t + 1] + part[1][2 * t + 1] + part[2][2 * t + 1] + part[3][2 * t + 1]) * inv;
        reinterpret_cast<__nv_bfloat162*>(p.ctx + (long long)b * p.ld_ctx + h * kAcD)[t] = __floats2bfloat162_rn(o0, o1);
    }
}

}  // namespace fame

// K1s: "skinny" linear layer for at most 32 rows:  Y[M<=32, N] = epi( X[M, K] . W[N, K]^T )
//
// The demographic tower of FAME (BEHRTModel_Demo, 10_FAME.py:175-206) runs a 12-layer BERT over ONE token per
// patient, so at the reference batch size (32 patients per GPU) every one of its 48 forward and 48 data-gradient
// products has M = 32 rows: 0.15 GFLOP against 1.2-4.7 MB of weights, i.e. pure weight streaming (HBM / L2 bound by
// two orders of magnitude).  A 128x256 tcgen05 tile would use 3-12 CTAs of the 148 and leave 75 % of each MMA empty;
// this kernel instead spreads the weight matrix over N/8 CTAs x 8 warps (each warp one K slice of one 8-column
// block), feeds mma.sync.m16n8k16 straight from 16-byte global loads (no shared-memory staging: the contraction index
// may be permuted freely as long as A and B use the same permutation, so a thread's 8 consecutive k values serve two
// MMAs), reduces the 8 K-slices through shared memory and applies the same epilogue family as the tcgen05 GEMM.
// Tensor cores are used through the legacy warp-level path on purpose: the op is bandwidth bound, M = 32 cannot fill
// a tcgen05 tile, and the goal is one ~3 us launch instead of ~25 us.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_sm100.cuh"   // GemmParams, gelu_erf, kAct*, kRes*
#include "rowwise.cuh"      // bf16x8_to_float

namespace fame {

constexpr int kSkWarps = 8;
constexpr int kSkThreads = kSkWarps * 32;
constexpr int kSkMaxM = 32;
constexpr int kSkBN = 8;

struct SkinnyParams {
    const __nv_bfloat16* x;   // [M, ldx]
    const __nv_bfloat16* w;   // [N, ldw]
    long long ldx, ldw;
    int M, N, K;              // M <= 32, N % 8 == 0, K % 32 == 0
    const float* bias;
    const void* residual;
    long long ldr;
    int res_mode;             // kRes*
    void* y;
    long long ldy;
    int y_f32;
    int act;
    float alpha;
    DropCfg drop;             // dropout after the activation, before the residual add
    __nv_bfloat16* pre_act;   // optional [M, ld_pre]: x W^T + bias BEFORE the activation (saved for the GELU backward)
    long long ld_pre;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kSkThreads)
skinny_gemm_kernel(const SkinnyParams p) {
    __shared__ float red[kSkWarps][kSkMaxM][kSkBN + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * kSkBN;
    // All of this CTA's weight bytes (8 rows x K) are requested from HBM at once, before the first dependent load: the
    // 4-chunk software pipeline below then runs against L2 (A/B on one box, three runs each: 3.53 -> 3.45-3.50 ms per training step).
    {
        const char* wbase = reinterpret_cast<const char*>(p.w + (long long)n0 * p.ldw);
        const int lines = (p.K * 2 + 127) >> 7;
        for (int i = threadIdx.x; i < kSkBN * lines; i += kSkThreads) {
            const int r = i / lines, l = i - r * lines;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(wbase + (long long)r * p.ldw * 2 + l * 128));
        }
    }
    // K slice of this warp, in 32-element chunks
    const int chunks = p.K >> 5;
    const int per = (chunks + kSkWarps - 1) / kSkWarps;
    const int c_lo = warp * per, c_hi = min(chunks, c_lo + per);

    float acc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;

    const uint4* wrow = reinterpret_cast<const uint4*>(p.w + (long long)(n0 + g) * p.ldw) + t;   // + 4 per chunk
    const uint4* xrow[4];
    bool xok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = g + 8 * r;
        xok[r] = row < p.M;
        xrow[r] = reinterpret_cast<const uint4*>(p.x + (long long)(xok[r] ? row : 0) * p.ldx) + t;
    }
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);

    constexpr int U = 4;   // chunks in flight per warp: 4 x (1 weight + 4 activation) 16-byte loads per thread
    for (int c = c_lo; c < c_hi; c += U) {
        uint4 wv[U], xv[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool live = c + u < c_hi;
            wv[u] = live ? __ldg(wrow + (c + u) * 4) : zero;
#pragma unroll
            for (int r = 0; r < 4; ++r) xv[u][r] = (live && xok[r]) ? __ldg(xrow[r] + (c + u) * 4) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                // rows 16 mt + g (fragment rows "g") and 16 mt + g + 8 (fragment rows "g + 8")
                const uint4& lo = xv[u][2 * mt];
                const uint4& hi = xv[u][2 * mt + 1];
                mma_bf16_16816(acc[mt], lo.x, hi.x, lo.y, hi.y, wv[u].x, wv[u].y);
                mma_bf16_16816(acc[mt], lo.z, hi.z, lo.w, hi.w, wv[u].z, wv[u].w);
            }
        }
    }
    // C fragment: c0,c1 -> (row g, cols 2t, 2t+1); c2,c3 -> (row g + 8, same cols)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        red[warp][16 * mt + g][2 * t] = acc[mt][0];
        red[warp][16 * mt + g][2 * t + 1] = acc[mt][1];
        red[warp][16 * mt + g + 8][2 * t] = acc[mt][2];
        red[warp][16 * mt + g + 8][2 * t + 1] = acc[mt][3];
    }
    __syncthreads();
    const int row = threadIdx.x >> 3, col = threadIdx.x & 7;   // 32 x 8 outputs, one per thread
    if (row >= p.M) return;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kSkWarps; ++w) v += red[w][row][col];
    const int n = n0 + col;
    v *= p.alpha;
    if (p.bias != nullptr) v += __ldg(p.bias + n);
    if (p.pre_act != nullptr) {
        const __nv_bfloat16 pb = __float2bfloat16_rn(v);
        p.pre_act[(long long)row * p.ld_pre + n] = pb;
        v = __bfloat162float(pb);      // the activation sees the rounded value, as when it is applied by a second kernel
    }
    if (p.act == kActGelu) v = gelu_erf(v);
    else if (p.act == kActRelu) v = fmaxf(v, 0.f);
    if (p.drop.thresh16 != 0u) {
        const uint32_t rs = drop_row_seed(drop_site_seed(p.drop), (uint32_t)row);
        v = drop_keep(rs, (uint32_t)n >> p.drop.group_shift, p.drop.thresh16) ? v * drop_inv_keep(p.drop.thresh16) : 0.f;
    }
    if (p.res_mode == kResAddF32) {
        v += __ldg(reinterpret_cast<const float*>(p.residual) + (long long)row * p.ldr + n);
    } else if (p.res_mode == kResAddBf16) {
        v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[(long long)row * p.ldr + n]);
    } else if (p.res_mode == kResReluMaskBf16) {
        const float a = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[(long long)row * p.ldr + n]);
        v = a > 0.f ? v : 0.f;
    } else if (p.res_mode == kResGeluBwdBf16) {
        // erf-GELU backward: d pre = d h * (Phi(x) + x phi(x)), x = the saved pre-activation; d h rounded to bf16 first,
        // as when the product and the GELU backward are two kernels
        const float x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[(long long)row * p.ldr + n]);
        const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
        const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
        v = __bfloat162float(__float2bfloat16_rn(v)) * (cdf + x * pdf);
    }
    if (p.y_f32) reinterpret_cast<float*>(p.y)[(long long)row * p.ldy + n] = v;
    else reinterpret_cast<__nv_bfloat16*>(p.y)[(long long)row * p.ldy + n] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of a <= 32-row layer:  dW[n, k] (+)= sum_{m < M} dY[m, n] * X[m, k]      (M <= 32)
// 32 multiply-adds per 4-byte output: bound by the f32 gradient write (9.4 MB for a 3072 x 768 weight), not by math,
// so plain FP32 FMAs from shared-memory tiles; CTA = 64 (n) x 128 (k) outputs, thread = 4 x 8 outputs.
constexpr int kWsBN = 64, kWsBK = 128;

__global__ void __launch_bounds__(256)
wgrad_small_kernel(const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ x,
                   long long ld_x, float* __restrict__ out, long long ld_out, int M, int N, int K, int accumulate,
                   float* __restrict__ dbias) {
    __shared__ __align__(16) float s_dy[kSkMaxM][kWsBN];
    __shared__ __align__(16) float s_x[kSkMaxM][kWsBK];
    const int n0 = blockIdx.y * kWsBN, k0 = blockIdx.x * kWsBK;
    // stage the two operand tiles with 16-byte loads, all three issued before the first use (a per-element loop with
    // bounds-checked 2-byte loads serialised 24 memory round trips: 17 us per launch)
    {
        const bool vec = ((ld_dy | ld_x) & 7) == 0 && (N & 7) == 0 && (K & 7) == 0 &&
                         ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
        if (vec) {
            const int m_d = threadIdx.x >> 3, c_d = (threadIdx.x & 7) * 8;          // 32 rows x 8 chunks
            uint4 vd = make_uint4(0u, 0u, 0u, 0u), vx[2];
            if (m_d < M && n0 + c_d < N) vd = __ldg(reinterpret_cast<const uint4*>(dy + (long long)m_d * ld_dy + n0 + c_d));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ch = threadIdx.x + 256 * h, m_x = ch >> 4, c_x = (ch & 15) * 8;   // 32 rows x 16 chunks
                vx[h] = make_uint4(0u, 0u, 0u, 0u);
                if (m_x < M && k0 + c_x < K) vx[h] = __ldg(reinterpret_cast<const uint4*>(x + (long long)m_x * ld_x + k0 + c_x));
            }
            float f[8];
            bf16x8_to_float(vd, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s_dy[m_d][c_d + j] = f[j];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ch = threadIdx.x + 256 * h, m_x = ch >> 4, c_x = (ch & 15) * 8;
                bf16x8_to_float(vx[h], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s_x[m_x][c_x + j] = f[j];
            }
        } else {
            for (int i = threadIdx.x; i < kSkMaxM * kWsBN; i += 256) {
                const int m = i / kWsBN, n = i % kWsBN;
                s_dy[m][n] = (m < M && n0 + n < N) ? __bfloat162float(dy[(long long)m * ld_dy + n0 + n]) : 0.f;
            }
            for (int i = threadIdx.x; i < kSkMaxM * kWsBK; i += 256) {
                const int m = i / kWsBK, k = i % kWsBK;
                s_x[m][k] = (m < M && k0 + k < K) ? __bfloat162float(x[(long long)m * ld_x + k0 + k]) : 0.f;
            }
        }
    }
    __syncthreads();
    // bias gradient = column sums of dY, from the tile the first k-block of each n-block has staged anyway (saves the
    // separate column-sum launch: 48 of the ~270 launches of the demographic tower's step)
    if (dbias != nullptr && blockIdx.x == 0 && threadIdx.x < kWsBN && n0 + (int)threadIdx.x < N) {
        float sacc = 0.f;
#pragma unroll 8
        for (int m = 0; m < kSkMaxM; ++m) sacc += s_dy[m][threadIdx.x];
        if (accumulate) dbias[n0 + threadIdx.x] += sacc;
        else dbias[n0 + threadIdx.x] = sacc;
    }
    const int tk = threadIdx.x & 15, tn = threadIdx.x >> 4;   // 16 threads along k (8 each), 16 along n (4 each)
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int m = 0; m < kSkMaxM; ++m) {
        const float4 d = *reinterpret_cast<const float4*>(&s_dy[m][tn * 4]);
        const float4 xa = *reinterpret_cast<const float4*>(&s_x[m][tk * 8]);
        const float4 xb = *reinterpret_cast<const float4*>(&s_x[m][tk * 8 + 4]);
        const float dv[4] = {d.x, d.y, d.z, d.w};
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + tn * 4 + i;
        if (n >= N) continue;
        float* o = out + (long long)n * ld_out + k0 + tk * 8;
        if (k0 + tk * 8 + 8 <= K && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            float4 a = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            float4 b = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
            if (accumulate) {
                const float4 pa = *reinterpret_cast<float4*>(o), pb = *reinterpret_cast<float4*>(o + 4);
                a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
                b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
            }
            *reinterpret_cast<float4*>(o) = a;
            *reinterpret_cast<float4*>(o + 4) = b;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k0 + tk * 8 + j < K) o[j] = accumulate ? o[j] + acc[i][j] : acc[i][j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Transposed bf16 shadows of weight matrices (dgrad of the skinny path: dX = dY . W needs W^T rows contiguous in the
// contraction index).  One launch for a table of matrices: entry = {src, dst, rows, cols, first tile}.
struct TransposeEntry {
    const __nv_bfloat16* src;   // [rows, cols]
    __nv_bfloat16* dst;         // [cols, rows]
    int rows, cols;
    int tile0;                  // first 64x64 tile index of this matrix in the launch
    int tiles_x;                // tiles along cols
};

__global__ void __launch_bounds__(256)
transpose_bf16_table_kernel(const TransposeEntry* __restrict__ table, int n_entries) {
    __shared__ __nv_bfloat16 tile[64][66];   // row stride 33 words: column reads are bank-conflict free
    __shared__ int s_e;
    if (threadIdx.x == 0) {
        int e = 0;
        while (e + 1 < n_entries && table[e + 1].tile0 <= (int)blockIdx.x) ++e;
        s_e = e;
    }
    __syncthreads();
    const TransposeEntry en = table[s_e];
    const int tl = blockIdx.x - en.tile0;
    const int r0 = (tl / en.tiles_x) * 64, c0 = (tl % en.tiles_x) * 64;
    const bool vec = (en.rows & 7) == 0 && (en.cols & 7) == 0 &&
                     ((reinterpret_cast<uintptr_t>(en.src) | reinterpret_cast<uintptr_t>(en.dst)) & 15) == 0;
    if (vec) {
        // 64 x 64 tile = 512 chunks of 8 bf16; thread handles chunks threadIdx.x and threadIdx.x + 256
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ch = threadIdx.x + 256 * h, r = ch >> 3, cc = (ch & 7) * 8;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r0 + r < en.rows && c0 + cc < en.cols)
                v = __ldg(reinterpret_cast<const uint4*>(en.src + (long long)(r0 + r) * en.cols + c0 + cc));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint32_t*>(&tile[r][cc + 2 * j]) = w[j];
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ch = threadIdx.x + 256 * h, c = ch >> 3, rr = (ch & 7) * 8;   // output row c, 8 source rows
            if (c0 + c < en.cols && r0 + rr < en.rows) {
                __align__(16) __nv_bfloat16 o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = tile[rr + j][c];
                *reinterpret_cast<uint4*>(en.dst + (long long)(c0 + c) * en.rows + r0 + rr) = *reinterpret_cast<uint4*>(o);
            }
        }
        return;
    }
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = r0 + ty + 4 * i, c = c0 + tx;
        if (r < en.rows && c < en.cols) tile[ty + 4 * i][tx] = en.src[(long long)r * en.cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int c = c0 + ty + 4 * i, r = r0 + tx;
        if (r < en.rows && c < en.cols) en.dst[(long long)c * en.rows + r] = tile[tx][ty + 4 * i];
    }
}

}  // namespace fame

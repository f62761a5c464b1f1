// K7 fusion head forward, K8 joint BCE + LEDDI loss (statistics pass + loss / gradient pass).  fp32 CUDA-core
// kernels: these stages are a few MFLOP per patient and bound by memory / latency, not by the tensor pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rowwise.cuh"

namespace fame {

// ================================================================================================ K7 fusion fwd
// MultimodalTransformer_EDDI_Sigmoid.forward after the encoders (10_FAME.py:276-308):
//   proj_m  = relu(W_m e_m + b_m)                      m in {demo, lab, text},  768 -> 256
//   gated   = [w_d proj_d | w_l proj_l | w_t proj_t] * sigmoid(sig_weights)
//   pre     = W3 gated + b3 (768 -> 512);  logits = W4 relu(pre) + b4 (512 -> 3)      [dropout: identity here]
//   modality_logits_m = Wc_m proj_m + bc_m (256 -> 3)
// One CTA = 8 patients; inputs staged in shared memory, weights (pre-transposed, L2 resident) streamed coalesced.
constexpr int kFuRows = 8;
constexpr int kFuThreads = 256;
constexpr int kFuSmemBytes = (3 * kFuRows * 768 + kFuRows * 768 + kFuRows * 512) * 4;

struct FusionParams {
    const float* emb[3];   // demo, lab, text  [B,768]
    const float* wp_t;     // [3][768][256]  projector weights, transposed
    const float* bp;       // [3][256]
    float w_mod[3];        // EDDI modality weights (the "mortality" entry, 10_FAME.py:283-285)
    const float* sig_w;    // [768]
    const float* w3_t;     // [768][512]  fusion_mlp.0 weight, transposed
    const float* b3;       // [512]
    const float* w4;       // [3][512]    fusion_mlp.3 weight
    const float* b4;       // [3]
    const float* wc;       // [3][3][256] classifier_{demo,lab,text}.weight
    const float* bc;       // [3][3]
    float* proj;           // [B,768] relu outputs (unweighted), nullable
    float* gated;          // [B,768] nullable
    float* pre_relu;       // [B,512] nullable
    float* logits;         // [B,3]
    float* mod_logits;     // [3][B][3] nullable
    float* sig_out;        // [768] sigmoid(sig_w), nullable
    int B;
    const float* w_mod_dev;  // optional device copy of w_mod, read at run time (CUDA-graph replays follow updates)
    __device__ __forceinline__ float wmod(int m) const { return w_mod_dev != nullptr ? __ldg(w_mod_dev + m) : w_mod[m]; }
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(kFuThreads)
fusion_fwd_kernel(const FusionParams p) {
    extern __shared__ float fsm[];
    float* xin = fsm;                          // [3][8][768]; after phase 1 xin[m][r][0..255] holds proj_m
    float* g = xin + 3 * kFuRows * 768;        // [8][768] gated
    float* hid = g + kFuRows * 768;            // [8][512] relu(pre)
    const int t = threadIdx.x;
    const int row0 = blockIdx.x * kFuRows;
    const int nrow = min(kFuRows, p.B - row0);

    for (int m = 0; m < 3; ++m)
        for (int i = t; i < kFuRows * 192; i += kFuThreads) {
            const int r = i / 192, c4 = i % 192;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrow) v = __ldg(reinterpret_cast<const float4*>(p.emb[m] + (long long)(row0 + r) * 768) + c4);
            reinterpret_cast<float4*>(xin + (m * kFuRows + r) * 768)[c4] = v;
        }
    if (blockIdx.x == 0 && p.sig_out != nullptr)
        for (int i = t; i < 768; i += kFuThreads) p.sig_out[i] = 1.0f / (1.0f + expf(-p.sig_w[i]));
    __syncthreads();

    // ---- phase 1: projectors; thread t owns output column t of each modality for all 8 rows
    float pr[3][kFuRows];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        float acc[kFuRows];
#pragma unroll
        for (int r = 0; r < kFuRows; ++r) acc[r] = 0.f;
        const float* w = p.wp_t + (long long)m * 768 * 256 + t;
        const float* xm = xin + m * kFuRows * 768;
        // software-pipelined weight stream: the next 16 weights are in flight while the current 16 are consumed
        // (a 4-deep dependent loop left the kernel latency bound: 345 us at 32 patients)
        constexpr int KU = 16;
        float wn[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) wn[u] = __ldg(w + u * 256);
        for (int k = 0; k < 768; k += KU) {
            float wc[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) wc[u] = wn[u];
            if (k + KU < 768) {
#pragma unroll
                for (int u = 0; u < KU; ++u) wn[u] = __ldg(w + (k + KU + u) * 256);
            }
#pragma unroll
            for (int r = 0; r < kFuRows; ++r) {
#pragma unroll
                for (int u = 0; u < KU; u += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(xm + r * 768 + k + u);
                    acc[r] = fmaf(x.x, wc[u], acc[r]);
                    acc[r] = fmaf(x.y, wc[u + 1], acc[r]);
                    acc[r] = fmaf(x.z, wc[u + 2], acc[r]);
                    acc[r] = fmaf(x.w, wc[u + 3], acc[r]);
                }
            }
        }
        const float b = __ldg(p.bp + m * 256 + t);
#pragma unroll
        for (int r = 0; r < kFuRows; ++r) pr[m][r] = fmaxf(acc[r] + b, 0.f);
    }
    __syncthreads();  // everyone is done reading xin
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const float sg = 1.0f / (1.0f + expf(-__ldg(p.sig_w + m * 256 + t)));
#pragma unroll
        for (int r = 0; r < kFuRows; ++r) {
            xin[(m * kFuRows + r) * 768 + t] = pr[m][r];
            const float gv = (p.wmod(m) * pr[m][r]) * sg;
            g[r * 768 + m * 256 + t] = gv;
            if (r < nrow) {
                if (p.proj != nullptr) p.proj[(long long)(row0 + r) * 768 + m * 256 + t] = pr[m][r];
                if (p.gated != nullptr) p.gated[(long long)(row0 + r) * 768 + m * 256 + t] = gv;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: hidden layer 768 -> 512; thread t owns columns t and t + 256
    {
        float a0[kFuRows], a1[kFuRows];
#pragma unroll
        for (int r = 0; r < kFuRows; ++r) a0[r] = a1[r] = 0.f;
        const float* w = p.w3_t + t;
        constexpr int KU = 8;
        float wan[KU], wbn[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            wan[u] = __ldg(w + u * 512);
            wbn[u] = __ldg(w + u * 512 + 256);
        }
        for (int k = 0; k < 768; k += KU) {
            float wa[KU], wb[KU];
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                wa[u] = wan[u];
                wb[u] = wbn[u];
            }
            if (k + KU < 768) {
#pragma unroll
                for (int u = 0; u < KU; ++u) {
                    wan[u] = __ldg(w + (k + KU + u) * 512);
                    wbn[u] = __ldg(w + (k + KU + u) * 512 + 256);
                }
            }
#pragma unroll
            for (int r = 0; r < kFuRows; ++r) {
#pragma unroll
                for (int u = 0; u < KU; u += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(g + r * 768 + k + u);
                    a0[r] = fmaf(x.x, wa[u], a0[r]); a0[r] = fmaf(x.y, wa[u + 1], a0[r]);
                    a0[r] = fmaf(x.z, wa[u + 2], a0[r]); a0[r] = fmaf(x.w, wa[u + 3], a0[r]);
                    a1[r] = fmaf(x.x, wb[u], a1[r]); a1[r] = fmaf(x.y, wb[u + 1], a1[r]);
                    a1[r] = fmaf(x.z, wb[u + 2], a1[r]); a1[r] = fmaf(x.w, wb[u + 3], a1[r]);
                }
            }
        }
        const float b0 = __ldg(p.b3 + t), b1 = __ldg(p.b3 + t + 256);
#pragma unroll
        for (int r = 0; r < kFuRows; ++r) {
            const float v0 = a0[r] + b0, v1 = a1[r] + b1;
            hid[r * 512 + t] = fmaxf(v0, 0.f);
            hid[r * 512 + t + 256] = fmaxf(v1, 0.f);
            if (r < nrow && p.pre_relu != nullptr) {
                p.pre_relu[(long long)(row0 + r) * 512 + t] = v0;
                p.pre_relu[(long long)(row0 + r) * 512 + t + 256] = v1;
            }
        }
    }
    __syncthreads();

    // ---- phase 3: 3 fused logits and 9 modality logits per row; warp w owns row w
    const int warp = t >> 5, lane = t & 31;
    if (warp < nrow) {
        const int r = warp;
        float l[3] = {0.f, 0.f, 0.f};
        for (int k = lane; k < 512; k += 32) {
            const float h = hid[r * 512 + k];
#pragma unroll
            for (int o = 0; o < 3; ++o) l[o] = fmaf(h, __ldg(p.w4 + o * 512 + k), l[o]);
        }
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            l[o] = warp_sum(l[o]);
            if (lane == 0) p.logits[(long long)(row0 + r) * 3 + o] = l[o] + __ldg(p.b4 + o);
        }
        if (p.mod_logits != nullptr) {
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                float c[3] = {0.f, 0.f, 0.f};
                for (int k = lane; k < 256; k += 32) {
                    const float x = xin[(m * kFuRows + r) * 768 + k];
#pragma unroll
                    for (int o = 0; o < 3; ++o) c[o] = fmaf(x, __ldg(p.wc + (m * 3 + o) * 256 + k), c[o]);
                }
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    c[o] = warp_sum(c[o]);
                    if (lane == 0)
                        p.mod_logits[((long long)m * p.B + row0 + r) * 3 + o] = c[o] + __ldg(p.bc + m * 3 + o);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K7, few patients
// The fused kernel above gives one CTA per 8 patients: at the training batch size (32 per GPU) that is 4 CTAs, each
// streaming all 2.75 MB of fp32 head weights through 16 KB of loads in flight -- 230-345 us of pure latency.  For small
// batches the head is therefore split into three launches whose CTAs each own 32 OUTPUT COLUMNS (x 8 patients) and
// whose 8 warps split the 768-long contraction, so the weight stream is spread over 24 / 16 column groups x B/8
// row groups: same arithmetic in fp32, same outputs, ~10 us in total.
constexpr int kFsCols = 32;      // output columns per CTA (lane = column: weight rows are read 128 bytes at a time)
constexpr int kFsKU = 12;        // weight loads in flight per thread (96 per warp slice / 8 trips)

// y[r, col] = sum_k x[r, k] * wt[k, col]   for the CTA's 8 rows and 32 columns; wt is [768][ldw] (transposed weight)
__device__ __forceinline__ void fusion_small_tile(const float* __restrict__ xs /* smem [8][768] */,
                                                  const float* __restrict__ wt, int ldw, int col,
                                                  float (*red)[kFuRows][kFsCols], float (&out)[1], int warp, int lane) {
    float acc[kFuRows];
#pragma unroll
    for (int r = 0; r < kFuRows; ++r) acc[r] = 0.f;
    const int kb = warp * 96;
    const float* w = wt + (long long)kb * ldw + col;
    for (int k0 = 0; k0 < 96; k0 += kFsKU) {
        float wv[kFsKU];
#pragma unroll
        for (int u = 0; u < kFsKU; ++u) wv[u] = __ldg(w + (long long)(k0 + u) * ldw);
#pragma unroll
        for (int u = 0; u < kFsKU; ++u)
#pragma unroll
            for (int r = 0; r < kFuRows; ++r) acc[r] = fmaf(xs[r * 768 + kb + k0 + u], wv[u], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kFuRows; ++r) red[warp][r][lane] = acc[r];
    __syncthreads();
    // thread (r = warp, c = lane) adds the 8 K-slices in a fixed order
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][warp][lane];
    out[0] = v;
}

__global__ void __launch_bounds__(256)
fusion_small_proj_kernel(const FusionParams p, float* __restrict__ proj, float* __restrict__ gated) {
    __shared__ __align__(16) float xs[kFuRows * 768];
    __shared__ float red[8][kFuRows][kFsCols];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m = blockIdx.x / 8, col_in_m = (blockIdx.x % 8) * kFsCols + lane;
    const int row0 = blockIdx.y * kFuRows;
    const int nrow = min(kFuRows, p.B - row0);
    for (int i = t; i < kFuRows * 192; i += 256) {
        const int r = i / 192, c4 = i % 192;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrow) v = __ldg(reinterpret_cast<const float4*>(p.emb[m] + (long long)(row0 + r) * 768) + c4);
        reinterpret_cast<float4*>(xs + r * 768)[c4] = v;
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && p.sig_out != nullptr)
        for (int i = t; i < 768; i += 256) p.sig_out[i] = 1.0f / (1.0f + expf(-p.sig_w[i]));
    __syncthreads();
    float v[1];
    fusion_small_tile(xs, p.wp_t + (long long)m * 768 * 256, 256, col_in_m, red, v, warp, lane);
    const int r = warp;
    if (r < nrow) {
        const int c = m * 256 + col_in_m;
        const float pr = fmaxf(v[0] + __ldg(p.bp + c), 0.f);
        const float sg = 1.0f / (1.0f + expf(-__ldg(p.sig_w + c)));
        proj[(long long)(row0 + r) * 768 + c] = pr;
        gated[(long long)(row0 + r) * 768 + c] = (p.wmod(m) * pr) * sg;
    }
}

__global__ void __launch_bounds__(256)
fusion_small_hidden_kernel(const FusionParams p, const float* __restrict__ gated, float* __restrict__ pre_relu) {
    __shared__ __align__(16) float xs[kFuRows * 768];
    __shared__ float red[8][kFuRows][kFsCols];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int col = blockIdx.x * kFsCols + lane;
    const int row0 = blockIdx.y * kFuRows;
    const int nrow = min(kFuRows, p.B - row0);
    for (int i = t; i < kFuRows * 192; i += 256) {
        const int r = i / 192, c4 = i % 192;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrow) v = *(reinterpret_cast<const float4*>(gated + (long long)(row0 + r) * 768) + c4);
        reinterpret_cast<float4*>(xs + r * 768)[c4] = v;
    }
    __syncthreads();
    float v[1];
    fusion_small_tile(xs, p.w3_t, 512, col, red, v, warp, lane);
    if (warp < nrow) pre_relu[(long long)(row0 + warp) * 512 + col] = v[0] + __ldg(p.b3 + col);
}

// logits (512 -> 3) and the 9 modality logits (256 -> 3 each): one warp per patient
__global__ void __launch_bounds__(256)
fusion_small_logits_kernel(const FusionParams p, const float* __restrict__ proj, const float* __restrict__ pre_relu) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= p.B) return;
    float l[3] = {0.f, 0.f, 0.f};
    for (int k = lane; k < 512; k += 32) {
        const float h = fmaxf(pre_relu[(long long)row * 512 + k], 0.f);
#pragma unroll
        for (int o = 0; o < 3; ++o) l[o] = fmaf(h, __ldg(p.w4 + o * 512 + k), l[o]);
    }
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        l[o] = warp_sum(l[o]);
        if (lane == 0) p.logits[(long long)row * 3 + o] = l[o] + __ldg(p.b4 + o);
    }
    if (p.mod_logits != nullptr) {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            float c[3] = {0.f, 0.f, 0.f};
            for (int k = lane; k < 256; k += 32) {
                const float x = proj[(long long)row * 768 + m * 256 + k];
#pragma unroll
                for (int o = 0; o < 3; ++o) c[o] = fmaf(x, __ldg(p.wc + (m * 3 + o) * 256 + k), c[o]);
            }
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                c[o] = warp_sum(c[o]);
                if (lane == 0) p.mod_logits[((long long)m * p.B + row) * 3 + o] = c[o] + __ldg(p.bc + m * 3 + o);
            }
        }
    }
}

// ================================================================================================ K8 loss
// Statistics layout (int64, identical on every rank so that one SUM all-reduce makes them global):
//   [0..2]   sum_b |sigmoid(z)-y|  per outcome            fixed point 2^24
//   [3..5]   sum_b bce term        per outcome            fixed point 2^24
//   [6..77]  group error sums [outcome][attr][code 0..7]  fixed point 2^24
//   [78..101] group counts [attr][code]                   integer
//   [102]    patients                                     integer
//   [103]    error flag (an attribute code outside 0..7)
// Every PATIENT's term is rounded to fixed point first and only integers are added afterwards, so each statistic is
// an exact integer sum: independent of thread / block / rank order.  N data-parallel ranks therefore reproduce the
// statistics of the single-process run on the concatenated batch bit for bit, and so does a re-run.
// |sigmoid(z) - y| lies in [0, 1]: 2^-24 is the float32 resolution at 1.0, the rounding error per patient <= 3e-8.
constexpr int kLossStatsLen = 104;
constexpr int kLossSlots = 8;
constexpr double kFixErr = 16777216.0;     // 2^24
constexpr double kFixBce = 16777216.0;     // 2^24
constexpr int kLossTripsPerFlush = 63;     // 63 trips x 4 patients x 2^24 < 2^32: uint32 accumulators cannot overflow

struct LossStatsParams {
    const float* logits;     // [B,3]
    const float* labels;     // [B,3]
    const long long* attr[3];  // age, ethnicity, insurance  [B]
    const float* pos_weight; // [3]
    long long* stats;        // [kLossStatsLen], zero-initialised by the caller's memset (done by the host wrapper)
    int B;
};

// exact warp sum of 32 uint32 values as a 64-bit integer: two hardware REDUX adds over the 16-bit halves
__device__ __forceinline__ unsigned long long warp_sum_u32_exact(unsigned v) {
    const unsigned lo = __reduce_add_sync(0xffffffffu, v & 0xffffu);
    const unsigned hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return ((unsigned long long)hi << 16) + lo;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (A variant that groups lanes by subgroup code with match.any / REDUX and keeps the sums in shared memory was
// measured SLOWER at scale -- 14.7 % vs 26.5 % of the copy bandwidth at 32 M patients: the warp collectives cost more
// than the predicated adds they replace.  The variant with all 96 subgroup accumulators in registers -- predicated
// adds over the 8 code slots -- needed 225 registers: one CTA per SM, 14.7 % of the copy bandwidth after the
// deterministic fixed-point change.)
// Subgroup accumulators are THREAD-PRIVATE COLUMNS OF SHARED MEMORY: bins[bin][thread], bin = outcome x attribute x
// code (72 fixed-point sums) + attribute x code (24 counts).  The code indexes the bin directly (what registers cannot
// do), the column index is the thread, so there are no atomics and no bank conflicts (bank = thread % 32), a patient
// costs 12 read-modify-writes, and the kernel needs ~64 registers: 4 CTAs of 128 threads per SM with 16 patients in
// flight per thread.  Every kLossTripsPerFlush trips (uint32 bins cannot overflow before that) and at the end, 96
// threads add up one bin each across the 128 columns (rotated start => conflict-free) into 64-bit block totals.
constexpr int kLsThreads = 128;
constexpr int kLsBins = 96;
constexpr int kLsSmemBytes = kLsBins * kLsThreads * 4 + kLossStatsLen * 8;

__global__ void __launch_bounds__(kLsThreads, 4)
loss_stats_kernel(const LossStatsParams p) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    unsigned (*bins)[kLsThreads] = reinterpret_cast<unsigned (*)[kLsThreads]>(ls_smem);
    unsigned long long* sh = reinterpret_cast<unsigned long long*>(ls_smem + kLsBins * kLsThreads * 4);
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned e_sum[3] = {0u, 0u, 0u};
    unsigned long long b_sum[3] = {0ull, 0ull, 0ull};
    unsigned n_local = 0, bad = 0;
    float pw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) pw[i] = __ldg(p.pos_weight + i);
#pragma unroll 8
    for (int b = 0; b < kLsBins; ++b) bins[b][tid] = 0u;
    for (int i = tid; i < kLossStatsLen; i += kLsThreads) sh[i] = 0ull;
    __syncthreads();

    auto put = [&](int idx, unsigned long long v) {
        if (lane == 0 && v != 0ull) atomicAdd(&sh[idx], v);
    };
    auto flush = [&]() {     // block-collective: the trip count is uniform over the block
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            put(i, warp_sum_u32_exact(e_sum[i]));
            put(3 + i, warp_sum_u64(b_sum[i]));
            e_sum[i] = 0u;
            b_sum[i] = 0ull;
        }
        put(102, __reduce_add_sync(0xffffffffu, n_local));
        n_local = 0;
        __syncthreads();
        if (tid < kLsBins) {
            unsigned long long tot = 0ull;
#pragma unroll 8
            for (int j = 0; j < kLsThreads; ++j) tot += bins[tid][(j + tid) & (kLsThreads - 1)];
            // bins 0..71 = [outcome][attr][code] sums -> stats[6 + ...]; bins 72..95 = [attr][code] counts -> stats[78 + ...]
            if (tot != 0ull) atomicAdd(&sh[6 + tid], tot);
        }
        __syncthreads();
#pragma unroll 8
        for (int b = 0; b < kLsBins; ++b) bins[b][tid] = 0u;
    };

    int trips = 0;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(p.labels) |
                          reinterpret_cast<uintptr_t>(p.attr[0]) | reinterpret_cast<uintptr_t>(p.attr[1]) |
                          reinterpret_cast<uintptr_t>(p.attr[2])) & 15) == 0;
    const long long per_block = (long long)kLsThreads * kPatPerThread;
    // the loop bound is uniform per block (base, not base + thread offset), so the block never diverges around flush()
    for (long long base = blockIdx.x * per_block; base < p.B; base += gridDim.x * per_block) {
        PatientQuad q;
        load_patient_quad(q, p.logits, 3, p.labels, p.attr, base + (long long)tid * kPatPerThread, p.B, vec_ok);
#pragma unroll
        for (int u = 0; u < kPatPerThread; ++u) {
            if (u < q.n) {
                unsigned e[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float z = q.z[u][i], y = q.y[u][i];
                    const float pr = 1.0f / (1.0f + expf(-z));
                    e[i] = __float2uint_rn(fminf(fabsf(pr - y), 1.0f) * 16777216.0f);
                    // softplus(-z) = -log sigmoid(z), stable:  max(-z, 0) + log1p(exp(-|z|))
                    const float sp = fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z)));
                    const float term = pw[i] * y * sp + (1.0f - y) * (sp + z);
                    b_sum[i] += __float2ull_rn(fmaxf(term, 0.f) * 16777216.0f);
                    e_sum[i] += e[i];
                }
                ++n_local;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const long long c = q.code[a][u];
                    if (c < 0 || c >= kLossSlots) {
                        bad = 1;
                    } else {
                        const int s = a * kLossSlots + (int)c;
                        bins[72 + s][tid] += 1u;
#pragma unroll
                        for (int i = 0; i < 3; ++i) bins[i * 24 + s][tid] += e[i];
                    }
                }
            }
        }
        if (++trips == kLossTripsPerFlush) {
            flush();
            trips = 0;
        }
    }
    flush();
    put(103, __reduce_add_sync(0xffffffffu, bad));
    __syncthreads();
    for (int i = tid; i < kLossStatsLen; i += kLsThreads) {
        const unsigned long long v = sh[i];
        if (v != 0ull) atomicAdd(reinterpret_cast<unsigned long long*>(p.stats + i), v);
    }
}

struct LossGradParams {
    const float* logits;
    const float* labels;
    const long long* attr[3];
    const float* pos_weight;
    const long long* stats;   // GLOBAL statistics (after the all-reduce when data parallel)
    const float* sig_w;       // [n_sig] for the L1 term, nullable
    int n_sig;
    float lambda_edd, lambda_l1;
    float* dlogits;           // [B,3]  d(total)/d(logits) for the LOCAL patients, nullable
    float* loss_out;          // [4]: total, bce, leddi, l1   (written by block 0)
    int B;
};

// loss = BCE + lambda_edd * 10 * LEDDI + lambda_l1 * |sig_w|_1   (10_FAME.py:420-444), with
// LEDDI = mean_{i,a} sqrt( mean_{g present} (e_{i,a,g} - e_i)^2 + 1e-8 ).  Gradient flows through every mean.
__global__ void __launch_bounds__(256)
loss_fwd_bwd_kernel(const LossGradParams p) {
    __shared__ float coef[3][3][kLossSlots];  // d R_{i,a} / d e_b for a member of group g (excluding the common term)
    __shared__ float cst[3][3];               // common term of d R_{i,a} / d e_b
    __shared__ float s_leddi, s_bce, s_l1;
    const double Bt = (double)p.stats[102];
    if (threadIdx.x < 9) {
        const int i = threadIdx.x / 3, a = threadIdx.x % 3;
        const double ebar = (double)p.stats[i] / kFixErr / Bt;
        double dev[kLossSlots], ss = 0.0, sd = 0.0;
        int ng = 0;
        for (int s = 0; s < kLossSlots; ++s) {
            const long long c = p.stats[78 + a * kLossSlots + s];
            dev[s] = 0.0;
            if (c > 0) {
                dev[s] = (double)p.stats[6 + (i * 3 + a) * kLossSlots + s] / kFixErr / (double)c - ebar;
                ss += dev[s] * dev[s];
                sd += dev[s];
                ++ng;
            }
        }
        const double R = sqrt(ss / (double)ng + 1e-8);
        for (int s = 0; s < kLossSlots; ++s) {
            const long long c = p.stats[78 + a * kLossSlots + s];
            coef[i][a][s] = c > 0 ? (float)(dev[s] / ((double)c * R * (double)ng)) : 0.f;
        }
        cst[i][a] = (float)(-sd / (Bt * R * (double)ng));
        // reuse coef slot storage is not possible for R itself; keep it in a register and reduce below
        double r9 = R;
        // 9 threads of warp 0: sum R over (i,a)
        unsigned m9 = 0x1ffu;
        for (int o = 8; o > 0; o >>= 1) {
            const double other = __shfl_down_sync(m9, r9, o);
            if (threadIdx.x + o < 9) r9 += other;
        }
        if (threadIdx.x == 0) {
            s_leddi = (float)(r9 / 9.0);
            s_bce = (float)(((double)p.stats[3] + (double)p.stats[4] + (double)p.stats[5]) / kFixBce / (3.0 * Bt));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x >= 32 && threadIdx.x < 64) {
        float l1 = 0.f;
        if (p.sig_w != nullptr)
            for (int k = threadIdx.x - 32; k < p.n_sig; k += 32) l1 += fabsf(p.sig_w[k]);
        l1 = warp_sum(l1);
        if (threadIdx.x == 32) s_l1 = l1;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.loss_out != nullptr) {
        const float l1 = p.lambda_l1 * s_l1;
        // a sensitive-attribute code outside 0..7 was seen by the statistics pass (on any rank): its patient is in no
        // subgroup sum, so the loss is not the reference's -- report NaN instead of a silently different number
        const float bad = p.stats[103] != 0 ? __int_as_float(0x7fc00000) : 0.f;
        p.loss_out[0] = s_bce + p.lambda_edd * (10.0f * s_leddi) + l1 + bad;
        p.loss_out[1] = s_bce + bad;
        p.loss_out[2] = s_leddi + bad;
        p.loss_out[3] = l1;
    }
    if (p.dlogits == nullptr) return;
    const float inv3B = (float)(1.0 / (3.0 * Bt));
    const float k_edd = p.lambda_edd * 10.0f / 9.0f;
    float pwv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) pwv[i] = __ldg(p.pos_weight + i);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(p.labels) |
                          reinterpret_cast<uintptr_t>(p.attr[0]) | reinterpret_cast<uintptr_t>(p.attr[1]) |
                          reinterpret_cast<uintptr_t>(p.attr[2]) | reinterpret_cast<uintptr_t>(p.dlogits)) & 15) == 0;
    const long long per_block = (long long)blockDim.x * kPatPerThread;
    for (long long base = blockIdx.x * per_block; base < p.B; base += gridDim.x * per_block) {
        const long long b0 = base + (long long)threadIdx.x * kPatPerThread;
        PatientQuad q;
        load_patient_quad(q, p.logits, 3, p.labels, p.attr, b0, p.B, vec_ok);
        float gz[kPatPerThread][3];
#pragma unroll
        for (int u = 0; u < kPatPerThread; ++u) {
            int code[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const long long c = q.code[a][u];
                code[a] = (c < 0 || c >= kLossSlots) ? -1 : (int)c;   // out of range: member of no subgroup
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float z = q.z[u][i], y = q.y[u][i];
                const float pr = 1.0f / (1.0f + expf(-z));
                float g = (-pwv[i] * y * (1.0f - pr) + (1.0f - y) * pr) * inv3B;
                float dRde = 0.f;
#pragma unroll
                for (int a = 0; a < 3; ++a) dRde += (code[a] >= 0 ? coef[i][a][code[a]] : 0.f) + cst[i][a];
                const float d = pr - y;
                const float sgn = d > 0.f ? 1.0f : (d < 0.f ? -1.0f : 0.f);
                gz[u][i] = g + k_edd * dRde * sgn * pr * (1.0f - pr);
            }
        }
        if (q.n == kPatPerThread && vec_ok) {
            float4* dst = reinterpret_cast<float4*>(p.dlogits + 3 * b0);
            const float* gf = &gz[0][0];
#pragma unroll
            for (int k = 0; k < 3; ++k) dst[k] = make_float4(gf[4 * k], gf[4 * k + 1], gf[4 * k + 2], gf[4 * k + 3]);
        } else {
#pragma unroll
            for (int u = 0; u < kPatPerThread; ++u)
                if (u < q.n)
#pragma unroll
                    for (int i = 0; i < 3; ++i) p.dlogits[3 * (b0 + u) + i] = gz[u][i];
        }
    }
}

// ================================================================================================ focal loss (f-1)
// Text-only baseline of the reference (02_BioClinicalBERT.py:18-38, 138-152): per outcome i
//   bce = BCEWithLogits(pos_weight_i)(z, y) elementwise;  pt = exp(-bce);  fl = alpha * (1 - pt)^gamma * bce;
//   loss = sum_i mean_b fl[b, i].          d fl / d z = alpha * (gamma (1-pt)^(gamma-1) pt bce + (1-pt)^gamma) * d bce / d z,
//   d bce / d z = -pw y (1 - sigmoid z) + (1 - y) sigmoid z.
// One thread per (patient, outcome); the loss is accumulated in float64 (one atomic per block).
__global__ void __launch_bounds__(256)
focal_loss_fwd_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                          const float* __restrict__ pos_weight, float gamma, float alpha, float inv_batch, int n,
                          double* __restrict__ loss_out, float* __restrict__ dlogits,
                          const long long* __restrict__ batch_total_dev) {
    __shared__ double part[8];
    // data parallel: the mean runs over the GLOBAL batch (all-reduced patient count in device memory), so that the
    // SUM of the ranks' losses / gradients is the single-process result on the concatenated batch
    if (batch_total_dev != nullptr) inv_batch = 1.0f / (float)(*batch_total_dev);
    double acc = 0.0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
        const int i = idx % 3;
        const float z = __ldg(logits + idx), y = __ldg(labels + idx), pw = __ldg(pos_weight + i);
        const float sp = fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z)));          // softplus(-z) = -log sigmoid(z)
        const float bce = pw * y * sp + (1.0f - y) * (sp + z);
        const float pt = expf(-bce), om = 1.0f - pt;
        const float omg = gamma == 2.0f ? om * om : powf(om, gamma);
        const float omg1 = gamma == 2.0f ? om : (om > 0.f ? powf(om, gamma - 1.0f) : (gamma == 1.0f ? 1.0f : 0.f));
        acc += (double)(alpha * omg * bce);
        if (dlogits != nullptr) {
            const float sg = 1.0f / (1.0f + expf(-z));
            const float dbce = -pw * y * (1.0f - sg) + (1.0f - y) * sg;
            dlogits[idx] = alpha * (gamma * omg1 * pt * bce + omg) * dbce * inv_batch;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        atomicAdd(loss_out, t * (double)inv_batch);
    }
}

// h = relu(x) in place;  dh = pre > 0 ? dh : 0 in place  (the 256-wide hidden layer of the text-only classifier)
__global__ void relu_fwd_kernel(float* __restrict__ x, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = fmaxf(x[i], 0.f);
}
__global__ void relu_bwd_kernel(float* __restrict__ dh, const float* __restrict__ pre, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dh[i] = pre[i] > 0.f ? dh[i] : 0.f;
}

}  // namespace fame

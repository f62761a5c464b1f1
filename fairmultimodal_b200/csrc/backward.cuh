// Backward-pass and optimizer kernels of the FAME training step (train_step, 10_FAME.py:401-449): everything that
// is not a tensor-core GEMM.  These are the hand-written counterparts of what torch.autograd dispatches for the
// reference (LayerNorm / GELU / softmax / embedding / mean backward, clip_grad_norm_, AdamW).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "gemm_sm100.cuh"
#include "rowwise.cuh"
#include "dropout.cuh"

namespace fame {

template <bool kF32>
__device__ __forceinline__ void load8(const void* __restrict__ base, long long elem_off, float* f) {
    if (kF32) {
        const float* p = reinterpret_cast<const float*>(base) + elem_off;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off)), f);
    }
}
__device__ __forceinline__ void store8(__nv_bfloat16* yb, float* yf, long long elem_off, const float* o) {
    if (yb != nullptr) *reinterpret_cast<uint4*>(yb + elem_off) = float_to_bf16x8(o);
    if (yf != nullptr) {
        float4* q = reinterpret_cast<float4*>(yf + elem_off);
        q[0] = make_float4(o[0], o[1], o[2], o[3]);
        q[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// ---------------------------------------------------------------------------------------- LayerNorm backward
// y = (x - mean) * rstd * gamma + beta.   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
// dgamma += sum_rows dy * xhat,  dbeta += sum_rows dy   (f32 atomics, one per column per block)
// x: pre-LN input (bf16 or f32), stats: {mean, rstd} saved by the forward kernel.  cols % 8 == 0, <= 1024.
// CH = 8-column chunks per lane (3 for cols <= 768: 96 instead of 128 accumulator / operand registers, which lets two
// CTAs share an SM -- at one CTA the 8 resident warps kept 24 KB of loads in flight and the kernel ran at 1.8 TB/s).
// (Measured and dropped, r02: the column partials in shared memory instead of registers -- 78 registers, three CTAs per
// SM -- ran the 17 344 x 768 case in 75 instead of 48 us: the 36 16-byte shared-memory read-modify-writes per lane and
// row cost more than the third CTA brings, and without any column partials both variants take 41 us.)
template <bool kXF32, bool kDyF32, int CH = kLnMaxChunks>
__global__ void __launch_bounds__(kLnWarpsPerBlock * 32, CH <= 3 ? 2 : 1)
layernorm_bwd_kernel(const void* __restrict__ x, const void* __restrict__ dy, const float2* __restrict__ stats,
                     const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dx_f32,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int cols,
                     __nv_bfloat16* __restrict__ dx_drop, const DropCfg drop,
                     const __nv_bfloat16* __restrict__ residual = nullptr) {
    // residual (bf16, optional): the LayerNorm input was x + residual, added inside the forward LayerNorm kernel (the
    // attention-output projection of the lab tower: its K = 768 GEMM is shorter than a residual-adding epilogue)
    // dx_drop (optional): dx with the dropout mask of the layer that produced the residual branch re-applied
    // (t = residual + dropout(linear(.)): the residual path takes dx, the linear layer's backward takes dx_drop)
    __shared__ float red[2][kLnWarpsPerBlock][1024 / 4];  // staged in 4 passes of 256 columns
    const uint32_t drop_site = dx_drop != nullptr ? drop_site_seed(drop) : 0u;
    const float drop_inv = drop_inv_keep(drop.thresh16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = cols >> 3;
    float dg[CH][8], db[CH][8];
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dg[i][j] = db[i][j] = 0.f;

    for (int row = blockIdx.x * kLnWarpsPerBlock + warp; row < rows; row += gridDim.x * kLnWarpsPerBlock) {
        const float2 st = stats[row];
        float xh[CH][8], g[CH][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int ch = lane + 32 * i;
            if (ch < nchunks) {
                float xv[8], dv[8];
                load8<kXF32>(x, (long long)row * cols + 8 * ch, xv);
                load8<kDyF32>(dy, (long long)row * cols + 8 * ch, dv);
                if (residual != nullptr) {
                    float rv[8];
                    load8<false>(residual, (long long)row * cols + 8 * ch, rv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) xv[j] += rv[j];
                }
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch);
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch + 1);
                const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] = (xv[j] - st.x) * st.y;
                    g[i][j] = dv[j] * gm[j];
                    s1 += g[i][j];
                    s2 += g[i][j] * xh[i][j];
                    dg[i][j] += dv[j] * xh[i][j];
                    db[i][j] += dv[j];
                }
            }
        }
        s1 = warp_sum(s1) / (float)cols;
        s2 = warp_sum(s2) / (float)cols;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int ch = lane + 32 * i;
            if (ch < nchunks) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = st.y * (g[i][j] - s1 - xh[i][j] * s2);
                store8(dx_bf16, dx_f32, (long long)row * cols + 8 * ch, o);
                if (dx_drop != nullptr) {
                    const uint32_t rs = drop_row_seed(drop_site, (uint32_t)row);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const uint32_t bits = drop_pair_bits(rs, (uint32_t)(8 * ch + j) >> 1);
                        o[j] = (bits & 0xffffu) >= drop.thresh16 ? o[j] * drop_inv : 0.f;
                        o[j + 1] = (bits >> 16) >= drop.thresh16 ? o[j + 1] * drop_inv : 0.f;
                    }
                    store8(dx_drop, nullptr, (long long)row * cols + 8 * ch, o);
                }
            }
        }
    }
    if (dgamma == nullptr) return;
    // reduce the per-warp column partials across the block, 256 columns (one chunk index i) at a time
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[0][warp][lane * 8 + j] = dg[i][j];
            red[1][warp][lane * 8 + j] = db[i][j];
        }
        __syncthreads();
        const int c = threadIdx.x;  // 256 threads <-> 256 columns of this pass
        const int col = 256 * i + c;
        if (col < cols) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w = 0; w < kLnWarpsPerBlock; ++w) {
                a += red[0][w][c];
                b += red[1][w][c];
            }
            atomicAdd(dgamma + col, a);
            atomicAdd(dbeta + col, b);
        }
    }
}

// ---------------------------------------------------------------------------------------- GELU fwd / bwd (elementwise)
// Training keeps the pre-activation (the GEMM epilogue's fused GELU discards it).  Same polynomial as the epilogue.
__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ h, long long n8) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float f[8];
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(pre) + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
        reinterpret_cast<uint4*>(h)[i] = float_to_bf16x8(f);
    }
}
// dpre = dh * (Phi(x) + x phi(x))
__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dh,
                                __nv_bfloat16* __restrict__ dpre, long long n8) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float x[8], d[8];
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(pre) + i), x);
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(dh) + i), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float cdf = 0.5f * (1.0f + erff(x[j] * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * __expf(-0.5f * x[j] * x[j]);
            d[j] *= cdf + x[j] * pdf;
        }
        reinterpret_cast<uint4*>(dpre)[i] = float_to_bf16x8(d);
    }
}

// ---------------------------------------------------------------------------------------- column sums (bias grads)
// out[c] (+)= sum_r x[r, c].  grid = (ceil(cols / 256), row_splits); f32 atomics across row splits.
template <bool kF32>
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    auto ld1 = [&](int r) {
        return kF32 ? __ldg(reinterpret_cast<const float*>(x) + (long long)r * ld + c)
                    : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[(long long)r * ld + c]);
    };
    // 8 independent loads in flight per thread: with 32 rows (the demographic tower) a dependent row loop costs
    // 32 memory round trips (15 us measured); this costs 4
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    int r = r0;
    for (; r + 8 <= r1; r += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld1(r + u);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += v[u];
    }
    for (; r < r1; ++r) acc[0] += ld1(r);
    atomicAdd(out + c, ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7])));
}

// Vectorised variant for the large bf16 activations-gradient matrices (cols % 8 == 0, ld % 8 == 0, 16-byte aligned):
// a block covers 256 columns x a slab of rows; thread (cg, rl) streams 16-byte chunks of column group cg for rows
// rl, rl + 8, ... (4 loads in flight), the 8 row lanes are combined through shared memory, one atomic per column.
__global__ void __launch_bounds__(256)
colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows, int cols, float* __restrict__ out) {
    __shared__ float red[8][256];
    const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + cg * 8;
    const int per = (rows + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (c0 < cols) {
        int r = r0 + rl;
        for (; r + 24 < r1; r += 32) {
            float f[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + (long long)(r + 8 * u) * ld + c0)), f[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += f[u][j];
        }
        for (; r < r1; r += 8) {
            float f[8];
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + (long long)r * ld + c0)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl][cg * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < cols) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        atomicAdd(out + c, s);
    }
}

// ---------------------------------------------------------------------------------------- sequence-mean backward
// dx[b*L + l, :] = dout[b, :] / L     (bf16)
__global__ void seq_mean_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dx, int batch, int L,
                                    int cols) {
    const long long n8 = (long long)batch * L * (cols >> 3);
    const float inv = 1.0f / (float)L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % (cols >> 3));
        const long long row = i / (cols >> 3);
        const int b = (int)(row / L);
        const float4 a = __ldg(reinterpret_cast<const float4*>(dout + (long long)b * cols) + 2 * ch);
        const float4 c = __ldg(reinterpret_cast<const float4*>(dout + (long long)b * cols) + 2 * ch + 1);
        const float o[8] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv, c.x * inv, c.y * inv, c.z * inv, c.w * inv};
        reinterpret_cast<uint4*>(dx)[i] = float_to_bf16x8(o);
    }
}

// ---------------------------------------------------------------------------------------- lab embedding backward
// x[b,l,:] = lab[b,l] * w + bias + pos[l,:]:  dpos[l,:] = sum_b dx[b,l,:];  dw += sum dx * lab;  dbias += sum dx
// One CTA per SM walks over positions l; a thread owns one 8-column chunk for one slice of the batch (hidden / 8 chunks
// x S slices = the block), 8 independent 16-byte loads in flight.  dpos is the fixed-order sum of the S slice partials
// (deterministic); dw / dbias are kept in registers over all of the CTA's positions and leave as ONE atomic per
// column per CTA.  (One block per l with its own atomics -- 542 x 1 536 adds onto 1 536 addresses -- took 57 us at
// 32 x 542 x 768: the L2 atomic units serialise per address.)
constexpr int kLebThreads = 384;

__global__ void __launch_bounds__(kLebThreads)
lab_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const float* __restrict__ lab, float* __restrict__ dpos,
                     float* __restrict__ dw, float* __restrict__ dbias, int batch, int L, int hidden) {
    extern __shared__ __align__(16) float leb_part[];          // [S][hidden]
    const int nch = hidden >> 3;                               // <= kLebThreads (checked by the launcher)
    const int S = kLebThreads / nch;
    const int q = threadIdx.x / nch, ch = threadIdx.x - q * nch;
    const bool live = q < S;
    const int per = (batch + S - 1) / S;
    const int b_lo = q * per, b_hi = min(batch, b_lo + per);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    float aw[8], ab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) aw[j] = ab[j] = 0.f;
    for (int l = blockIdx.x; l < L; l += gridDim.x) {
        if (live) {
            float sp[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) sp[j] = 0.f;
            for (int b = b_lo; b < b_hi; b += 8) {
                uint4 raw[8];
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = b + u < b_hi;
                    raw[u] = ok ? __ldg(reinterpret_cast<const uint4*>(dx + ((long long)(b + u) * L + l) * hidden) + ch) : zero;
                    v[u] = ok ? __ldg(lab + (long long)(b + u) * L + l) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float f[8];
                    bf16x8_to_float(raw[u], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        sp[j] += f[j];
                        aw[j] = fmaf(f[j], v[u], aw[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) ab[j] += sp[j];
            float4* dst = reinterpret_cast<float4*>(leb_part + (long long)q * hidden + 8 * ch);
            dst[0] = make_float4(sp[0], sp[1], sp[2], sp[3]);
            dst[1] = make_float4(sp[4], sp[5], sp[6], sp[7]);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < hidden; c += kLebThreads) {
            float t = leb_part[c];
            for (int s2 = 1; s2 < S; ++s2) t += leb_part[(long long)s2 * hidden + c];
            dpos[(long long)l * hidden + c] = t;
        }
        __syncthreads();
    }
    // the CTA's dw, then dbias: slice partials through shared memory, one atomic per column
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        if (live) {
            const float* src = pass == 0 ? aw : ab;
            float4* dst = reinterpret_cast<float4*>(leb_part + (long long)q * hidden + 8 * ch);
            dst[0] = make_float4(src[0], src[1], src[2], src[3]);
            dst[1] = make_float4(src[4], src[5], src[6], src[7]);
        }
        __syncthreads();
        float* out = pass == 0 ? dw : dbias;
        for (int c = threadIdx.x; c < hidden; c += kLebThreads) {
            float t = leb_part[c];
            for (int s2 = 1; s2 < S; ++s2) t += leb_part[(long long)s2 * hidden + c];
            atomicAdd(out + c, t);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------- attention backward (elementwise)
// Per score row (one warp): recompute P = softmax(scale * S) from the f32 scores, D = sum_k P dP, and emit
//   P (bf16, feeds dV = P^T dO) and dS = scale * P * (dP - D) (bf16, feeds dQ = dS K and dK = dS^T Q).
// s / dp: f32 [rows, ld] (rows = batch * heads * seq), p / ds: bf16 [rows, ld]; columns >= seq are written as 0.
__global__ void __launch_bounds__(256)
attn_bwd_softmax_kernel(const float* __restrict__ s, const float* __restrict__ dp, __nv_bfloat16* __restrict__ p,
                        __nv_bfloat16* __restrict__ ds, long long rows, int seq, int ld, float scale) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= rows) return;
    const float* sr = s + row * ld;
    const float* dr = dp + row * ld;
    const float c = scale * 1.4426950408889634f;
    float m = -INFINITY;
    for (int k = lane; k < seq; k += 32) m = fmaxf(m, sr[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f, dot = 0.f;
    for (int k = lane; k < seq; k += 32) {
        const float e = exp2f((sr[k] - m) * c);
        sum += e;
        dot += e * dr[k];
    }
    sum = warp_sum(sum);
    dot = warp_sum(dot);
    const float inv = 1.0f / sum;
    const float D = dot * inv;
    for (int k = lane; k < ld; k += 32) {
        float pv = 0.f, dv = 0.f;
        if (k < seq) {
            pv = exp2f((sr[k] - m) * c) * inv;
            dv = scale * pv * (dr[k] - D);
        }
        p[row * ld + k] = __float2bfloat16(pv);
        ds[row * ld + k] = __float2bfloat16(dv);
    }
}

// ---------------------------------------------------------------------------------------- embedding scatter (demo tower)
// d_sum[t, :] = gradient of the pre-LayerNorm embedding sum for token t (f32).  Adds it to word[id_t] (unless id_t is the
// padding index), pos[t % seq] and type[0].
__global__ void __launch_bounds__(256)
bert_embed_bwd_kernel(const float* __restrict__ d_sum, const long long* __restrict__ ids, float* __restrict__ dword,
                      float* __restrict__ dpos, float* __restrict__ dtype0, int tokens, int seq, int hidden, int vocab,
                      int pad_idx) {
    const int t = blockIdx.x;
    long long id = ids[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
        const float g = d_sum[(long long)t * hidden + c];
        if (id != pad_idx) atomicAdd(dword + id * hidden + c, g);
        atomicAdd(dpos + (long long)(t % seq) * hidden + c, g);
        atomicAdd(dtype0 + c, g);
    }
}

// demographic tables: dE_k[clamp(id_k[b])] += dout[b] / 4
__global__ void __launch_bounds__(128)
demo_add_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ i0, const long long* __restrict__ i1,
                    const long long* __restrict__ i2, const long long* __restrict__ i3, float* __restrict__ t0,
                    float* __restrict__ t1, float* __restrict__ t2, float* __restrict__ t3, int n0, int n1, int n2, int n3,
                    int hidden) {
    const int b = blockIdx.x;
    auto clampi = [](long long v, int n) { return (long long)(v < 0 ? 0 : (v > n - 1 ? n - 1 : v)); };
    const long long r0 = clampi(i0[b], n0), r1 = clampi(i1[b], n1), r2 = clampi(i2[b], n2), r3 = clampi(i3[b], n3);
    for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
        const float g = dout[(long long)b * hidden + c] * 0.25f;
        atomicAdd(t0 + r0 * hidden + c, g);
        atomicAdd(t1 + r1 * hidden + c, g);
        atomicAdd(t2 + r2 * hidden + c, g);
        atomicAdd(t3 + r3 * hidden + c, g);
    }
}

// ---------------------------------------------------------------------------------------- small fp32 GEMM (fusion backward)
// C[m, n] = alpha * sum_k A(m, k) B(k, n) (+ C if accumulate), arbitrary element strides: A(m,k) = a[m*sam + k*sak],
// B(k,n) = b[k*sbk + n*sbn].  32 x 32 output tile per block, 16-deep k slices through shared memory.  The fusion head's
// matrices are at most 768 wide and the batch is 32 per GPU: latency matters here, not FLOPs.
// 128-deep k slices: a slice is 2 x 16 independent loads per thread, so the whole product costs K / 128 global-memory
// round trips (4 at K = 512) instead of K / 16 -- with 16-deep slices each call took 19-28 us, all of it load latency,
// and four of them sit on the chain between the towers' forward and backward passes.
constexpr int kSgKT = 128;

__global__ void __launch_bounds__(256)
sgemm_small_kernel(const float* __restrict__ a, long long sam, long long sak, const float* __restrict__ b, long long sbk,
                   long long sbn, float* __restrict__ c, long long ldc, int M, int N, int K, float alpha, int accumulate) {
    __shared__ float As[kSgKT][33], Bs[kSgKT][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads; each owns 4 rows x 1 column
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += kSgKT) {
#pragma unroll 4
        for (int i = threadIdx.x; i < kSgKT * 32; i += 256) {
            const int kk = i % kSgKT, mm = i / kSgKT;
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? a[m * sam + k * sak] : 0.f;
            const int nn = i & 31, kb = i >> 5;
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < N && k2 < K) ? b[k2 * sbk + n * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll 16
        for (int kk = 0; kk < kSgKT; ++kk) {
            const float bv = Bs[kk][tx];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[kk][ty * 4 + r], bv, acc[r]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + ty * 4 + r, n = n0 + tx;
        if (m < M && n < N) {
            float* cp = c + (long long)m * ldc + n;
            *cp = accumulate ? *cp + alpha * acc[r] : alpha * acc[r];
        }
    }
}

// fusion head backward, elementwise parts (10_FAME.py:287-296 differentiated):
//   dhid  = (W4^T dlogits) * [pre > 0]                                  (done by kernel A below, 512 wide)
//   dgate = dhid W3  -> dsig_w[k] += sum_b dgate * (w_m proj) * s(1-s) + lambda_l1 sign(sig_w[k])
//           dproj = dgate * s * w_m * [proj > 0]
__global__ void fusion_bwd_hidden_kernel(const float* __restrict__ dlogits, const float* __restrict__ w4,
                                         const float* __restrict__ pre, float* __restrict__ dhid, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 512) return;
    const int b = i / 512, k = i % 512;
    const float g = dlogits[3 * b] * w4[k] + dlogits[3 * b + 1] * w4[512 + k] + dlogits[3 * b + 2] * w4[1024 + k];
    dhid[i] = pre[i] > 0.f ? g : 0.f;
}
__global__ void fusion_bwd_gate_kernel(const float* __restrict__ dgated, const float* __restrict__ proj,
                                       const float* __restrict__ sig_w, float w0, float w1, float w2, float lambda_l1,
                                       float* __restrict__ dproj, float* __restrict__ dsig, int B,
                                       const float* __restrict__ w_dev) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 768) return;
    const float wm = w_dev != nullptr ? __ldg(w_dev + k / 256) : (k < 256 ? w0 : (k < 512 ? w1 : w2));
    const float sw = sig_w[k];
    const float s = 1.0f / (1.0f + expf(-sw));
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
        const float dg = dgated[(long long)b * 768 + k], pr = proj[(long long)b * 768 + k];
        acc += dg * (wm * pr);
        dproj[(long long)b * 768 + k] = pr > 0.f ? dg * s * wm : 0.f;
    }
    const float sgn = sw > 0.f ? 1.0f : (sw < 0.f ? -1.0f : 0.f);
    dsig[k] = acc * s * (1.0f - s) + lambda_l1 * sgn;
}

// ---------------------------------------------------------------------------------------- optimizer (K9)
// clip_grad_norm_(max_norm) + AdamW over ONE flat f32 parameter / gradient / state buffer (10_FAME.py:446-447).
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
    float acc = 0.f;
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[(n4 << 2) + threadIdx.x];
        acc += v * v;
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)red[w];
        atomicAdd(out, s);
    }
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

struct AdamWParams {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    const double* sumsq;  // global squared gradient norm (device)
    float max_norm, lr, beta1, beta2, eps, weight_decay;
    float bc1, bc2;       // 1 - beta^step
    float* grad_norm_out; // optional: total norm written by block 0
    // CUDA-graph friendly overrides (device memory, read at run time instead of being baked into the launch):
    const int* step_dev;      // optimizer step count (>= 1), nullable
    const float* hyper_dev;   // {lr, weight_decay}, nullable
    __nv_bfloat16* p_bf16;    // optional bf16 shadow of the updated parameters (tensor-core operands), nullable
};

__global__ void __launch_bounds__(256)
clip_adamw_kernel(const AdamWParams a) {
    const float total = (float)sqrt(*a.sumsq);
    const float coef = fminf(a.max_norm / (total + 1e-6f), 1.0f);   // torch.nn.utils.clip_grad_norm_
    if (a.grad_norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *a.grad_norm_out = total;
    float lr = a.lr, wd = a.weight_decay, bc1 = a.bc1, bc2 = a.bc2;
    if (a.hyper_dev != nullptr) {
        lr = a.hyper_dev[0];
        wd = a.hyper_dev[1];
    }
    if (a.step_dev != nullptr) {
        const double t = (double)*a.step_dev;
        bc1 = (float)(1.0 - pow((double)a.beta1, t));
        bc2 = (float)(1.0 - pow((double)a.beta2, t));
    }
    const float step = lr / bc1;
    const float rbc2 = rsqrtf(bc2);
    const float decay = 1.0f - lr * wd;
    const long long n4 = a.n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 p = reinterpret_cast<float4*>(a.p)[i];
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.g) + i);
        float4 m = reinterpret_cast<float4*>(a.m)[i];
        float4 v = reinterpret_cast<float4*>(a.v)[i];
        float* pp = &p.x; float* mm = &m.x; float* vv = &v.x; const float* gg = &g4.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float g = gg[j] * coef;
            pp[j] *= decay;
            mm[j] = a.beta1 * mm[j] + (1.0f - a.beta1) * g;
            vv[j] = a.beta2 * vv[j] + (1.0f - a.beta2) * g * g;
            pp[j] -= step * mm[j] / (sqrtf(vv[j]) * rbc2 + a.eps);
        }
        reinterpret_cast<float4*>(a.p)[i] = p;
        reinterpret_cast<float4*>(a.m)[i] = m;
        reinterpret_cast<float4*>(a.v)[i] = v;
        if (a.p_bf16 != nullptr) {
            uint2 o;
            o.x = pack2(p.x, p.y);
            o.y = pack2(p.z, p.w);
            reinterpret_cast<uint2*>(a.p_bf16)[i] = o;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float g = a.g[i] * coef;
        float p = a.p[i] * decay;
        const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
        p -= step * m / (sqrtf(v) * rbc2 + a.eps);
        a.p[i] = p; a.m[i] = m; a.v[i] = v;
        if (a.p_bf16 != nullptr) a.p_bf16[i] = __float2bfloat16(p);
    }
}

// AdamW for a region whose gradient AND Adam moments are identically zero (the query / key projection weights of the
// demographic BERT: one key per sequence, so the softmax is 1 and d loss / d (Q, K) = 0 exactly; 14.2 M of the 97.9 M
// parameters).  With g = m = v = 0 the update term is 0 / (0 + eps) = 0 and the step reduces to the decoupled weight
// decay p <- p (1 - lr wd): 8 bytes per parameter instead of 30 (+ 2 for the bf16 shadow).
__global__ void __launch_bounds__(256)
decay_only_kernel(float* __restrict__ p, long long n, float lr, float wd, const float* __restrict__ hyper_dev,
                  __nv_bfloat16* __restrict__ p_bf16) {
    if (hyper_dev != nullptr) {
        lr = hyper_dev[0];
        wd = hyper_dev[1];
    }
    const float decay = 1.0f - lr * wd;
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(p)[i];
        v.x *= decay; v.y *= decay; v.z *= decay; v.w *= decay;
        reinterpret_cast<float4*>(p)[i] = v;
        if (p_bf16 != nullptr) {
            uint2 o;
            o.x = pack2(v.x, v.y);
            o.y = pack2(v.z, v.w);
            reinterpret_cast<uint2*>(p_bf16)[i] = o;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float v = p[i] * decay;
        p[i] = v;
        if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16(v);
    }
}

// f32 -> bf16 cast of the flat parameter buffer regions that feed the tensor-core GEMMs
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = __float2bfloat16(x[i]);
}

}  // namespace fame

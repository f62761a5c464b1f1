// K1: bf16 GEMM on tcgen05 tensor cores, accumulators in TMEM, operands staged by TMA.
//
//   C[b][m, n] = epi( sum_k A[b](m, k) * B[b](n, k) )        b = optional 2-level batch index (b0, b1)
//
// One kernel serves the forward, data-gradient and weight-gradient products of every nn.Linear on the FAME hot path
// and the five batched products of the attention backward, through the operand "major":
//   K-major  operand: stored [rows = m or n, cols = k]  (k contiguous)  -- forward X and W, dY in dgrad
//   MN-major operand: stored [rows = k, cols = m or n]  (m/n contiguous) -- W in dgrad, dY and X in wgrad, V/K/Q/dO
//   forward  Y  = X . W^T        A = X  (K),  B = W  (K)          HF modeling_bert.py:179-181,295,340,353; 10_FAME.py:214
//   dgrad    dX = dY . W         A = dY (K),  B = W  (MN)
//   wgrad    dW = dY^T . X       A = dY (MN), B = X  (MN)         contraction over tokens
// Epilogue: + bias[n], GELU(erf) / ReLU, + residual (bf16 or f32) or ReLU-mask by another tensor; bf16 output through
// swizzled smem + TMA store, f32 output (small matrices, weight gradients) through direct vector stores.
//
// CTA pairs (cluster of 2, tcgen05 cta_group::2): the two CTAs of a pair own vertically adjacent 128-row output
// tiles of the same n block and execute ONE 256x256x16 MMA per step, issued by the pair leader.  Each CTA stages its
// own A tile (128 x 64) and only HALF of the B tile (128 of the 256 n rows); the tensor cores read the other half from
// the peer's shared memory.  Per k-block a CTA therefore moves 32 KB through TMA and its MMAs read 32 KB of shared
// memory instead of 48 + 48 KB: with single-CTA 128x256 MMAs the kernel sat at 1150-1380 TFLOP/s on every shape
// (tensor pipe 65-80 % active) because operand staging (TMA writes + UMMA reads) saturated the 128 B/clk
// shared-memory port; halving the L2 traffic alone (TMA multicast of B, tried first) changed nothing.
// Barriers: both producers credit the LEADER's full barrier (cta_group::2 TMA); the leader's tcgen05.commit
// multicasts to the empty / tmem_full barriers of both CTAs; the epilogue warps of both CTAs arrive on the leader's
// tmem_empty barrier.
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0   : TMA producer   (one lane)  global -> smem ring, kStages x {A 128x64, B 256x64} bf16, SW128
//   warp 1   : MMA issuer     (one lane)  tcgen05.mma 128x256x16, 4 per k-block, commit -> frees ring slot
//   warp 2   : TMEM allocator (512 columns = 2 accumulator buffers of 256 f32 columns)
//   warps 4-11: epilogue, two warpgroups of 128 threads, each owning 128 accumulator columns; runs concurrently
//              with the next tile's MMAs (double-buffered TMEM).
#pragma once
#include "sm100_ptx.cuh"
#include "dropout.cuh"

namespace fame {

constexpr int kGemmBM = 128;
constexpr int kGemmBN = 256;
constexpr int kGemmBK = 64;
constexpr int kGemmStages = 5;                      // 5 x 32 KB ring + 4 x 16 KB store staging = 224 KB
constexpr int kGemmABytes = kGemmBM * kGemmBK * 2;        // 16 KB: this CTA's 128 rows of A
constexpr int kGemmBBytes = (kGemmBN / 2) * kGemmBK * 2;  // 16 KB: this CTA's half (128 n rows) of the pair's B tile
constexpr int kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr int kGemmCBoxBytes = 128 * 64 * 2;        // 16 KB staging box per epilogue warpgroup
constexpr int kGemmSubTile = 64 * 64 * 2;           // 8 KB: one {64 mn x 64 k} box of an MN-major operand
constexpr int kGemmThreads = 384;
// two staging boxes per epilogue warpgroup (one per 64-column chunk of its half): a chunk's TMA store may still be
// reading its box while the next chunk is written into the other one
constexpr int kGemmSmemBytes = kGemmStages * kGemmStageBytes + 4 * kGemmCBoxBytes + 1024 /*align slack*/ + 256;

enum { kActNone = 0, kActGelu = 1, kActRelu = 2 };
enum { kResNone = 0, kResAddBf16 = 1, kResAddF32 = 2, kResReluMaskBf16 = 3, kResGeluBwdBf16 = 4 /* skinny only */ };

struct GemmParams {
    int M, N, K;                    // per-batch output rows / cols and contraction length
    int nb0, nb1;                   // batch counts (1, 1 when unbatched); tile -> (b0, b1, m_blk, n_blk)
    const float* bias;              // [N] or nullptr
    const void* residual;           // [nb0][nb1][M, ldr] or nullptr
    long long ldr, rs_b0, rs_b1;    // leading dimension and batch strides of the residual (elements)
    int res_mode;
    void* y;                        // f32 output only: [nb0][nb1][M, ldy]; bf16 output goes through tmap_c
    long long ldy, ys_b0, ys_b1;
    int act;
    int y_f32;
    float alpha;                    // scales the accumulator before bias / activation
    int split_k;                    // >= 1.  > 1: the K range is cut into split_k slices, one tile-task per slice;
    int kb_per_split;               //        f32 output only
    int atomic_out;                 // 1: tiles are ADDED to y with float4 atomics (split-K / accumulate mode)
    DropCfg drop;                   // dropout after the activation, before the residual add (kDrop kernels only)
};

// erf-GELU (HF "gelu", modeling_bert.py:339-342):  0.5 x (1 + erf(x / sqrt 2)).
// erf(x/sqrt2) is evaluated as xc * Q(xc^2) with xc = clamp(x, +-3 sqrt2), Q of degree 8 fitted by
// scripts/fit_gelu_poly.py (max |erf error| 3.1e-5, max |gelu error| 6.7e-5 -- far below the bf16 output ulp).
// 13 FP32 instructions and no MUFU per element: the libdevice erff (~30 instructions, divergent branches) made the
// 3072-wide FFN epilogue slower than its K = 768 main loop.
__device__ __forceinline__ float gelu_erf(float x) {
    constexpr float c0 = 7.978010774e-01f, c1 = -1.326614171e-01f, c2 = 1.961016096e-02f, c3 = -2.209631959e-03f,
                    c4 = 1.854783768e-04f, c5 = -1.111321126e-05f, c6 = 4.429220439e-07f, c7 = -1.040250019e-08f,
                    c8 = 1.080706497e-10f;
    const float xc = fminf(fmaxf(x, -4.242640687f), 4.242640687f);
    const float u = xc * xc;
    float q = fmaf(c8, u, c7);
    q = fmaf(q, u, c6);
    q = fmaf(q, u, c5);
    q = fmaf(q, u, c4);
    q = fmaf(q, u, c3);
    q = fmaf(q, u, c2);
    q = fmaf(q, u, c1);
    q = fmaf(q, u, c0);
    const float h = 0.5f * x;
    return fmaf(h, xc * q, h);
}

// Two elements at a time on the packed FP32 pipe (FFMA2 / FMUL2): 12 packed instructions per pair instead of 13
// scalar ones per element.  The 3072-wide FFN1 epilogue is otherwise slower than its K = 768 main loop.
__device__ __forceinline__ void gelu_erf_x2(float& a, float& b) {
    constexpr float c0 = 7.978010774e-01f, c1 = -1.326614171e-01f, c2 = 1.961016096e-02f, c3 = -2.209631959e-03f,
                    c4 = 1.854783768e-04f, c5 = -1.111321126e-05f, c6 = 4.429220439e-07f, c7 = -1.040250019e-08f,
                    c8 = 1.080706497e-10f;
    const float xa = fminf(fmaxf(a, -4.242640687f), 4.242640687f);
    const float xb = fminf(fmaxf(b, -4.242640687f), 4.242640687f);
    const unsigned long long x = f32x2_pack(xa, xb);
    const unsigned long long u = f32x2_mul(x, x);
    unsigned long long q = f32x2_fma(f32x2_pack(c8, c8), u, f32x2_pack(c7, c7));
    q = f32x2_fma(q, u, f32x2_pack(c6, c6));
    q = f32x2_fma(q, u, f32x2_pack(c5, c5));
    q = f32x2_fma(q, u, f32x2_pack(c4, c4));
    q = f32x2_fma(q, u, f32x2_pack(c3, c3));
    q = f32x2_fma(q, u, f32x2_pack(c2, c2));
    q = f32x2_fma(q, u, f32x2_pack(c1, c1));
    q = f32x2_fma(q, u, f32x2_pack(c0, c0));
    const unsigned long long h = f32x2_mul(f32x2_pack(a, b), f32x2_pack(0.5f, 0.5f));
    const unsigned long long r = f32x2_fma(h, f32x2_mul(x, q), h);
    f32x2_unpack(r, a, b);
}

// kDrop: a separate instantiation carries the dropout epilogue, so the inference / no-dropout kernels are unchanged.
template <bool kAMn, bool kBMn, bool kDrop = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must start on 1024-byte boundaries.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kGemmStages * kGemmABytes;
    uint8_t* smem_c = smem + kGemmStages * kGemmStageBytes;  // 4 x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 4 * kGemmCBoxBytes);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + kGemmStages;        // [kStages]
    uint64_t* tmem_full = bars + 2 * kGemmStages;    // [2]
    uint64_t* tmem_empty = bars + 2 * kGemmStages + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kGemmStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const uint32_t rank = cluster_ctarank();          // 0 / 1 inside the CTA pair
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int m_tiles = (p.M + kGemmBM - 1) / kGemmBM;
    const int m_pairs = (m_tiles + 1) >> 1;           // a pair takes m blocks 2 mp and 2 mp + 1 (the second may be empty)
    const int n_tiles = (p.N + kGemmBN - 1) / kGemmBN;
    const int tiles_per_batch = m_pairs * n_tiles;
    // task = (output tile, k slice); the slice index is innermost so that the slices of one tile run concurrently
    const int num_tiles = tiles_per_batch * p.nb0 * p.nb1 * p.split_k;
    const int num_kb = (p.K + kGemmBK - 1) / kGemmBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kGemmStages; ++i) {
            mbar_init(&full_bar[i], 1);    // used in the leader only: its own arrive.expect_tx for both CTAs' bytes
            mbar_init(&empty_bar[i], 1);   // one multicast tcgen05.commit of the leader per use
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 16); // leader only: one arrival per epilogue warp of BOTH CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_pair(tmem_slot, 512);   // executed by the same warp of both CTAs of the pair
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();    // the peer's barriers are initialised before anything is multicast into its shared memory
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int task = pair; task < num_tiles; task += npairs) {
                const int tile = task / p.split_k, ks = task % p.split_k;
                const int bidx = tile / tiles_per_batch, t2 = tile % tiles_per_batch;
                const int b0 = bidx / p.nb1, b1 = bidx % p.nb1;
                const int m_blk = 2 * (t2 / n_tiles) + (int)rank, n_blk = t2 % n_tiles;
                const int kb_lo = ks * p.kb_per_split, kb_hi = min(num_kb, kb_lo + p.kb_per_split);
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    // both CTAs' loads complete on the leader's barrier, which expects the bytes of the whole pair
                    const uint32_t fb = map_to_cta(smem_u32(&full_bar[stage]), 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kGemmStageBytes);
                    uint8_t* sa = smem_a + stage * kGemmABytes;
                    uint8_t* sb = smem_b + stage * kGemmBBytes;
                    if (kAMn) {
#pragma unroll
                        for (int s = 0; s < kGemmBM / 64; ++s)
                            tma_load_4d_pair(sa + s * kGemmSubTile, &tmap_a, fb, m_blk * kGemmBM + 64 * s, kb * kGemmBK,
                                             b1, b0, kEvictNormal);
                    } else {
                        tma_load_4d_pair(sa, &tmap_a, fb, kb * kGemmBK, m_blk * kGemmBM, b1, b0, kEvictNormal);
                    }
                    // this CTA's half of the B tile: n rows [n_blk * 256 + rank * 128, + 128)
                    const int n_row0 = n_blk * kGemmBN + (int)rank * (kGemmBN / 2);
                    if (kBMn) {
#pragma unroll
                        for (int s = 0; s < kGemmBN / 128; ++s)
                            tma_load_4d_pair(sb + s * kGemmSubTile, &tmap_b, fb, n_row0 + 64 * s, kb * kGemmBK, b1, b0,
                                             kEvictLast);
                    } else {
                        tma_load_4d_pair(sb, &tmap_b, fb, kb * kGemmBK, n_row0, b1, b0, kEvictLast);
                    }
                    if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ------------------------------------------------ MMA issuer (pair leader only): 256 x 256 x 16 per step
            // the whole warp runs the loop so that descriptors / TMEM addresses are warp-uniform (uniform registers, one
            // instruction per tcgen05.mma); only the elected lane issues
            const bool leader = elect_one();
            constexpr uint32_t idesc = make_idesc_bf16(2 * kGemmBM, kGemmBN, kAMn ? 1 : 0, kBMn ? 1 : 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int task = pair; task < num_tiles; task += npairs) {
                const int kb_lo = (task % p.split_k) * p.kb_per_split, kb_hi = min(num_kb, kb_lo + p.kb_per_split);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kGemmBN;
                for (int kb = kb_lo; kb < kb_hi; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kGemmABytes);
                    const uint32_t b_addr = smem_u32(smem_b + stage * kGemmBBytes);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < kGemmBK / 16; ++k) {
                            // K-major: 16 k-elements = 32 B inside the 128 B row; SBO = 8 rows.  MN-major: 16 k-rows of
                            // 128 B = 2 KB; LBO = distance between 64-wide MN sub-tiles, SBO = 8 k-rows.
                            const uint64_t adesc = kAMn ? make_smem_desc_sw128(a_addr + k * 2048, kGemmSubTile, 1024)
                                                        : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                            const uint64_t bdesc = kBMn ? make_smem_desc_sw128(b_addr + k * 2048, kGemmSubTile, 1024)
                                                        : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                            umma_bf16_ss_pair(d_tmem, adesc, bdesc, idesc, (kb != kb_lo) || (k != 0));
                        }
                        umma_commit_pair(&empty_bar[stage], 3);   // frees the slot in both CTAs' producers
                    }
                    if (++stage == kGemmStages) { stage = 0; phase ^= 1; }
                }
                if (leader) umma_commit_pair(&tmem_full[acc], 3);     // accumulators of both CTAs are complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue (2 warpgroups x 4 warps)
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int half = (warp - 4) >> 2;  // which 128-column half of the accumulator (= warpgroup)
        const int r_local = q * 32 + lane; // row inside the tile
        const bool wg_leader = (q == 0 && lane == 0);
        uint8_t* cbox0 = smem_c + half * 2 * kGemmCBoxBytes;       // boxes [2 half, 2 half + 1]: chunk cc uses box cc
        const uint32_t crow_s0 = smem_u32(cbox0 + r_local * 128);  // shared-window address: st.shared, not a generic store
        const int sw = r_local & 7;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t drop_site = kDrop ? drop_site_seed(p.drop) : 0u;
        for (int task = pair; task < num_tiles; task += npairs) {
            const int tile = task / p.split_k;
            const bool first_slice = (task % p.split_k) == 0;   // bias / residual are added by one slice only
            const int bidx = tile / tiles_per_batch, t2 = tile % tiles_per_batch;
            const int b0 = bidx / p.nb1, b1 = bidx % p.nb1;
            const int m_blk = 2 * (t2 / n_tiles) + (int)rank, n_blk = t2 % n_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * kGemmBM + r_local;
            const bool row_ok = row < p.M;
            const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * kGemmBN + half * 128;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int col0 = n_blk * kGemmBN + half * 128 + cc * 64;
                const bool active = col0 < p.N;  // warpgroup-uniform
                float v[64];
                const bool with_bias = p.bias != nullptr && first_slice;
                if (active) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_x32(t_addr + cc * 64, r0);
                    tmem_ld_x32(t_addr + cc * 64 + 32, r1);
                    tmem_ld_wait();
                    if (with_bias && col0 + 64 <= p.N) {
                        // alpha * acc + bias as one FMA per element (the common case: a full 64-column chunk)
#pragma unroll
                        for (int j = 0; j < 64; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                            const uint32_t* r = j < 32 ? r0 + j : r1 + (j - 32);
                            v[j] = fmaf(__uint_as_float(r[0]), p.alpha, b.x);
                            v[j + 1] = fmaf(__uint_as_float(r[1]), p.alpha, b.y);
                            v[j + 2] = fmaf(__uint_as_float(r[2]), p.alpha, b.z);
                            v[j + 3] = fmaf(__uint_as_float(r[3]), p.alpha, b.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            v[j] = __uint_as_float(r0[j]) * p.alpha;
                            v[32 + j] = __uint_as_float(r1[j]) * p.alpha;
                        }
                    }
                }
                if (cc == 1) {
                    // accumulator fully read: hand the TMEM buffer back to the MMA warp before the math
                    tc_fence_before();
                    __syncwarp();
                    // relaxed arrival: a release-scoped one made every epilogue warp wait here for its own outstanding
                    // stores (ncu r02: MEMBAR.ALL.CTA + ERRBAR = 12 % of the FFN1 launch's samples)
                    if (lane == 0) mbar_arrive_cluster_relaxed(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
                }
                if (!active) continue;
                if (with_bias && col0 + 64 > p.N) {      // ragged last chunk: bias added under the column guard
#pragma unroll
                    for (int j = 0; j < 64; j += 4) {
                        if (col0 + j < p.N) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                    }
                }
                if (p.act == kActGelu) {
#pragma unroll
                    for (int j = 0; j < 64; j += 2) gelu_erf_x2(v[j], v[j + 1]);
                } else if (p.act == kActRelu) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.0f);
                }
                if (kDrop && p.drop.thresh16 != 0u) {
                    // nn.Dropout on the layer output (after the activation, before the residual add)
                    const uint32_t rs = drop_row_seed(drop_site, (uint32_t)row);
                    const float inv = drop_inv_keep(p.drop.thresh16);
                    if (p.drop.group_shift == 0) {
#pragma unroll
                        for (int j = 0; j < 64; j += 2) {
                            const uint32_t bits = drop_pair_bits(rs, (uint32_t)(col0 + j) >> 1);
                            v[j] = (bits & 0xffffu) >= p.drop.thresh16 ? v[j] * inv : 0.f;
                            v[j + 1] = (bits >> 16) >= p.drop.thresh16 ? v[j + 1] * inv : 0.f;
                        }
                    } else {
                        // group_shift >= 6 (host-checked): the 64 columns of this chunk (col0 % 64 == 0) share one draw
                        const float m = drop_keep(rs, (uint32_t)col0 >> p.drop.group_shift, p.drop.thresh16) ? inv : 0.f;
#pragma unroll
                        for (int j = 0; j < 64; ++j) v[j] *= m;
                    }
                }
                if (p.res_mode != kResNone && row_ok && first_slice) {
                    const long long roff = (long long)b0 * p.rs_b0 + (long long)b1 * p.rs_b1 + (long long)row * p.ldr + col0;
                    if (p.res_mode == kResAddF32) {
                        const float* rp = reinterpret_cast<const float*>(p.residual) + roff;
#pragma unroll
                        for (int j = 0; j < 64; j += 4) {
                            if (col0 + j < p.N) {
                                const float4 rv = __ldg(reinterpret_cast<const float4*>(rp + j));
                                v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
                            }
                        }
                    } else {
                        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
                        const bool mask = p.res_mode == kResReluMaskBf16;
#pragma unroll
                        for (int j = 0; j < 64; j += 8) {
                            if (col0 + j < p.N) {
                                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + j));
                                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float2 f = __bfloat1622float2(h2[t]);
                                    if (mask) {  // ReLU backward: keep the gradient where the saved activation is > 0
                                        v[j + 2 * t] = f.x > 0.f ? v[j + 2 * t] : 0.f;
                                        v[j + 2 * t + 1] = f.y > 0.f ? v[j + 2 * t + 1] : 0.f;
                                    } else {
                                        v[j + 2 * t] += f.x;
                                        v[j + 2 * t + 1] += f.y;
                                    }
                                }
                            }
                        }
                    }
                }
                if (p.y_f32) {
                    if (row_ok) {
                        float* yp = reinterpret_cast<float*>(p.y) + (long long)b0 * p.ys_b0 + (long long)b1 * p.ys_b1 +
                                    (long long)row * p.ldy + col0;
                        if (p.atomic_out) {
#pragma unroll
                            for (int j = 0; j < 64; j += 4) {
                                if (col0 + j < p.N)
                                    atomicAdd(reinterpret_cast<float4*>(yp + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 64; j += 4) {
                                if (col0 + j < p.N)
                                    *reinterpret_cast<float4*>(yp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            }
                        }
                    }
                } else {
                    // the TMA store issued out of THIS box two chunks ago must have finished reading it (the one out of
                    // the other box may still be in flight)
                    uint8_t* cbox = cbox0 + cc * kGemmCBoxBytes;
                    const uint32_t crow_s = crow_s0 + cc * kGemmCBoxBytes;
                    if (wg_leader) tma_store_wait_read_1();
                    named_bar_sync(1 + half, 128);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint4 o;
                        o.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
                        o.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
                        o.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
                        o.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"      // SW128: 16-byte chunk ^= row % 8
                                     ::"r"(crow_s + ((c ^ sw) << 4)), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + half, 128);
                    if (wg_leader) {
                        tma_store_4d(&tmap_c, cbox, col0, m_blk * kGemmBM, b1, b0);
                        tma_store_commit();
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (wg_leader) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();    // neither CTA leaves while the other may still signal barriers in its shared memory
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace fame

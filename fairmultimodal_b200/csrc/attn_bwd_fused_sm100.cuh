// K2c: fused attention backward on tcgen05 / TMEM -- P and dS never leave the SM.
//
// autograd of F.scaled_dot_product_attention inside nn.MultiheadAttention of BEHRTModel_Lab (10_FAME.py:212-215, reached
// by total_loss.backward() at 10_FAME.py:445), per (sequence, head):
//     P  = softmax(Q K^T * scale)                        dV = P^T dO
//     dP = dO V^T                                         dS = scale * P * (dP - rowsum(dO * O))
//     dQ = dS K,   dK = dS^T Q
// attn_bwd_pds_kernel (attn_bwd_sm100.cuh) wrote P and dS to HBM ([B, H, L, L] bf16 each: 2 x 4.8 GB at 1024 patients)
// and three batched GEMMs read them back.  Here the two score products, the softmax backward AND the products that
// consume P / dS run in one kernel: the bf16 P / dS tiles are written back into TMEM over the fp32 scores they came
// from and feed the next tcgen05.mma as its A operand (the forward kernel's S -> P aliasing, attn_pair_sm100.cuh).
//
// One template, two passes (no cross-CTA reduction, no atomics, no fp32 scratch):
//   kKV = false  "dQ pass"   stationary rows = 128 QUERIES (Q_i, dO_i in shared memory), streamed = 64-key half blocks
//                            (K_jh, V_jh):   S = Q_i K_jh^T,  dP = dO_i V_jh^T  ->  dS  ->  dQ_i += dS K_jh
//   kKV = true   "dK/dV pass" stationary rows = 128 KEYS (K_j, V_j), streamed = 64-query half blocks (Q_ih, dO_ih):
//                            S^T = K_j Q_ih^T, dP^T = V_j dO_ih^T -> P^T, dS^T -> dV_j += P^T dO_ih, dK_j += dS^T Q_ih
// Both passes recompute the scores (7 tile products per 128 x 128 pair instead of 5 with a dQ reduction across CTAs);
// the softmax statistics come from the forward's row log-sum-exp, so there is no max / sum pass.
//
//   warp 16 lane 0 : TMA producer   stationary pair per item; streamed half-block ring (4 stages x 32 KB), SW128 boxes
//   warp 17        : MMA issuer     scores of half block g -> TMEM stage g & 1 (cols [128 s, +64) and [128 s + 64, +64));
//                                   accumulators at cols 256 (dQ | dV) and 384 (dK); runs two half blocks ahead
//   warps 0-7      : softmax-backward group of even half blocks: thread = one stationary row x 32 of the 64 streamed
//                    columns (warps q and q + 4 share TMEM lane quadrant q and take columns [0, 32) / [32, 64))
//   warps 8-15     : group of odd half blocks
// Sixteen softmax warps (four per scheduler) instead of eight: with two per scheduler the S -> dS -> next-S chain of a
// stage was exposed (ncu r02: 2 300 - 2 700 clk per half block against 576 / 768 clk of tensor pipe, the warps' issue
// slots 25 % used).  A thread writes its bf16 results over columns it has itself read, so no ordering is needed
// between the two warps of a quadrant: P of columns [32 h, 32 h + 32) lands in [32 h, 32 h + 16).
// TMEM: 2 stages x 128 score columns + 2 x 96 accumulator columns.  Streamed traffic: 32 KB per 64-row half block
// (head_dim 96 is loaded as two 64-column boxes) against 576 / 768 clk of tensor pipe: ~43-57 B/clk/SM from L2 -- the
// kernel sits at the L2 -> SM bandwidth of the chip (~6.3 KB/clk), which is why the ring is four deep.
#pragma once
#include "dropout.cuh"
#include "sm100_ptx.cuh"
#include "attn_common.cuh"

namespace fame {

// Measurement switches (profiles/r02_attn_bwd_switch_experiments.log, r02_attn_bwd_event_trace_cta0.log) are compiled in
// only with -DFAME_ATTN_INSTRUMENT: left in as run-time branches they cost the kernel 15 % (18.9 -> 22 ms per config-3
// step).  With the macro, FAME_ATTN_DEBUG selects: 1 skip the exponential / FMA math, 2 also skip the TMEM score loads,
// 4 skip the accumulating MMAs, 8 stream half the bytes, 16 record the event trace below (results are then wrong).
#ifdef FAME_ATTN_INSTRUMENT
#define AF_DBG(bit) ((p.debug & (bit)) != 0)
#else
#define AF_DBG(bit) (false)
#endif
// AF_DBG(16): CTA 0 records (event, clock64) pairs of its MMA issuer and of softmax warp 0
__device__ long long g_af_trace[2][4096];
__device__ __forceinline__ void af_trace(int who, int& n, int ev) {
    if (n < 2047) {
        g_af_trace[who][2 * n] = ev;
        g_af_trace[who][2 * n + 1] = clock64();
        ++n;
    }
}

constexpr int kAfThreads = 576;          // 16 softmax-backward warps + TMA producer + MMA issuer

template <int D>
struct AfCfg {
    static constexpr int kBoxes = (D + 63) / 64;
    static constexpr int kStatBytes = kBoxes * kFaBoxBytes;        // one stationary operand: 128 rows
    static constexpr int kHalfBoxBytes = 64 * 64 * 2;              // one streamed box: 64 rows x 64 bf16 columns
    static constexpr int kStreamBytes = kBoxes * kHalfBoxBytes;    // one streamed operand: 64 rows
    static constexpr int kStages = 4;
    static constexpr int kStageBytes = 2 * kStreamBytes;
    static constexpr int kStatsBytes = 2 * 2 * 3 * 64 * 4;         // [warpgroup][buffer][lse | delta * scale | row seed][64]
    static constexpr int kSmemBytes = 2 * kStatBytes + kStages * kStageBytes + kStatsBytes + 1024 /*barriers*/ + 1024 /*align*/;
};

struct AfParams {
    const float* lse;          // [batch, heads, seq]  row log-sum-exp of the forward, log2 units of the scaled scores
    const float* delta;        // [batch, heads, seq]  rowsum(dO * O)
    __nv_bfloat16* dqkv;       // [batch * seq, ld]    packed gradient: dQ | dK | dV, head-major inside each third
    long long ld;
    int batch, seq, heads;
    int q_col0, k_col0, v_col0;
    float scale, scale_log2e;
    DropCfg drop;              // dropout of the attention probabilities in the forward (kDrop instantiations only)
    int debug;                 // measurement switches (FAME_ATTN_DEBUG, results are then WRONG): 1 skip the exponential / FMA
                               // math, 2 also skip the TMEM score loads, 4 skip the accumulating MMAs, 8 stream half the bytes
};

// kDrop: the forward dropped entries of P (mask m, scale c = 1 / (1 - p)):  O = (m c P) V.  Then
//   dV = (m c P)^T dO;   dP = m c (dO V^T);   dS = scale P (dP - delta),  delta = rowsum(dO * O) = sum_k P_k dP_k.
// Mask row = (sequence, head, query), mask unit = key, two keys per 32-bit hash (dropout.cuh).
template <int D, bool kKV, bool kDrop>
__global__ void __launch_bounds__(kAfThreads, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmap_qkv128, const __grid_constant__ CUtensorMap tmap_do128,
                      const __grid_constant__ CUtensorMap tmap_qkv64, const __grid_constant__ CUtensorMap tmap_do64,
                      const AfParams p, const int num_items, const int rtiles) {
    using Cfg = AfCfg<D>;
    constexpr int NB = Cfg::kBoxes;
    constexpr int ST = Cfg::kStages;
    constexpr uint32_t kColAcc0 = 256, kColAcc1 = 384;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a1 = smem;                              // stationary operand of the S product   (Q_i | K_j)
    uint8_t* smem_a2 = smem_a1 + Cfg::kStatBytes;         // stationary operand of the dP product  (dO_i | V_j)
    uint8_t* smem_b = smem_a2 + Cfg::kStatBytes;          // [ST][B1 | B2]: streamed (K_jh, V_jh | Q_ih, dO_ih)
    float* smem_stats = reinterpret_cast<float*>(smem_b + ST * Cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_stats) + Cfg::kStatsBytes);
    uint64_t* a_full = bars;                   // stationary pair landed
    uint64_t* a_empty = bars + 1;              // last score product of the item done
    uint64_t* b_full = bars + 2;               // [ST]
    uint64_t* b_empty = b_full + ST;           // [ST]
    uint64_t* sd_full = b_empty + ST;          // [2] scores of a half block complete in TMEM stage s
    uint64_t* ds_full = sd_full + 2;           // [2] bf16 P / dS of stage s written by its warpgroup (4 warp arrivals)
    uint64_t* acc_full = ds_full + 2;          // accumulators of the item complete
    uint64_t* acc_empty = acc_full + 1;        // accumulators read by the epilogue (8 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.seq;
    const int nhb = (S + 63) >> 6;             // streamed half blocks per item

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv128);
        tma_prefetch_desc(&tmap_do128);
        tma_prefetch_desc(&tmap_qkv64);
        tma_prefetch_desc(&tmap_do64);
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < ST; ++i) {
            mbar_init(&b_full[i], 1);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sd_full[i], 1);
            mbar_init(&ds_full[i], 8);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 16);
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // columns of the stationary / streamed operands inside the packed tensors
    const int a1_col0 = kKV ? p.k_col0 : p.q_col0;        // a2: V (qkv) | dO (dctx, column 0)
    const int b1_col0 = kKV ? p.q_col0 : p.k_col0;        // b2: dO (dctx) | V (qkv)

    if (warp == 16) {
        if (lane == 0) {
            // ------------------------------------------------------------------------------------ TMA producer
            int st = 0;
            uint32_t aph = 0, bph = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int rt = item % rtiles, bh = item / rtiles;
                const int h = bh % p.heads, b = bh / p.heads;
                const int row0 = b * S;
                mbar_wait(a_empty, aph ^ 1);
                aph ^= 1;
                mbar_arrive_expect_tx(a_full, 2 * Cfg::kStatBytes);
#pragma unroll
                for (int x = 0; x < NB; ++x) {
                    tma_load_2d(smem_a1 + x * kFaBoxBytes, &tmap_qkv128, a_full, a1_col0 + h * D + x * 64, row0 + rt * 128,
                                kEvictFirst);
                    if (kKV)
                        tma_load_2d(smem_a2 + x * kFaBoxBytes, &tmap_qkv128, a_full, p.v_col0 + h * D + x * 64,
                                    row0 + rt * 128, kEvictFirst);
                    else
                        tma_load_2d(smem_a2 + x * kFaBoxBytes, &tmap_do128, a_full, h * D + x * 64, row0 + rt * 128,
                                    kEvictFirst);
                }
                for (int hb = 0; hb < nhb; ++hb) {
                    mbar_wait(&b_empty[st], bph ^ 1);
                    const int nbx = AF_DBG(8) ? 1 : NB;      // measurement: half of the streamed bytes
                    mbar_arrive_expect_tx(&b_full[st], 2 * nbx * Cfg::kHalfBoxBytes);
                    uint8_t* b1 = smem_b + st * Cfg::kStageBytes;
                    uint8_t* b2 = b1 + Cfg::kStreamBytes;
#pragma unroll
                    for (int x = 0; x < NB; ++x) {
                        if (x >= nbx) break;
                        tma_load_2d(b1 + x * Cfg::kHalfBoxBytes, &tmap_qkv64, &b_full[st], b1_col0 + h * D + x * 64,
                                    row0 + hb * 64, kEvictLast);
                        if (kKV)
                            tma_load_2d(b2 + x * Cfg::kHalfBoxBytes, &tmap_do64, &b_full[st], h * D + x * 64, row0 + hb * 64,
                                        kEvictLast);
                        else
                            tma_load_2d(b2 + x * Cfg::kHalfBoxBytes, &tmap_qkv64, &b_full[st], p.v_col0 + h * D + x * 64,
                                        row0 + hb * 64, kEvictLast);
                    }
                    if (++st == ST) { st = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == 17) {
        if (blockIdx.x < num_items) {
            // ------------------------------------------------------------------------------------ MMA issuer
            // The whole warp runs the loop (warp-uniform control flow and operands: descriptors and TMEM addresses live
            // in uniform registers and every tcgen05.mma is ONE instruction); only the elected lane issues.  With the
            // loop inside `if (lane == 0)` each MMA cost ~13 instructions (R2UR + an ELECT / BRA.U.ANY loop per operand
            // set), and the single issuing thread, not the tensor pipe, paced the kernel (ncu r02: 3 300 clk per half
            // block, tensor pipe 24 % active).
            const bool leader = elect_one();
            const bool tr = AF_DBG(16) && blockIdx.x == 0 && leader;
            int tn = 0;
            constexpr uint32_t idesc_sc = make_idesc_bf16(128, 64, 0, 0);     // scores: [128 stationary rows] x [64 streamed rows]
            constexpr uint32_t idesc_ac = make_idesc_bf16(128, D, 0, 1);      // accumulators: A from TMEM, B MN-major
            const uint32_t a1_addr = smem_u32(smem_a1), a2_addr = smem_u32(smem_a2);
            // cursor of the NEXT score pair to issue (two half blocks ahead of the accumulating products, across items)
            int n_item = blockIdx.x, n_hb = 0, n_st = 0, n_g = 0;
            uint32_t n_aph = 0, n_bph = 0;
            auto issue_scores = [&]() {
                if (n_hb == 0) {
                    mbar_wait(a_full, n_aph);
                    n_aph ^= 1;
                }
                if (tr) af_trace(0, tn, 100 + (n_hb == 0 ? 50 : 0));      // before the waits of a score pair
                mbar_wait(&b_full[n_st], n_bph);
                tc_fence_after();
                if (tr) af_trace(0, tn, 101);                               // operands ready, issuing
                const uint32_t b1_addr = smem_u32(smem_b + n_st * Cfg::kStageBytes);
                const uint32_t b2_addr = b1_addr + Cfg::kStreamBytes;
                const uint32_t col = tmem_base + (n_g & 1) * 128;
                const bool last = n_hb == nhb - 1;
                if (leader) {
#pragma unroll
                    for (int tt = 0; tt < D / 16; ++tt) {
                        const uint32_t aoff = (tt >> 2) * kFaBoxBytes + (tt & 3) * 32;
                        const uint32_t boff = (tt >> 2) * Cfg::kHalfBoxBytes + (tt & 3) * 32;
                        umma_bf16_ss(col, make_smem_desc_sw128(a1_addr + aoff, 16, 1024),
                                     make_smem_desc_sw128(b1_addr + boff, 16, 1024), idesc_sc, tt != 0);
                    }
#pragma unroll
                    for (int tt = 0; tt < D / 16; ++tt) {
                        const uint32_t aoff = (tt >> 2) * kFaBoxBytes + (tt & 3) * 32;
                        const uint32_t boff = (tt >> 2) * Cfg::kHalfBoxBytes + (tt & 3) * 32;
                        umma_bf16_ss(col + 64, make_smem_desc_sw128(a2_addr + aoff, 16, 1024),
                                     make_smem_desc_sw128(b2_addr + boff, 16, 1024), idesc_sc, tt != 0);
                    }
                    umma_commit(&sd_full[n_g & 1]);
                    if (last) umma_commit(a_empty);   // the accumulating products read TMEM and the streamed tiles only
                }
                ++n_g;
                if (++n_st == ST) { n_st = 0; n_bph ^= 1; }
                if (++n_hb == nhb) {
                    n_hb = 0;
                    n_item += gridDim.x;
                }
            };
            issue_scores();
            if (n_item < num_items) issue_scores();
            int c_st = 0, g = 0;
            uint32_t eph = 0;
            bool first_item = true;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, first_item = false) {
                for (int hb = 0; hb < nhb; ++hb, ++g) {
                    const int s = g & 1;
                    if (tr) af_trace(0, tn, 200);                           // waiting for the stage's dS
                    mbar_wait(&ds_full[s], (g >> 1) & 1);
                    if (tr) af_trace(0, tn, 201);
                    if (hb == 0 && !first_item) {          // the previous item's accumulators have been read
                        mbar_wait(acc_empty, eph);
                        eph ^= 1;
                    }
                    tc_fence_after();
                    const uint32_t b1_addr = smem_u32(smem_b + c_st * Cfg::kStageBytes);
                    const uint32_t b2_addr = b1_addr + Cfg::kStreamBytes;
                    const uint32_t col = tmem_base + s * 128;
                    if (leader && AF_DBG(4)) {
                        // measurement only: free the stage and the streamed tiles without the accumulating products
                        umma_commit(&b_empty[c_st]);
                        if (hb == nhb - 1) umma_commit(acc_full);
                    } else if (leader) {
                        if (kKV) {
                            // dV_j += P^T . dO_ih ;  dK_j += dS^T . Q_ih   (bf16 P^T at cols [0,16) u [32,48) of the stage, dS^T
                            // at [64,80) u [96,112): each softmax thread writes over columns it has read itself)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16_ts(tmem_base + kColAcc0, col + (kk >> 1) * 32 + (kk & 1) * 8,
                                             make_smem_desc_sw128(b2_addr + kk * 2048, Cfg::kHalfBoxBytes, 1024), idesc_ac,
                                             (hb | kk) != 0);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16_ts(tmem_base + kColAcc1, col + 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                                             make_smem_desc_sw128(b1_addr + kk * 2048, Cfg::kHalfBoxBytes, 1024), idesc_ac,
                                             (hb | kk) != 0);
                        } else {
                            // dQ_i += dS . K_jh   (bf16 dS at cols [64,80) u [96,112) of the stage)
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                umma_bf16_ts(tmem_base + kColAcc0, col + 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                                             make_smem_desc_sw128(b1_addr + kk * 2048, Cfg::kHalfBoxBytes, 1024), idesc_ac,
                                             (hb | kk) != 0);
                        }
                        umma_commit(&b_empty[c_st]);
                        if (hb == nhb - 1) umma_commit(acc_full);
                    }
                    if (++c_st == ST) c_st = 0;
                    // next score pair: at once while it belongs to THIS item; the first two of the next item wait for
                    // that item's stationary tiles (a TMA round trip after this item's last score product) -- issuing
                    // them here would hold back this item's remaining accumulating products behind that wait (ncu r02:
                    // 13-15 % of all samples in the acc_full wait), so they are issued after the item's last product
                    if (n_item == item) issue_scores();
                }
                while (n_item < num_items && n_g < g + 2) issue_scores();   // pair n needs acc(n - 2) issued: 2 stages
            }
        }
    } else {
        // -------------------------------------------------------------------------------- softmax-backward warpgroups
        const int q = warp & 3;       // TMEM lane quadrant
        const int w = warp >> 3;      // group = parity of the CTA's half-block counter it serves
        const int hcol = (warp >> 2) & 1;   // which 32 of the 64 streamed columns this warp owns
        const int r = q * 32 + lane;  // stationary row inside the tile
        const int tw = threadIdx.x & 255;   // thread inside the group
        const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
        const uint32_t st_base = lane_base + w * 128 + hcol * 32;
        float* stats = smem_stats + w * (2 * 3 * 64);     // [buffer][lse | delta * scale | row seed][64]
        const float sc = p.scale_log2e;
        const uint32_t drop_site = kDrop ? drop_site_seed(p.drop) : 0u;
        const float drop_inv = kDrop ? drop_inv_keep(p.drop.thresh16) : 1.0f;
        const uint32_t thr = p.drop.thresh16;
        int g0 = 0, nproc = 0;        // half blocks of this CTA before the current item; half blocks this warpgroup has done
        int tns = 0;
        uint32_t fph = 0, aph = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, g0 += nhb) {
            const int rt = item % rtiles, bh = item / rtiles;
            const int h = bh % p.heads, b = bh / p.heads;
            const int srow = rt * 128 + r;                 // stationary row: query (dQ pass) or key (dK/dV pass)
            const bool row_ok = srow < S;
            const long long stat0 = (long long)bh * S;     // first entry of this (sequence, head) in lse / delta
            float lse_r = INFINITY, dls_r = 0.f;
            uint32_t rs_r = 0u;
            if (!kKV) {
                if (row_ok) {
                    lse_r = __ldg(p.lse + stat0 + srow);
                    dls_r = __ldg(p.delta + stat0 + srow) * p.scale;
                }
                if (kDrop) rs_r = drop_row_seed(drop_site, (uint32_t)(stat0 + srow));
            }
            // dK/dV pass: per-COLUMN statistics (lse, delta * scale of the 64 queries of a half block), loaded by threads
            // 0..63 of the warpgroup one half block AHEAD (the global-load latency sits off the S -> dS chain) and staged
            // through shared memory
            const int hb_first = ((g0 & 1) == w) ? 0 : 1;
            float pf_lse = INFINITY, pf_dls = 0.f;
            auto prefetch_cols = [&](int hb) {
                pf_lse = INFINITY;
                pf_dls = 0.f;
                if (kKV && tw < 64 && hb < nhb) {
                    const int qi = hb * 64 + tw;
                    if (qi < S) {                          // +inf -> P = 0 outside the sequence
                        pf_lse = __ldg(p.lse + stat0 + qi);
                        pf_dls = __ldg(p.delta + stat0 + qi);
                    }
                }
            };
            prefetch_cols(hb_first);
            for (int hb = hb_first; hb < nhb; hb += 2) {
                const int c0 = hb * 64;                    // first streamed row (key | query) of the half block
                float* sbuf = stats + (nproc & 1) * (3 * 64);
                if (kKV) {
                    if (tw < 64) {
                        sbuf[tw] = pf_lse;
                        sbuf[64 + tw] = pf_dls * p.scale;
                    } else if (kDrop && tw < 128) {
                        reinterpret_cast<uint32_t*>(sbuf)[128 + tw - 64] = drop_row_seed(drop_site, (uint32_t)(stat0 + c0 + tw - 64));
                    }
                    named_bar_sync(1 + w, 256);
                    prefetch_cols(hb + 2);
                }
                ++nproc;
                const bool trs = AF_DBG(16) && blockIdx.x == 0 && warp == 0 && lane == 0;
                if (trs) af_trace(1, tns, 300);
                mbar_wait(&sd_full[w], fph);
                fph ^= 1;
                tc_fence_after();
                if (trs) af_trace(1, tns, 301);
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t s[16], dp[16];
                    uint32_t pk[8], dk[8];
                    if (AF_DBG(2)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = dk[i] = 0u;
                        if (kKV) tmem_st_x8(st_base + c * 8, pk);
                        tmem_st_x8(st_base + 64 + c * 8, dk);
                        continue;
                    }
                    tmem_ld_x16(st_base + c * 16, s);
                    tmem_ld_x16(st_base + 64 + c * 16, dp);
                    tmem_ld_wait();
                    if (AF_DBG(1)) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            pk[i] = s[2 * i] ^ s[2 * i + 1];
                            dk[i] = dp[2 * i] ^ dp[2 * i + 1];
                        }
                        if (kKV) tmem_st_x8(st_base + c * 8, pk);
                        tmem_st_x8(st_base + 64 + c * 8, dk);
                        continue;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const int cc = hcol * 32 + c * 16 + i;   // column inside the half block
                        float p0, p1, d0, d1;
                        float dp0 = __uint_as_float(dp[i]), dp1 = __uint_as_float(dp[i + 1]);
                        if (!kKV) {
                            p0 = ex2_approx(fmaf(__uint_as_float(s[i]), sc, -lse_r));
                            p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), sc, -lse_r));
                            if (c0 + cc >= S) p0 = 0.f;    // keys beyond the sequence (rows of the next one / padding)
                            if (c0 + cc + 1 >= S) p1 = 0.f;
                            if (kDrop) {
                                const uint32_t bits = drop_pair_bits(rs_r, (uint32_t)(c0 + cc) >> 1);
                                dp0 = (bits & 0xffffu) >= thr ? dp0 * drop_inv : 0.f;
                                dp1 = (bits >> 16) >= thr ? dp1 * drop_inv : 0.f;
                            }
                            d0 = p0 * fmaf(dp0, p.scale, -dls_r);
                            d1 = p1 * fmaf(dp1, p.scale, -dls_r);
                        } else {
                            const float2 ls = *reinterpret_cast<const float2*>(sbuf + cc);
                            const float2 dl = *reinterpret_cast<const float2*>(sbuf + 64 + cc);
                            p0 = ex2_approx(fmaf(__uint_as_float(s[i]), sc, -ls.x));
                            p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), sc, -ls.y));
                            if (!row_ok) { p0 = 0.f; p1 = 0.f; }   // key rows beyond the sequence
                            float m0 = 1.f, m1 = 1.f;
                            if (kDrop) {
                                const uint2 rs = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(sbuf) + 128 + cc);
                                const uint32_t b0 = drop_pair_bits(rs.x, (uint32_t)srow >> 1);
                                const uint32_t b1 = drop_pair_bits(rs.y, (uint32_t)srow >> 1);
                                const uint32_t u0 = (srow & 1) ? (b0 >> 16) : (b0 & 0xffffu);
                                const uint32_t u1 = (srow & 1) ? (b1 >> 16) : (b1 & 0xffffu);
                                m0 = u0 >= thr ? drop_inv : 0.f;
                                m1 = u1 >= thr ? drop_inv : 0.f;
                            }
                            d0 = p0 * fmaf(dp0 * m0, p.scale, -dl.x);
                            d1 = p1 * fmaf(dp1 * m1, p.scale, -dl.y);
                            pk[i >> 1] = pack_bf16x2(p0 * m0, p1 * m1);
                        }
                        dk[i >> 1] = pack_bf16x2(d0, d1);
                    }
                    // bf16 tiles over the fp32 scores this thread has read: 16 columns [32 h + 16 c, +16) -> 8 packed
                    // columns at [32 h + 8 c, +8) of the S (P^T) and dP (dS) blocks
                    if (kKV) tmem_st_x8(st_base + c * 8, pk);
                    tmem_st_x8(st_base + 64 + c * 8, dk);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ds_full[w]);
                if (trs) af_trace(1, tns, 302);
            }
            // ---- epilogue: accumulators -> bf16 -> the packed gradient tensor
            if (AF_DBG(16) && blockIdx.x == 0 && warp == 0 && lane == 0) af_trace(1, tns, 400);
            mbar_wait(acc_full, aph);
            aph ^= 1;
            tc_fence_after();
            if (AF_DBG(16) && blockIdx.x == 0 && warp == 0 && lane == 0) af_trace(1, tns, 401);
            __nv_bfloat16* dst_row = p.dqkv + (long long)(b * S + srow) * p.ld + h * D;
            if (kKV) {
                // group 0: dV (accumulator 0), group 1: dK (accumulator 1); each warp of a quadrant pair stores D / 2 columns
                constexpr int HALF = D / 2;
                const uint32_t acc = lane_base + (w == 0 ? kColAcc0 : kColAcc1) + hcol * HALF;
                __nv_bfloat16* dst = dst_row + (w == 0 ? p.v_col0 : p.k_col0) + hcol * HALF;
#pragma unroll
                for (int c = 0; c < HALF / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld_x16(acc + c * 16, o);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 16; i += 8) {
                            uint4 u;
                            u.x = pack_bf16x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1]));
                            u.y = pack_bf16x2(__uint_as_float(o[i + 2]), __uint_as_float(o[i + 3]));
                            u.z = pack_bf16x2(__uint_as_float(o[i + 4]), __uint_as_float(o[i + 5]));
                            u.w = pack_bf16x2(__uint_as_float(o[i + 6]), __uint_as_float(o[i + 7]));
                            *reinterpret_cast<uint4*>(dst + c * 16 + i) = u;
                        }
                    }
                }
            } else {
                // dQ: the four warps of a quadrant (2 groups x 2 column halves) store D / 4 columns each
                constexpr int QUART = D / 4;
                const int part = 2 * w + hcol;
                const uint32_t acc = lane_base + kColAcc0 + part * QUART;
                __nv_bfloat16* dst = dst_row + p.q_col0 + part * QUART;
#pragma unroll
                for (int c = 0; c < QUART / 8; ++c) {
                    uint32_t o[8];
                    tmem_ld_x8(acc + c * 8, o);
                    tmem_ld_wait();
                    if (row_ok) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[0]), __uint_as_float(o[1]));
                        u.y = pack_bf16x2(__uint_as_float(o[2]), __uint_as_float(o[3]));
                        u.z = pack_bf16x2(__uint_as_float(o[4]), __uint_as_float(o[5]));
                        u.w = pack_bf16x2(__uint_as_float(o[6]), __uint_as_float(o[7]));
                        *reinterpret_cast<uint4*>(dst + c * 8) = u;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            if (AF_DBG(16) && blockIdx.x == 0 && warp == 0 && lane == 0) af_trace(1, tns, 402);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace fame

// K2: fused attention forward on tcgen05 / TMEM -- persistent, two query tiles per CTA.
//
//   ctx[b, q, h, :] = softmax_k( Q.K^T * scale  (+ -inf on masked / out-of-sequence keys) ) . V
//
// Replaces F.scaled_dot_product_attention as reached from BertSelfAttention (HF modeling_bert.py:192-206; additive
// key-padding mask of HF:709-713; 10_FAME.py:140) with head_dim 64, and nn.MultiheadAttention inside
// nn.TransformerEncoderLayer of BEHRTModel_Lab (10_FAME.py:212-215, 222; 8 heads => head_dim 96, 542 tokens).
//
// Why this shape.  At head_dim 64 a 128x128 score tile costs 16384 exponentials (MUFU: 16 / clk / SM = 1024 clk) but
// only 512 clk of tensor pipe for its two MMAs, so the kernel is bound by the softmax warps, not by tcgen05.  The
// design therefore (a) keeps TWO independent score tiles in flight per SM, one per softmax warpgroup, so that the
// MUFU pipe never waits for an MMA round trip, (b) takes everything that is not an exponential off the softmax
// warps: the output accumulator lives in TMEM (tcgen05.mma accumulates P.V across key blocks; it is rescaled in place
// only when a row maximum grows by more than 2^8 -- "lazy rescale"), max uses 3-input FMNMX3, scale/sum use packed
// FFMA2/FADD2, and (c) is persistent: one CTA per SM loops over (sequence, head, 256-query) work items while the TMA
// producer and the MMA issuer run ahead across item boundaries, so launch / prologue / epilogue latency is hidden.
//
//   warp 8 lane 0 : TMA producer    Q pair (2 x 128 rows) per item, K_j / V_j ring (128 keys per stage), SW128 boxes
//   warp 9 lane 0 : MMA issuer      S_t = Q_t . K_j^T (SS)  -> TMEM cols [t*128, t*128+128)        t = 0, 1
//                                   O_t (+)= P_t . V_j (TS) -> TMEM cols [256 + t*128, .. + D)
//   warps 0-3     : softmax warpgroup of tile 0 (thread = one query row, all 128 keys of the block)
//   warps 4-7     : softmax warpgroup of tile 1
// Ping-pong (p.pingpong): left alone, both warpgroups run their exponential phases at the SAME time (their score tiles
// are issued back to back), share the MUFU pipe, finish together and then both wait while the single MMA thread issues
// P.V and the next scores of both tiles: measured 2 475 clk per 128 x 128 tile at head_dim 64 against a MUFU floor of
// 1 024 (ncu r01c: XU 41 % active, a quarter of all samples in the s_full wait).  Two named barriers make the
// warpgroups take turns on the exponential loop, so that tile 1's exponentials run under tile 0's MMAs and vice versa.
// P_t (bf16) overwrites the first 64 columns of S_t; the in-order MMA pipe guarantees S_t(j+1) is written only after
// P_t(j) was consumed.  A commit on s_full[t] for block j also covers P.V of block j-1 (same issuing thread), so the
// softmax warps may rescale O_t right after that wait.
//
// Tried and measured slower (kept out): (0, r02) one MMA-issuing warp PER QUERY TILE (two independent instruction
// streams, K / V slots and the Q pair released by two arrivals): 0.48 instead of 0.38 ms at 256 x 12 x 512 x 64 with
// 64-key blocks, 2.48 instead of 2.39 ms at 1024 x 8 x 542 x 96 -- the ~65-130 clk a small tcgen05.mma costs is paid in
// the tensor pipe / operand fetch, not in the issuing thread; (1) a third, rotating score buffer with per-(tile, buffer) barriers so that a
// tile's next score block is in TMEM before its warpgroup finishes the current one (0.50 ms vs 0.41 ms at 256 x 12 x
// 512 x 64: the extra commits and cursor arithmetic in the single MMA-issuing thread cost more than the wait saved);
// (2) fetching the key-mask bytes one block ahead (0.45 ms: the extra live registers spill in the softmax loop).
#pragma once
#include "sm100_ptx.cuh"
#include "attn_common.cuh"

namespace fame {

constexpr int kApThreads = 320;
constexpr float kApLazyLog2 = 8.0f;     // rescale the accumulator only when a row max grows by more than 2^8

// KB = keys per block.  128: one CTA per SM (512 TMEM columns).  64 (head_dim 64 only): a CTA needs 2 x 64 score + 2 x 64
// output columns = 256 and ~100 registers per thread, so TWO CTAs share an SM -- four query tiles and sixteen softmax
// warps in flight instead of two and eight.  With eight warps (two per scheduler) the S -> softmax -> P.V -> next S chain
// of a tile (~2 500 clk per key block against 1 024 clk of MUFU work) is not hidden; the second CTA fills the gaps.
template <int D, int KB = 128>
struct ApCfg {
    static constexpr int kBoxes = (D + 63) / 64;
    static constexpr int kTileBytes = kBoxes * kFaBoxBytes;       // Q tile: 128 rows x (64 | 128) bf16 columns
    static constexpr int kKvBoxBytes = KB * 64 * 2;               // one K / V box: KB rows x 64 bf16 columns
    static constexpr int kKvBytes = kBoxes * kKvBoxBytes;
    static constexpr int kQBufs = (D <= 64 && KB == 128) ? 2 : 1; // Q pairs in flight
    static constexpr int kStages = (D <= 64) ? 3 : 2;             // K/V ring depth
    static constexpr int kCtasPerSm = KB == 64 ? 2 : 1;
    static constexpr int kTmemCols = KB == 64 ? 256 : 512;
    static constexpr int kColO = 2 * KB;                          // O_t at kColO + t * kOStride
    static constexpr int kOStride = KB == 64 ? 64 : 128;
    static constexpr int kSmemBytes = kTileBytes * 2 * kQBufs + kKvBytes * 2 * kStages + 1024 /*barriers*/ + 1024 /*align*/;
    static_assert(KB == 128 || (KB == 64 && D == 64), "64-key blocks are built for head_dim 64");
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// kDrop: dropout of the attention probabilities (training; torch's F.multi_head_attention_forward applies
// dropout_p to softmax(QK^T) before the product with V, and HF's sdpa path passes dropout_p the same way, HF:205).
// The row sum l (softmax denominator) and the saved log-sum-exp use the UNDROPPED exponentials; dropped entries of P
// are zeroed before P.V and the 1 / (1 - p) scale is folded into the final 1 / l.  Mask row = (sequence, head, query),
// mask column = key (dropout.cuh); attn_bwd_pds_kernel<D, true> regenerates the same bits.
template <int D, bool kDrop = false, int KB = 128>
__global__ void __launch_bounds__(kApThreads, (KB == 64 ? 2 : 1))
attn_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_kv,
                     const FaParams p, const int num_items, const int qpairs) {
    using Cfg = ApCfg<D, KB>;
    constexpr int KVT = Cfg::kKvBytes;          // one K or V stage
    constexpr int KVBOX = Cfg::kKvBoxBytes;
    constexpr int NC = KB / 32;                 // 32-column chunks of a score block
    constexpr int NB = Cfg::kBoxes;
    constexpr int ST = Cfg::kStages;
    constexpr int QB = Cfg::kQBufs;
    constexpr int TILE = Cfg::kTileBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;                          // [QB][2][TILE]
    uint8_t* smem_k = smem_q + QB * 2 * TILE;        // [ST][KVT]
    uint8_t* smem_v = smem_k + ST * KVT;             // [ST][KVT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + ST * KVT);
    uint64_t* q_full = bars;                 // [QB]
    uint64_t* q_empty = q_full + QB;         // [QB]
    uint64_t* k_full = q_empty + QB;         // [ST]
    uint64_t* v_full = k_full + ST;          // [ST]
    uint64_t* kv_empty = v_full + ST;        // [ST]
    uint64_t* s_full = kv_empty + ST;        // [2]  S_t written (and P.V of the previous block of tile t complete)
    uint64_t* p_full = s_full + 2;           // [2]  P_t in TMEM, O_t rescaled (4 warp arrivals)
    uint64_t* o_full = p_full + 2;           // [2]  last P.V of the item complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.seq;
    const int nblk = (S + KB - 1) / KB;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < QB; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
        }
        for (int i = 0; i < ST; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColO = Cfg::kColO;
    constexpr uint32_t OST = Cfg::kOStride;

    // item -> (sequence b, head h, query pair qp); consecutive items share K/V of one (b, h) through L2
    auto t1_active = [&](int item) { return (item % qpairs) * 256 + 128 < S; };
    // key blocks of an item: all of them, or only those that hold an attended key of its sequence (p.kv_len); at
    // least one, so that an all-masked sequence still produces its zero rows
    auto nblk_of = [&](int item) {
        if (p.kv_len == nullptr) return nblk;
        const int len = abs(__ldg(p.kv_len + (item / qpairs) / p.heads));   // sign: prefix mask or not (mask_kv_len_kernel)
        return max(1, min(nblk, (len + KB - 1) / KB));
    };

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------------------------------------------ TMA producer
            int qb = 0, st = 0;
            uint32_t qph = 0, kph = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int qp = item % qpairs, bh = item / qpairs;
                const int h = bh % p.heads, b = bh / p.heads;
                const int row0 = b * S;
                const int nt = t1_active(item) ? 2 : 1;
                mbar_wait(&q_empty[qb], qph ^ 1);
                mbar_arrive_expect_tx(&q_full[qb], TILE * nt);
                for (int t = 0; t < nt; ++t)
#pragma unroll
                    for (int x = 0; x < NB; ++x)
                        tma_load_2d(smem_q + (qb * 2 + t) * TILE + x * kFaBoxBytes, &tmap_qkv, &q_full[qb],
                                    p.q_col0 + h * D + x * 64, row0 + qp * 256 + t * 128, kEvictFirst);
                if (++qb == QB) { qb = 0; qph ^= 1; }
                const int nb_item = nblk_of(item);
                for (int j = 0; j < nb_item; ++j) {
                    mbar_wait(&kv_empty[st], kph ^ 1);
                    mbar_arrive_expect_tx(&k_full[st], KVT);
#pragma unroll
                    for (int x = 0; x < NB; ++x)
                        tma_load_2d(smem_k + st * KVT + x * KVBOX, &tmap_kv, &k_full[st],
                                    p.k_col0 + h * D + x * 64, row0 + j * KB, kEvictLast);
                    mbar_arrive_expect_tx(&v_full[st], KVT);
#pragma unroll
                    for (int x = 0; x < NB; ++x)
                        tma_load_2d(smem_v + st * KVT + x * KVBOX, &tmap_kv, &v_full[st],
                                    p.v_col0 + h * D + x * 64, row0 + j * KB, kEvictLast);
                    if (++st == ST) { st = 0; kph ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        if (blockIdx.x < num_items) {
            // ------------------------------------------------------------------------------------ MMA issuer
            // The whole warp runs the loop (warp-uniform control flow: descriptors and TMEM addresses stay in uniform
            // registers, every tcgen05.mma is one instruction); only the elected lane issues.  Inside `if (lane == 0)`
            // each MMA cost ~13 instructions (R2UR + an ELECT / BRA.U.ANY loop), which put ~900 clk of issue time per
            // tile and key block on the softmax -> MMA -> softmax chain.
            const bool leader = elect_one();
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, KB, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
            // cursor over the NEXT score block to issue (runs one block ahead of the P.V products, across items)
            int n_item = blockIdx.x, n_j = 0, n_qb = 0, ks = 0;
            uint32_t n_qph = 0, kph = 0;
            bool n_t1 = t1_active(n_item);
            int n_nblk = nblk_of(n_item);
            auto issue_s = [&](int t) {
                if (t == 0) {
                    if (n_j == 0) mbar_wait(&q_full[n_qb], n_qph);
                    mbar_wait(&k_full[ks], kph);
                }
                tc_fence_after();
                const uint32_t q_addr = smem_u32(smem_q + (n_qb * 2 + t) * TILE);
                const uint32_t k_addr = smem_u32(smem_k + ks * KVT);
                if (leader) {
#pragma unroll
                    for (int tt = 0; tt < D / 16; ++tt) {
                        const uint32_t off = (tt >> 2) * kFaBoxBytes + (tt & 3) * 32;
                        const uint32_t koff = (tt >> 2) * KVBOX + (tt & 3) * 32;
                        umma_bf16_ss(tmem_base + t * KB, make_smem_desc_sw128(q_addr + off, 16, 1024),
                                     make_smem_desc_sw128(k_addr + koff, 16, 1024), idesc_qk, tt != 0);
                    }
                    umma_commit(&s_full[t]);
                    if ((t == 1 || !n_t1) && n_j == n_nblk - 1) umma_commit(&q_empty[n_qb]);
                }
                if (t == 1 || !n_t1) {          // last tile of this block: advance the cursor
                    if (++ks == ST) { ks = 0; kph ^= 1; }
                    if (++n_j == n_nblk) {
                        n_j = 0;
                        n_item += gridDim.x;
                        if (++n_qb == QB) { n_qb = 0; n_qph ^= 1; }
                        n_t1 = n_item < num_items && t1_active(n_item);
                        if (n_item < num_items) n_nblk = nblk_of(n_item);
                    }
                }
            };
            {
                const bool c_t1 = n_t1;
                issue_s(0);
                if (c_t1) issue_s(1);
            }
            int vs = 0;
            uint32_t vph = 0, pph0 = 0, pph1 = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const bool t1 = t1_active(item);
                const int nb_item = nblk_of(item);
                for (int j = 0; j < nb_item; ++j) {
                    mbar_wait(&v_full[vs], vph);
                    const uint32_t v_addr = smem_u32(smem_v + vs * KVT);
                    auto issue_pv = [&](int t, bool free_kv) {
                        if (leader) {
#pragma unroll
                            for (int kk = 0; kk < KB / 16; ++kk)
                                umma_bf16_ts(tmem_base + kColO + t * OST, tmem_base + t * KB + kk * 8,
                                             make_smem_desc_sw128(v_addr + kk * 2048, KVBOX, 1024), idesc_pv,
                                             (j | kk) != 0);
                            if (j == nb_item - 1) umma_commit(&o_full[t]);
                            if (free_kv) umma_commit(&kv_empty[vs]);
                        }
                    };
                    mbar_wait(&p_full[0], pph0);
                    pph0 ^= 1;
                    tc_fence_after();
                    issue_pv(0, !t1);
                    const bool nx = n_item < num_items;
                    const bool c_t1 = n_t1;
                    if (nx) issue_s(0);
                    if (t1) {
                        mbar_wait(&p_full[1], pph1);
                        pph1 ^= 1;
                        tc_fence_after();
                        issue_pv(1, true);
                    }
                    if (nx && c_t1) issue_s(1);
                    if (++vs == ST) { vs = 0; vph ^= 1; }
                }
            }
        }
    } else {
        // -------------------------------------------------------------------------------- softmax warpgroups
        const int q = warp & 3;       // TMEM lane quadrant of this warp
        const int t = warp >> 2;      // query tile (= warpgroup)
        const int r = q * 32 + lane;  // row inside the tile
        const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
        const uint32_t s_addr = lane_base + t * KB;
        const uint32_t o_addr = lane_base + kColO + t * OST;
        const float sc = p.scale_log2e;
        uint32_t sph = 0, oph = 0;
        const uint32_t drop_site = kDrop ? drop_site_seed(p.drop) : 0u;
        const bool pp = p.pingpong != 0;
        if (pp && t == 1) named_bar_arrive(1, 256);          // warpgroup 0 owns the first turn
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            if (t == 1 && !t1_active(item)) {
                if (pp)                                      // keep the turn counts of the two warpgroups equal
                    for (int j = nblk_of(item); j > 0; --j) {
                        named_bar_sync(2, 256);
                        named_bar_arrive(1, 256);
                    }
                continue;
            }
            const int qp = item % qpairs, bh = item / qpairs;
            const int h = bh % p.heads, b = bh / p.heads;
            const uint32_t drop_rs = kDrop ? drop_row_seed(drop_site, (uint32_t)(bh * S + qp * 256 + t * 128 + r)) : 0u;
            float m = -INFINITY, l = 0.f;   // running reference max (log2 units, scaled) and row sum
            const int nb_item = nblk_of(item);
            // keys [0, valid_to) are attended when the mask is absent or a prefix mask (kv_len >= 0): validity then needs
            // no memory access on the S -> softmax -> P chain; a mask with holes (kv_len < 0, or no kv_len) is read
            int valid_to = S;
            bool by_len = p.key_mask == nullptr;
            if (p.key_mask != nullptr && p.kv_len != nullptr) {
                const int len = __ldg(p.kv_len + b);
                if (len >= 0) {
                    by_len = true;
                    valid_to = min(S, len);
                }
            }
            for (int j = 0; j < nb_item; ++j) {
                // validity of the 128 keys of this block: inside the sequence and not masked by the caller
                const int key0 = j * KB;
                uint32_t vm[NC];
                bool all_valid = by_len && key0 + KB <= valid_to;
                if (!all_valid) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const int k = key0 + c * 32 + lane;
                        const bool ok = by_len ? k < valid_to : (k < S && p.key_mask[(long long)b * S + k] != 0);
                        vm[c] = __ballot_sync(0xffffffffu, ok);
                    }
                    uint32_t va = vm[0];
#pragma unroll
                    for (int c = 1; c < NC; ++c) va &= vm[c];
                    all_valid = va == 0xffffffffu;
                }
                mbar_wait(&s_full[t], sph);
                sph ^= 1;
                tc_fence_after();
                uint32_t s[NC][32];
#pragma unroll
                for (int c = 0; c < NC; ++c) tmem_ld_x32(s_addr + c * 32, s[c]);
                tmem_ld_wait();
                if (!all_valid) {
#pragma unroll
                    for (int c = 0; c < NC; ++c)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (!((vm[c] >> i) & 1u)) s[c][i] = 0xff800000u;   // -inf
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        mx0 = fmax3(mx0, __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
                        mx1 = fmax3(mx1, __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]));
                    }
                // scores are scaled by a positive constant, so the max commutes with the scaling
                const float m_blk = fmaxf(mx0, mx1) * sc;
                float m_new = fmaxf(m, m_blk);
                if (m != -INFINITY && m_new - m <= kApLazyLog2) m_new = m;   // lazy: keep the old reference max
                const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
                const float alpha = (m_new == m) ? 1.0f : ex2_approx(m - m_safe);   // m == -inf -> 0
                const unsigned long long sc2 = f32x2_pack(sc, sc), nm2 = f32x2_pack(-m_safe, -m_safe);
                unsigned long long sum2 = 0ull;   // (0.f, 0.f)
                if (pp) named_bar_sync(1 + t, 256);          // my turn on the MUFU pipe
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const unsigned long long x2 =
                            f32x2_fma(f32x2_pack(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
                        float x0, x1;
                        f32x2_unpack(x2, x0, x1);
                        float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                        sum2 = f32x2_add(sum2, f32x2_pack(p0, p1));
                        if (kDrop) {
                            const uint32_t bits = drop_pair_bits(drop_rs, (uint32_t)(key0 + c * 32 + i) >> 1);
                            p0 = (bits & 0xffffu) >= p.drop.thresh16 ? p0 : 0.f;
                            p1 = (bits >> 16) >= p.drop.thresh16 ? p1 : 0.f;
                        }
                        pk[i >> 1] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_x16(s_addr + c * 16, pk);
                }
                if (pp) named_bar_arrive(2 - t, 256);        // hand the turn to the other warpgroup
                if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
                    // P.V of block j-1 is complete (covered by the s_full commit): rescale O_t in place.  Done after
                    // the exponentials so that the 128 score registers are dead by now (register pressure).
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld_x32(o_addr + c * 32, o);
                        tmem_ld_wait();
                        uint32_t lo[16], hi[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            lo[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            hi[i] = __float_as_uint(__uint_as_float(o[16 + i]) * alpha);
                        }
                        tmem_st_x16(o_addr + c * 32, lo);
                        tmem_st_x16(o_addr + c * 32 + 16, hi);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
                float sa, sb;
                f32x2_unpack(sum2, sa, sb);
                l = l * alpha + (sa + sb);
                m = m_new;
            }
            // ---- epilogue: O_t / l -> bf16 -> global
            mbar_wait(&o_full[t], oph);
            oph ^= 1;
            tc_fence_after();
            const float inv = (l > 0.f ? 1.0f / l : 0.f) * (kDrop ? drop_inv_keep(p.drop.thresh16) : 1.0f);
            const int qrow = qp * 256 + t * 128 + r;
            // saved for the backward pass: P = 2^(s * scale_log2e - lse); +inf for an all-masked row (P = 0)
            if (p.lse != nullptr && qrow < S)
                p.lse[((long long)b * p.heads + h) * S + qrow] = l > 0.f ? m + log2f(l) : INFINITY;
            __nv_bfloat16* dst = p.ctx + (long long)(b * S + qrow) * p.ld_ctx + h * D;
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld_x32(o_addr + c * 32, o);
                tmem_ld_wait();
                if (qrow < S) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + c * 32 + i) = u;
                    }
                }
            }
            // the reads of O_t above are ordered before this warp's next p_full arrival (tcgen05.wait::ld + the
            // fence before that arrive), which is what the next item's first P.V (accumulate = 0) waits for
        }
        if (pp && t == 0) named_bar_sync(1, 256);            // consume warpgroup 1's last hand-over
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace fame

// K2 (general): flash attention forward on tcgen05 / TMEM for any sequence length, head_dim 64 or 96.
//
// Used for the lab (BEHRT) encoder -- nn.MultiheadAttention(768, 8 heads => head_dim 96) over L = 542 feature
// tokens (10_FAME.py:214, 222) -- and for BERT heads (head_dim 64) of any length.
//
// One CTA per (128-query tile, head, sequence); keys are streamed in blocks of 128 through a TMA ring.
//   control warps : warp 8 lane 0 = TMA producer, warp 9 lane 0 = MMA issuer
//   MMA1 (SS)     : S_j[128,128] = Q . K_j^T            -> TMEM S buffer j%2 (double buffered, 2 x 128 columns)
//   softmax       : two warpgroups split each key block: WG0 keys [0,64), WG1 keys [64,128).  Each keeps its OWN
//                   online-softmax state (running max m, sum l, fp32 output accumulator in registers), so no
//                   cross-warpgroup exchange is needed per block; the two states are merged once at the end
//                   (split-key merge).  P (bf16) is written back to TMEM in place of S.
//   MMA2 (TS)     : O_x,j[128,D] = P_x,j[tmem] . V_j[64 keys of WG x, MN-major smem]  -> TMEM O_x (fresh each block);
//                   the warpgroup folds it into its register accumulator: acc = acc * alpha_j + O_x,j
// head_dim 96 rows (192 B) do not fit one 128-byte swizzle atom: every tile is loaded as two 64-column boxes
// (the second box is only half used; its extra columns are never addressed by an MMA).
#pragma once
#include "dropout.cuh"
#include "sm100_ptx.cuh"

namespace fame {

constexpr int kFaThreads = 320;
constexpr int kFaBoxBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 bf16 columns, SW128

template <int D>
struct FaCfg {
    static constexpr int kBoxes = (D + 63) / 64;
    static constexpr int kTileBytes = kBoxes * kFaBoxBytes;
    static constexpr int kStages = (D <= 64) ? 3 : 2;
    static constexpr int kSmemBytes = kTileBytes * (1 + 2 * kStages) + 1024 /*barriers*/ + 1024 /*align*/;
};

struct FaParams {
    const uint8_t* key_mask;  // [batch, seq] (1 = attend) or nullptr
    __nv_bfloat16* ctx;
    long long ld_ctx;
    int batch, seq, heads;
    int q_col0, k_col0, v_col0;  // first column of Q / K / V of head 0 inside the packed tensor
    float scale_log2e;
    float* lse;               // optional [batch, heads, seq]: row log-sum-exp in log2 units of the scaled scores
    const int* kv_len;        // optional [batch]: 1 + last attended key (pair kernel: skips fully masked key blocks)
    DropCfg drop;             // dropout of the attention probabilities (pair kernel, kDrop instantiation only)
};

template <int D>
__global__ void __launch_bounds__(kFaThreads, 1)
attn_fwd_flash_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const FaParams p) {
    using Cfg = FaCfg<D>;
    constexpr int NB = Cfg::kBoxes;
    constexpr int ST = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;
    uint8_t* smem_k = smem + Cfg::kTileBytes;
    uint8_t* smem_v = smem_k + ST * Cfg::kTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + ST * Cfg::kTileBytes);
    uint64_t* q_full = bars;                 // 1
    uint64_t* k_full = bars + 1;             // [ST]
    uint64_t* k_empty = k_full + ST;         // [ST]
    uint64_t* v_full = k_empty + ST;         // [ST]
    uint64_t* v_empty = v_full + ST;         // [ST]
    uint64_t* s_full = v_empty + ST;         // [2]   S buffer written by MMA1
    uint64_t* p_full = s_full + 2;           // [2]   per warpgroup: P written to TMEM (4 warp arrivals)
    uint64_t* o_full = p_full + 2;           // [2]   per warpgroup: MMA2 complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int S = p.seq;
    const int nblk = (S + 127) >> 7;
    const int row0 = b * S;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(q_full, 1);
        for (int i = 0; i < ST; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kColO = 256;  // O_A at 256, O_B at 384

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            mbar_arrive_expect_tx(q_full, Cfg::kTileBytes);
#pragma unroll
            for (int x = 0; x < NB; ++x)
                tma_load_2d(smem_q + x * kFaBoxBytes, &tmap_qkv, q_full, p.q_col0 + h * D + x * 64,
                            row0 + qt * 128, kEvictFirst);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&k_full[st], Cfg::kTileBytes);
#pragma unroll
                for (int x = 0; x < NB; ++x)
                    tma_load_2d(smem_k + st * Cfg::kTileBytes + x * kFaBoxBytes, &tmap_qkv, &k_full[st],
                                p.k_col0 + h * D + x * 64, row0 + j * 128, kEvictLast);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&v_full[st], Cfg::kTileBytes);
#pragma unroll
                for (int x = 0; x < NB; ++x)
                    tma_load_2d(smem_v + st * Cfg::kTileBytes + x * kFaBoxBytes, &tmap_qkv, &v_full[st],
                                p.v_col0 + h * D + x * 64, row0 + j * 128, kEvictLast);
                if (++st == ST) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
            const uint32_t q_addr = smem_u32(smem_q);
            auto issue_s = [&](int j) {
                const int st = j % ST;
                mbar_wait(&k_full[st], (j / ST) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(smem_k + st * Cfg::kTileBytes);
#pragma unroll
                for (int t = 0; t < D / 16; ++t) {
                    const uint32_t off = (t >> 2) * kFaBoxBytes + (t & 3) * 32;
                    umma_bf16_ss(tmem_base + (j & 1) * 128, make_smem_desc_sw128(q_addr + off, 16, 1024),
                                 make_smem_desc_sw128(k_addr + off, 16, 1024), idesc_qk, t != 0);
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[j & 1]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (nblk > 1) issue_s(1);
            for (int j = 0; j < nblk; ++j) {
                const int st = j % ST;
                mbar_wait(&v_full[st], (j / ST) & 1);
                const uint32_t v_addr = smem_u32(smem_v + st * Cfg::kTileBytes);
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    mbar_wait(&p_full[x], j & 1);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int key = x * 64 + kk * 16;
                        umma_bf16_ts(tmem_base + kColO + x * 128, tmem_base + (j & 1) * 128 + x * 64 + kk * 8,
                                     make_smem_desc_sw128(v_addr + key * 128, kFaBoxBytes, 1024), idesc_pv, kk != 0);
                    }
                    umma_commit(&o_full[x]);
                }
                umma_commit(&v_empty[st]);
                if (j + 2 < nblk) issue_s(j + 2);
            }
        }
    } else {
        // ---------------------------------------------------- softmax warpgroups (warps 0-3: WG0, 4-7: WG1)
        const int q = warp & 3;
        const int x = warp >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
        float acc[D];
#pragma unroll
        for (int i = 0; i < D; ++i) acc[i] = 0.f;
        float m = -INFINITY, l = 0.f, alpha_prev = 1.f;

        auto fold_o = [&](int j) {
            mbar_wait(&o_full[x], j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t oo[32];
                tmem_ld_x32(lane_base + kColO + x * 128 + c * 32, oo);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], alpha_prev, __uint_as_float(oo[i]));
            }
        };

        for (int j = 0; j < nblk; ++j) {
            const int key0 = j * 128 + x * 64;
            // validity of my 64 keys: inside the sequence and not masked by the caller
            unsigned long long valid = 0ull;
            if (key0 + 64 <= S && p.key_mask == nullptr) {
                valid = ~0ull;
            } else {
                // warp-cooperative: lane i tests keys key0 + i and key0 + 32 + i, two ballots give the 64-bit mask
                const int ka = key0 + lane, kb = ka + 32;
                const bool oka = ka < S && (p.key_mask == nullptr || p.key_mask[(long long)b * S + ka] != 0);
                const bool okb = kb < S && (p.key_mask == nullptr || p.key_mask[(long long)b * S + kb] != 0);
                const unsigned lo = __ballot_sync(0xffffffffu, oka), hi = __ballot_sync(0xffffffffu, okb);
                valid = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            uint32_t s0[32], s1[32];
            const uint32_t s_addr = lane_base + (j & 1) * 128 + x * 64;
            tmem_ld_x32(s_addr, s0);
            tmem_ld_x32(s_addr + 32, s1);
            tmem_ld_wait();
            float mb = -INFINITY;
            if (valid == ~0ull) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    mb = fmaxf(mb, __uint_as_float(s0[i]));
                    mb = fmaxf(mb, __uint_as_float(s1[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!((valid >> i) & 1)) s0[i] = 0xff800000u;          // -inf
                    if (!((valid >> (32 + i)) & 1)) s1[i] = 0xff800000u;
                    mb = fmaxf(mb, __uint_as_float(s0[i]));
                    mb = fmaxf(mb, __uint_as_float(s1[i]));
                }
            }
            // scores are scaled by a positive constant, so the max commutes with the scaling
            const float m_new = fmaxf(m, mb * p.scale_log2e);
            const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = ex2_approx(m - m_safe);  // m == -inf -> 0
            float sum = 0.f;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(s0[i]), p.scale_log2e, -m_safe));
                const float p1 = ex2_approx(fmaf(__uint_as_float(s0[i + 1]), p.scale_log2e, -m_safe));
                sum += p0 + p1;
                pk[i >> 1] = pack_bf16x2(p0, p1);
            }
            tmem_st_x16(s_addr, pk);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(s1[i]), p.scale_log2e, -m_safe));
                const float p1 = ex2_approx(fmaf(__uint_as_float(s1[i + 1]), p.scale_log2e, -m_safe));
                sum += p0 + p1;
                pk[i >> 1] = pack_bf16x2(p0, p1);
            }
            tmem_st_x16(s_addr + 16, pk);
            tmem_st_wait();
            // O of the previous block must be folded in before MMA2 of this block may overwrite it
            if (j > 0) fold_o(j - 1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[x]);
            l = l * alpha + sum;
            m = m_new;
            alpha_prev = alpha;
        }
        fold_o(nblk - 1);

        // ---- merge the two warpgroups' states (split-key merge) through smem (the K/V ring is idle now)
        float* mrg = reinterpret_cast<float*>(smem_k);  // [(D + 2)][128]
        if (x == 1) {
#pragma unroll
            for (int i = 0; i < D; ++i) mrg[i * 128 + r] = acc[i];
            mrg[D * 128 + r] = m;
            mrg[(D + 1) * 128 + r] = l;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (x == 0) {
            const float m1 = mrg[D * 128 + r], l1 = mrg[(D + 1) * 128 + r];
            const float mt = fmaxf(m, m1);
            const float ms = (mt == -INFINITY) ? 0.f : mt;
            const float a0 = ex2_approx(m - ms), a1 = ex2_approx(m1 - ms);
            const float den = l * a0 + l1 * a1;
            const float inv = den > 0.f ? 1.0f / den : 0.f;
            const int qrow = qt * 128 + r;
            if (qrow < S) {
                __nv_bfloat16* dst = p.ctx + (long long)(row0 + qrow) * p.ld_ctx + h * D;
#pragma unroll
                for (int i = 0; i < D; i += 8) {
                    float o[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) o[t] = (acc[i + t] * a0 + mrg[(i + t) * 128 + r] * a1) * inv;
                    uint4 u;
                    u.x = pack_bf16x2(o[0], o[1]);
                    u.y = pack_bf16x2(o[2], o[3]);
                    u.z = pack_bf16x2(o[4], o[5]);
                    u.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(dst + i) = u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace fame

// K2b: attention backward, first half -- probabilities and score gradients on tcgen05 / TMEM.
//
// autograd of F.scaled_dot_product_attention inside nn.MultiheadAttention of BEHRTModel_Lab (10_FAME.py:212-215,
// reached by total_loss.backward() at 10_FAME.py:445) needs, per (sequence, head):
//     P  = softmax(Q K^T * scale)                       dV = P^T dO
//     dP = dO V^T                                        dS = scale * P * (dP - rowsum(dO * O))
//     dQ = dS K,  dK = dS^T Q
// The unfused path materialised S and dP in fp32 (two [B, H, L, L] round trips through HBM: 2.4 GB per layer at 32
// patients, 77 GB at 1024) before a row-wise softmax-backward kernel.  Here both score products stay in TMEM: one CTA
// walks (sequence, head, 128-query tile) items, for each 128-key block the MMA warp issues S = Q K_j^T and
// dP = dO V_j^T into one of two TMEM buffer pairs, and the two softmax warpgroups alternate over key blocks,
// recomputing P from the forward's row log-sum-exp (no max / sum pass) and writing P and dS once, in bf16, for the
// three tensor-core products that follow (fame_gemm_ex: dV, dK, dQ).
//
//   warp 8 lane 0 : TMA producer    Q_i, dO_i per item; K_j, V_j ring (2 stages); SW128 boxes
//   warp 9 lane 0 : MMA issuer      S -> TMEM cols [256 b, +128), dP -> [256 b + 128, +128), b = key-block parity
//   warps 0-3     : warpgroup of even key blocks (thread = one query row)
//   warps 4-7     : warpgroup of odd key blocks
#pragma once
#include "dropout.cuh"
#include "sm100_ptx.cuh"
#include "attn_common.cuh"
#include "rowwise.cuh"      // bf16x8_to_float

namespace fame {

constexpr int kAbThreads = 320;

template <int D>
struct AbCfg {
    static constexpr int kBoxes = (D + 63) / 64;
    static constexpr int kTileBytes = kBoxes * kFaBoxBytes;
    static constexpr int kStages = 2;
    static constexpr int kStageOutBytes = 8 * 4096;   // per softmax warp: 32 rows x 64 B of P and of dS
    static constexpr int kSmemBytes = kTileBytes * (2 + 2 * kStages) + kStageOutBytes + 1024 /*barriers*/ + 1024 /*align*/;
};

struct AbParams {
    const float* lse;          // [batch, heads, seq]  row log-sum-exp of the forward, log2 units of the scaled scores
    const float* delta;        // [batch, heads, seq]  rowsum(dO * O)
    __nv_bfloat16* p;          // [batch, heads, seq, ldp]
    __nv_bfloat16* ds;         // [batch, heads, seq, ldp]
    long long ldp;
    int batch, seq, heads;
    int q_col0, k_col0, v_col0;   // first column of Q / K / V of head 0 inside the packed qkv tensor
    float scale, scale_log2e;
    DropCfg drop;                 // dropout of the attention probabilities in the forward (kDrop instantiation only)
};

// kDrop: the forward dropped entries of P (mask m, scale c = 1 / (1 - p)):  O = (m c P) V.  Then
//   dV = (m c P)^T dO   -> the P written here is m c P;     dP = m c (dO V^T);
//   dS = scale P (dP - delta),  delta = rowsum(dO * O) = sum_k P_k dP_k  still holds.
template <int D, bool kDrop = false>
__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_pds_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                    const AbParams p, const int num_items, const int qtiles) {
    using Cfg = AbCfg<D>;
    constexpr int NB = Cfg::kBoxes;
    constexpr int ST = Cfg::kStages;
    constexpr int TILE = Cfg::kTileBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;                    // [TILE]
    uint8_t* smem_do = smem_q + TILE;          // [TILE]
    uint8_t* smem_k = smem_do + TILE;          // [ST][TILE]
    uint8_t* smem_v = smem_k + ST * TILE;      // [ST][TILE]
    uint8_t* smem_out = smem_v + ST * TILE;    // [8 warps][2][32 rows][64 B] output staging
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + Cfg::kStageOutBytes);
    uint64_t* q_full = bars;                   // Q_i and dO_i landed
    uint64_t* q_empty = bars + 1;              // last MMAs of the item done
    uint64_t* kv_full = bars + 2;              // [ST]
    uint64_t* kv_empty = kv_full + ST;         // [ST]
    uint64_t* sd_full = kv_empty + ST;         // [2] S and dP of a key block complete in TMEM buffer pair b
    uint64_t* sd_free = sd_full + 2;           // [2] buffer pair b read by its warpgroup (4 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sd_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.seq;
    const int nblk = (S + 127) >> 7;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv);
        tma_prefetch_desc(&tmap_do);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < ST; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sd_full[i], 1);
            mbar_init(&sd_free[i], 4);
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

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------------------------------------------ TMA producer
            int st = 0;
            uint32_t qph = 0, kph = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int qt = item % qtiles, bh = item / qtiles;
                const int h = bh % p.heads, b = bh / p.heads;
                const int row0 = b * S;
                mbar_wait(q_empty, qph ^ 1);
                qph ^= 1;
                mbar_arrive_expect_tx(q_full, 2 * TILE);
#pragma unroll
                for (int x = 0; x < NB; ++x) {
                    tma_load_2d(smem_q + x * kFaBoxBytes, &tmap_qkv, q_full, p.q_col0 + h * D + x * 64, row0 + qt * 128,
                                kEvictFirst);
                    tma_load_2d(smem_do + x * kFaBoxBytes, &tmap_do, q_full, h * D + x * 64, row0 + qt * 128, kEvictFirst);
                }
                for (int j = 0; j < nblk; ++j) {
                    mbar_wait(&kv_empty[st], kph ^ 1);
                    mbar_arrive_expect_tx(&kv_full[st], 2 * TILE);
#pragma unroll
                    for (int x = 0; x < NB; ++x) {
                        tma_load_2d(smem_k + st * TILE + x * kFaBoxBytes, &tmap_qkv, &kv_full[st],
                                    p.k_col0 + h * D + x * 64, row0 + j * 128, kEvictLast);
                        tma_load_2d(smem_v + st * TILE + x * kFaBoxBytes, &tmap_qkv, &kv_full[st],
                                    p.v_col0 + h * D + x * 64, row0 + j * 128, kEvictLast);
                    }
                    if (++st == ST) { st = 0; kph ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {
            // ------------------------------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
            const uint32_t q_addr = smem_u32(smem_q), do_addr = smem_u32(smem_do);
            int st = 0, g = 0;      // g = key blocks issued so far by this CTA: TMEM buffer pair g & 1
            uint32_t kph = 0, qph = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                mbar_wait(q_full, qph);
                qph ^= 1;
                for (int j = 0; j < nblk; ++j, ++g) {
                    const int bp = g & 1;
                    mbar_wait(&sd_free[bp], ((g >> 1) & 1) ^ 1);   // the pair's previous block has been read
                    mbar_wait(&kv_full[st], kph);
                    tc_fence_after();
                    const uint32_t k_addr = smem_u32(smem_k + st * TILE), v_addr = smem_u32(smem_v + st * TILE);
#pragma unroll
                    for (int tt = 0; tt < D / 16; ++tt) {
                        const uint32_t off = (tt >> 2) * kFaBoxBytes + (tt & 3) * 32;
                        umma_bf16_ss(tmem_base + bp * 256, make_smem_desc_sw128(q_addr + off, 16, 1024),
                                     make_smem_desc_sw128(k_addr + off, 16, 1024), idesc, tt != 0);
                    }
#pragma unroll
                    for (int tt = 0; tt < D / 16; ++tt) {
                        const uint32_t off = (tt >> 2) * kFaBoxBytes + (tt & 3) * 32;
                        umma_bf16_ss(tmem_base + bp * 256 + 128, make_smem_desc_sw128(do_addr + off, 16, 1024),
                                     make_smem_desc_sw128(v_addr + off, 16, 1024), idesc, tt != 0);
                    }
                    umma_commit(&kv_empty[st]);
                    umma_commit(&sd_full[bp]);
                    if (j == nblk - 1) umma_commit(q_empty);
                    if (++st == ST) { st = 0; kph ^= 1; }
                }
            }
        }
    } else {
        // -------------------------------------------------------------------------------- softmax-backward warpgroups
        const int q = warp & 3;       // TMEM lane quadrant
        const int w = warp >> 2;      // warpgroup = parity of the CTA's key-block counter it serves
        const int r = q * 32 + lane;  // query row inside the tile
        const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16) + w * 256;
        int g0 = 0;                   // key blocks of this CTA before the current item
        uint32_t fph = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, g0 += nblk) {
            const int qt = item % qtiles, bh = item / qtiles;
            const int qrow = qt * 128 + r;
            const bool row_ok = qrow < S;
            const long long orow = (long long)bh * S + qrow;       // row of P / dS and of lse / delta
            const float lse = row_ok ? __ldg(p.lse + orow) : INFINITY;   // +inf -> P = 0 for rows outside the sequence
            const float dl = row_ok ? __ldg(p.delta + orow) : 0.f;
            const uint32_t drop_rs = kDrop ? drop_row_seed(drop_site_seed(p.drop), (uint32_t)orow) : 0u;
            const float drop_inv = kDrop ? drop_inv_keep(p.drop.thresh16) : 1.0f;
            for (int j = 0; j < nblk; ++j) {
                if (((g0 + j) & 1) != w) continue;
                mbar_wait(&sd_full[w], fph);
                fph ^= 1;
                tc_fence_after();
                const int key0 = j * 128;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t s[32], dp[32];
                    tmem_ld_x32(lane_base + c * 32, s);
                    tmem_ld_x32(lane_base + 128 + c * 32, dp);
                    tmem_ld_wait();
                    if (c == 3) {
                        // both score tiles are in registers: hand the TMEM pair back before the math
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sd_free[w]);
                    }
                    const int kc = key0 + c * 32;
                    uint32_t pk[16], dk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), p.scale_log2e, -lse));
                        float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), p.scale_log2e, -lse));
                        if (kc + i >= S) p0 = 0.f;          // keys beyond the sequence (rows of the next one / padding)
                        if (kc + i + 1 >= S) p1 = 0.f;
                        float dp0 = __uint_as_float(dp[i]), dp1 = __uint_as_float(dp[i + 1]);
                        float m0 = 1.f, m1 = 1.f;
                        if (kDrop) {
                            const uint32_t bits = drop_pair_bits(drop_rs, (uint32_t)(kc + i) >> 1);
                            m0 = (bits & 0xffffu) >= p.drop.thresh16 ? drop_inv : 0.f;
                            m1 = (bits >> 16) >= p.drop.thresh16 ? drop_inv : 0.f;
                            dp0 *= m0;
                            dp1 *= m1;
                        }
                        const float d0 = p.scale * p0 * (dp0 - dl);
                        const float d1 = p.scale * p1 * (dp1 - dl);
                        if (kDrop) {
                            p0 *= m0;
                            p1 *= m1;
                        }
                        pk[i >> 1] = pack_bf16x2(p0, p1);
                        dk[i >> 1] = pack_bf16x2(d0, d1);
                    }
                    // transpose through shared memory so that global stores are row-contiguous: a thread owns one row
                    // (64 B per matrix and chunk); written back as 8 rows x 64 B per instruction (full 32-byte sectors
                    // instead of 32 scattered 16-byte pieces).  16-byte units are XOR-swizzled: conflict free both ways.
                    uint8_t* stg = smem_out + (warp * 4096);
                    const int swz_w = (lane >> 1) & 3;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        *reinterpret_cast<uint4*>(stg + lane * 64 + ((v ^ swz_w) << 4)) =
                            make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
                        *reinterpret_cast<uint4*>(stg + 2048 + lane * 64 + ((v ^ swz_w) << 4)) =
                            make_uint4(dk[4 * v], dk[4 * v + 1], dk[4 * v + 2], dk[4 * v + 3]);
                    }
                    __syncwarp();
                    const int seg = lane & 3;
                    if (kc + seg * 8 < p.ldp) {            // ldp is a multiple of 8: whole 16-byte groups
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int rw = it * 8 + (lane >> 2);             // row inside this warp's 32
                            const int qr = qt * 128 + q * 32 + rw;
                            if (qr < S) {
                                const long long off = ((long long)bh * S + qr) * p.ldp + kc + seg * 8;
                                const int u16 = (seg ^ ((rw >> 1) & 3)) << 4;
                                *reinterpret_cast<uint4*>(p.p + off) = *reinterpret_cast<const uint4*>(stg + rw * 64 + u16);
                                *reinterpret_cast<uint4*>(p.ds + off) =
                                    *reinterpret_cast<const uint4*>(stg + 2048 + rw * 64 + u16);
                            }
                        }
                    }
                    __syncwarp();
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

// delta[(b * heads + h) * seq + i] = sum_d dO[b*seq + i, h*D + d] * O[b*seq + i, h*D + d]; one warp per (token, head)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ d_o, const __nv_bfloat16* __restrict__ o, long long ld,
                  float* __restrict__ delta, int batch, int seq, int heads, int head_dim) {
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const long long total = (long long)batch * seq * heads;
    if (wid >= total) return;
    const long long tok = wid / heads;
    const int h = (int)(wid % heads);
    const __nv_bfloat16* a = d_o + tok * ld + h * head_dim;
    const __nv_bfloat16* b = o + tok * ld + h * head_dim;
    float acc = 0.f;
    for (int d = lane * 2; d < head_dim; d += 64) {
        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + d));
        const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(b + d));
        acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        const long long bi = tok / seq, i = tok % seq;
        delta[(bi * heads + h) * seq + i] = acc;
    }
}

// Same result, one warp per TOKEN: 32 / heads lanes share a head, each lane reads its contiguous head_dim * heads / 32
// elements of dO and O with 16-byte loads (the per-(token, head) kernel above moves 4 bytes per lane and load: 37 us
// for the 54 MB of a 32 x 542 x 768 step against 8 us of HBM time).  Requires 32 % heads == 0 and whole chunks per lane.
__global__ void __launch_bounds__(256)
attn_delta_token_kernel(const __nv_bfloat16* __restrict__ d_o, const __nv_bfloat16* __restrict__ o, long long ld,
                        float* __restrict__ delta, int batch, int seq, int heads, int head_dim) {
    const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tok >= (long long)batch * seq) return;
    const int lph = 32 / heads;                     // lanes per head
    const int h = lane / lph, q = lane - h * lph;
    const int per = head_dim / lph;                 // elements of this lane: a multiple of 8
    const uint4* a = reinterpret_cast<const uint4*>(d_o + tok * ld + h * head_dim + q * per);
    const uint4* b = reinterpret_cast<const uint4*>(o + tok * ld + h * head_dim + q * per);
    float acc = 0.f;
    for (int c = 0; c < (per >> 3); ++c) {
        float x[8], y[8];
        bf16x8_to_float(__ldg(a + c), x);
        bf16x8_to_float(__ldg(b + c), y);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(x[j], y[j], acc);
    }
    for (int s = lph >> 1; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (q == 0) {
        const long long bi = tok / seq, i = tok % seq;
        delta[(bi * heads + h) * seq + i] = acc;
    }
}

}  // namespace fame

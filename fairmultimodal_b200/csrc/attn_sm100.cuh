// K2: fused attention forward for the note encoder (head_dim 64, seq <= 512) on tcgen05 / TMEM.
//
//   ctx[b, q, h, :] = softmax_k( Q.K^T * scale + mask_bias[b, k] ) . V
//
// Replaces F.scaled_dot_product_attention as called by BertSelfAttention (HF modeling_bert.py:192-206) with the
// bidirectional key-padding mask of HF:709-713, reached from 10_FAME.py:140.
//
// One CTA per (128-query tile, head, sequence).  Because seq <= 512 the whole score row fits in TMEM
// (128 lanes x 512 f32 columns), so no online-softmax rescaling is needed:
//   1. TMA: Q tile, all K blocks, all V blocks of this (sequence, head) -> smem (SW128)
//   2. MMA1 (SS): S[128, 128 j..] = Q . K_j^T                          -> TMEM columns [0, 512)
//   3. 8 softmax warps: row max, p = exp2(x - max), row sum; P (bf16) is written back to TMEM in place
//      (keys [0,256) -> columns [0,128), keys [256,512) -> columns [256,384))
//   4. MMA2 (TS): O[128, 64] = P[tmem] . V[smem, MN-major]             -> TMEM columns [128, 192)
//   5. epilogue: O / rowsum -> bf16 -> global
#pragma once
#include "sm100_ptx.cuh"

namespace fame {

constexpr int kAttnD = 64;
constexpr int kAttnBQ = 128;
constexpr int kAttnMaxS = 512;
constexpr int kAttnTileBytes = 128 * kAttnD * 2;  // 16 KB: 128 rows x 64 bf16
constexpr int kAttnThreads = 320;                 // 8 softmax warps + control warp + TMEM-alloc warp
constexpr int kAttnSmemBytes = kAttnTileBytes * 9 /*Q + 4K + 4V*/ + 2048 /*bias*/ + 2048 /*max,sum*/ + 256 + 1024;

struct AttnParams {
    const uint8_t* key_mask;  // [batch, seq] or nullptr
    __nv_bfloat16* ctx;
    long long ld_ctx;
    int batch, seq, heads;
    float scale_log2e;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_d64_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;
    uint8_t* smem_k = smem + kAttnTileBytes;
    uint8_t* smem_v = smem + 5 * kAttnTileBytes;
    float* bias = reinterpret_cast<float*>(smem + 9 * kAttnTileBytes);  // [512]
    float* red_max = bias + 512;                                        // [2][128]
    float* red_sum = red_max + 256;                                     // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red_sum + 256);
    uint64_t* bar_q = bars;        // Q landed
    uint64_t* bar_k = bars + 1;    // [4] K block j landed
    uint64_t* bar_v = bars + 5;    // all V landed
    uint64_t* bar_s = bars + 6;    // MMA1 complete (S in TMEM)
    uint64_t* bar_p = bars + 7;    // P written to TMEM by all 8 softmax warps
    uint64_t* bar_o = bars + 8;    // MMA2 complete (O in TMEM)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int S = p.seq;
    const int nkb = (S + 127) >> 7;  // key blocks of 128
    const int HD = p.heads * kAttnD;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_qkv);
        mbar_init(bar_q, 1);
        for (int j = 0; j < 4; ++j) mbar_init(&bar_k[j], 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 8);
        mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 9) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    // additive key bias: -inf beyond the sequence, -1e30 for user-masked keys, 0 otherwise
    for (int k = threadIdx.x; k < kAttnMaxS; k += kAttnThreads) {
        float v = 0.f;
        if (k >= S) v = -INFINITY;
        else if (p.key_mask != nullptr && p.key_mask[(long long)b * S + k] == 0) v = -1e30f;
        bias[k] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------ control thread: TMA + both MMAs
            const int row0 = b * S;
            mbar_arrive_expect_tx(bar_q, kAttnTileBytes);
            tma_load_2d(smem_q, &tmap_qkv, bar_q, h * kAttnD, row0 + qt * kAttnBQ, kEvictFirst);
            for (int j = 0; j < nkb; ++j) {
                mbar_arrive_expect_tx(&bar_k[j], kAttnTileBytes);
                tma_load_2d(smem_k + j * kAttnTileBytes, &tmap_qkv, &bar_k[j], HD + h * kAttnD, row0 + j * 128,
                            kEvictLast);
            }
            mbar_arrive_expect_tx(bar_v, nkb * kAttnTileBytes);
            for (int j = 0; j < nkb; ++j)
                tma_load_2d(smem_v + j * kAttnTileBytes, &tmap_qkv, bar_v, 2 * HD + h * kAttnD, row0 + j * 128,
                            kEvictLast);

            // MMA1: S_j = Q . K_j^T   (M=128, N=128, K=64 -> 4 instructions of K=16)
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
            mbar_wait(bar_q, 0);
            const uint32_t q_addr = smem_u32(smem_q);
            for (int j = 0; j < nkb; ++j) {
                mbar_wait(&bar_k[j], 0);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(smem_k + j * kAttnTileBytes);
#pragma unroll
                for (int k = 0; k < kAttnD / 16; ++k) {
                    const uint64_t adesc = make_smem_desc_sw128(q_addr + k * 32, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(k_addr + k * 32, 16, 1024);
                    umma_bf16_ss(tmem_base + j * 128, adesc, bdesc, idesc_qk, k != 0);
                }
            }
            umma_commit(bar_s);

            // MMA2: O = P . V   (A = P from TMEM, B = V MN-major from smem; 16 keys per instruction)
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, kAttnD, 0, 1);
            mbar_wait(bar_v, 0);
            mbar_wait(bar_p, 0);
            tc_fence_after();
            const uint32_t v_addr = smem_u32(smem_v);
            const int nk16 = nkb * 8;
            for (int t = 0; t < nk16; ++t) {
                const int key = t * 16;
                const uint32_t p_col = (key >> 8) * 256 + ((key & 255) >> 1);
                const uint64_t bdesc = make_smem_desc_sw128(v_addr + key * 128, 16, 1024);
                umma_bf16_ts(tmem_base + 128, tmem_base + p_col, bdesc, idesc_pv, t != 0);
            }
            umma_commit(bar_o);
        }
    } else if (warp < 8) {
        // ---------------------------------------------------- softmax + epilogue warps
        const int q = warp & 3;      // TMEM lane quarter
        const int half = warp >> 2;  // keys [256*half, 256*half + 256)
        const int r = q * 32 + lane; // row inside the query tile
        const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
        const int nchunk = max(0, min(8, ((nkb * 128) - half * 256) >> 5));  // 32-key chunks in my half

        mbar_wait(bar_s, 0);
        tc_fence_after();
        float m = -INFINITY;
        for (int c = 0; c < nchunk; ++c) {
            uint32_t rr[32];
            tmem_ld_x32(lane_base + half * 256 + c * 32, rr);
            tmem_ld_wait();
            const float* bc = bias + half * 256 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaf(__uint_as_float(rr[j]), p.scale_log2e, bc[j]));
        }
        red_max[half * 128 + r] = m;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        m = fmaxf(red_max[r], red_max[128 + r]);

        float sum = 0.f;
        for (int c = 0; c < nchunk; ++c) {
            uint32_t rr[32];
            tmem_ld_x32(lane_base + half * 256 + c * 32, rr);
            tmem_ld_wait();
            const float* bc = bias + half * 256 + c * 32;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(rr[j]), p.scale_log2e, bc[j]) - m);
                const float p1 = ex2_approx(fmaf(__uint_as_float(rr[j + 1]), p.scale_log2e, bc[j + 1]) - m);
                sum += p0 + p1;
                pk[j >> 1] = pack_bf16x2(p0, p1);
            }
            tmem_st_x16(lane_base + half * 256 + c * 16, pk);
        }
        tmem_st_wait();
        red_sum[half * 128 + r] = sum;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = 1.0f / (red_sum[r] + red_sum[128 + r]);

        mbar_wait(bar_o, 0);
        tc_fence_after();
        uint32_t oo[32];
        tmem_ld_x32(lane_base + 128 + half * 32, oo);
        tmem_ld_wait();
        const int qrow = qt * kAttnBQ + r;
        if (qrow < S) {
            __nv_bfloat16* dst = p.ctx + (long long)(b * S + qrow) * p.ld_ctx + h * kAttnD + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(oo[j]) * inv, __uint_as_float(oo[j + 1]) * inv);
                o.y = pack_bf16x2(__uint_as_float(oo[j + 2]) * inv, __uint_as_float(oo[j + 3]) * inv);
                o.z = pack_bf16x2(__uint_as_float(oo[j + 4]) * inv, __uint_as_float(oo[j + 5]) * inv);
                o.w = pack_bf16x2(__uint_as_float(oo[j + 6]) * inv, __uint_as_float(oo[j + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + j) = o;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace fame

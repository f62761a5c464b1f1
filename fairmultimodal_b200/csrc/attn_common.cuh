// Shared by the attention forward (attn_pair_sm100.cuh) and backward (attn_bwd_sm100.cuh) kernels: the TMA box
// geometry of a packed [tokens, 3 * heads * head_dim] qkv tensor, the forward's parameter block and the MUFU exp2.
#pragma once
#include "dropout.cuh"
#include "sm100_ptx.cuh"

namespace fame {

constexpr int kFaBoxBytes = 128 * 64 * 2;  // one TMA box: 128 rows x 64 bf16 columns, SW128
// head_dim 96 rows (192 B) do not fit one 128-byte swizzle atom: every tile is loaded as two 64-column boxes (the
// second box is only half used; its extra columns are never addressed by an MMA).

struct FaParams {
    const uint8_t* key_mask;  // [batch, seq] (1 = attend) or nullptr
    __nv_bfloat16* ctx;
    long long ld_ctx;
    int batch, seq, heads;
    int q_col0, k_col0, v_col0;  // first column of Q / K / V of head 0 inside the packed tensor
    float scale_log2e;
    float* lse;               // optional [batch, heads, seq]: row log-sum-exp in log2 units of the scaled scores
    const int* kv_len;        // optional [batch]: 1 + last attended key (fully masked key blocks are skipped)
    DropCfg drop;             // dropout of the attention probabilities (kDrop instantiation only)
    int pingpong;             // 1: the two softmax warpgroups take turns on the exponential phase (see attn_pair_sm100.cuh)
};

// named-barrier halves of a turn hand-over between two 128-thread warpgroups (256 = waiting + signalling threads)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace fame

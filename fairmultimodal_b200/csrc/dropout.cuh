// Dropout masks of the training step, generated on the fly from a counter-based hash -- never stored: the forward
// kernel that applies a mask and the backward kernel that needs it again evaluate the same function of
// (site seed, step counter, row, column).  Replaces the nn.Dropout / SDPA dropout_p / MultiheadAttention dropout
// draws of the reference in train() mode (HF modeling_bert.py:111,205,297,355; torch TransformerEncoderLayer dropout,
// dropout1, dropout2 and self_attn.dropout; fusion_mlp[2], 10_FAME.py:255).  torch's Philox stream cannot be
// reproduced bit for bit by any other implementation; what is reproduced is the distribution: independent
// Bernoulli(1 - p) keeps, kept values scaled by 1 / (1 - p).
//
//   seed'   = mix32(seed + step * 0x632BE5AB)                    step read from device memory (CUDA-graph replays)
//   rowseed = mix32(seed' ^ (row * 0x9E3779B9))
//   unit    = column >> group_shift                               (group_shift 6: one draw per 64-wide attention head)
//   bits    = mix32(rowseed + (unit >> 1) * 0x85EBCA6B)           one 32-bit hash serves two units (16 bits each)
//   keep    = (unit & 1 ? bits >> 16 : bits & 0xffff) >= thresh16 thresh16 = round(p * 65536); 0 disables the site
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace fame {

struct DropCfg {
    const int* step;        // device step counter, may be nullptr (= 0)
    uint32_t seed;
    uint32_t thresh16;      // 0 = off
    int group_shift;
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {   // "lowbias32" integer finalizer
    h ^= h >> 16;
    h *= 0x7feb352dU;
    h ^= h >> 15;
    h *= 0x846ca68bU;
    h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t drop_site_seed(const DropCfg& c) {
    const uint32_t step = c.step != nullptr ? (uint32_t)__ldg(c.step) : 0u;
    return mix32(c.seed + step * 0x632BE5ABU);
}
__device__ __forceinline__ uint32_t drop_row_seed(uint32_t site_seed, uint32_t row) {
    return mix32(site_seed ^ (row * 0x9E3779B9U));
}
// 32 bits covering units 2 * pair and 2 * pair + 1 of a row
__device__ __forceinline__ uint32_t drop_pair_bits(uint32_t row_seed, uint32_t pair) {
    return mix32(row_seed + pair * 0x85EBCA6BU);
}
__device__ __forceinline__ bool drop_keep(uint32_t row_seed, uint32_t unit, uint32_t thresh16) {
    const uint32_t bits = drop_pair_bits(row_seed, unit >> 1);
    return ((unit & 1u) ? (bits >> 16) : (bits & 0xffffu)) >= thresh16;
}
__host__ __device__ __forceinline__ float drop_inv_keep(uint32_t thresh16) {
    return 65536.0f / (65536.0f - (float)thresh16);
}

// In-place dropout of a [rows, cols] tensor (bf16 or f32): the small sites that have no producing GEMM epilogue
// (embedding output of the demographic BERT, the fusion head's hidden layer and their gradients).
template <bool kF32>
__global__ void __launch_bounds__(256)
dropout_apply_kernel(void* __restrict__ x, long long ld, int rows, int cols, const DropCfg cfg) {
    const uint32_t site = drop_site_seed(cfg);
    const float inv = drop_inv_keep(cfg.thresh16);
    const long long total = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        const bool keep = drop_keep(drop_row_seed(site, (uint32_t)r), (uint32_t)c >> cfg.group_shift, cfg.thresh16);
        if (kF32) {
            float* p = reinterpret_cast<float*>(x) + (long long)r * ld + c;
            *p = keep ? *p * inv : 0.f;
        } else {
            __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(x) + (long long)r * ld + c;
            *p = keep ? __float2bfloat16_rn(__bfloat162float(*p) * inv) : __float2bfloat16_rn(0.f);
        }
    }
}

}  // namespace fame

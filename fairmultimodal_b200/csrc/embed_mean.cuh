// Demographic code embeddings, generalised to n tables (<= 8): the structured encoder of the average-fusion ablation
// (07_multimodal_average_fusion.py:156-203) adds the mean of SEVEN clamped embedding rows (age, segment, admission
// location, discharge location, gender, ethnicity, insurance) to the CLS state; 10_FAME.py:194-206 is the n = 4 case.
//
//   forward   out[b, :] = cls[b, :] + (1 / n) * sum_k table_k[clamp(ids_k[b], 0, rows_k - 1), :]
//   backward  dtable_k[clamp(ids_k[b]), :] += dout[b, :] / n        (d cls = dout: the caller reuses the tensor)
//
// One block per patient, fp32; summation order over k is fixed (k = 0 first), as torch evaluates (a + b + ...) / n.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fame {

constexpr int kEmMaxTables = 8;

struct EmbedMeanParams {
    const void* cls;                          // [batch, hidden] bf16 or f32, row stride ld_cls (elements)
    long long ld_cls;
    int cls_f32;
    int n_tables;
    const long long* ids[kEmMaxTables];       // [batch] each
    const float* table[kEmMaxTables];         // [rows_k, hidden]
    float* dtable[kEmMaxTables];              // backward only
    int rows[kEmMaxTables];
    float* out;                               // forward: [batch, hidden]
    const float* dout;                        // backward: [batch, hidden]
    int hidden;
};

__device__ __forceinline__ long long em_clamp(long long v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }

__global__ void __launch_bounds__(128)
embed_mean_add_kernel(const EmbedMeanParams p) {
    const int b = blockIdx.x;
    const float* row[kEmMaxTables];
#pragma unroll
    for (int k = 0; k < kEmMaxTables; ++k)
        row[k] = k < p.n_tables ? p.table[k] + em_clamp(p.ids[k][b], p.rows[k]) * p.hidden : nullptr;
    const float n = (float)p.n_tables;
    for (int c = threadIdx.x; c < p.hidden; c += blockDim.x) {
        float extra = 0.f;
#pragma unroll
        for (int k = 0; k < kEmMaxTables; ++k)
            if (k < p.n_tables) extra += row[k][c];
        const float cv = p.cls_f32 ? reinterpret_cast<const float*>(p.cls)[(long long)b * p.ld_cls + c]
                                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.cls)[(long long)b * p.ld_cls + c]);
        p.out[(long long)b * p.hidden + c] = cv + extra / n;   // (e_0 + ... + e_{n-1}) / n, as the reference divides
    }
}

__global__ void __launch_bounds__(128)
embed_mean_add_bwd_kernel(const EmbedMeanParams p) {
    const int b = blockIdx.x;
    long long r[kEmMaxTables];
#pragma unroll
    for (int k = 0; k < kEmMaxTables; ++k) r[k] = k < p.n_tables ? em_clamp(p.ids[k][b], p.rows[k]) : 0;
    const float inv = 1.0f / (float)p.n_tables;
    for (int c = threadIdx.x; c < p.hidden; c += blockDim.x) {
        const float g = p.dout[(long long)b * p.hidden + c] * inv;
#pragma unroll
        for (int k = 0; k < kEmMaxTables; ++k)
            if (k < p.n_tables) atomicAdd(p.dtable[k] + r[k] * p.hidden + c, g);
    }
}

}  // namespace fame

// K10 evaluation metric kernels: thresholded confusion counts per (outcome, sensitive attribute, subgroup) for
// EDDI / Equalized Odds, the 101-threshold F1 sweep histogram, and exact tie-aware AUROC / average-precision rank
// counts.  Replaces the numpy / sklearn passes of compute_eddi (10_FAME.py:54-82), calculate_tpr_and_fpr (84-97),
// calibrate_thresholds (470-481) and evaluate_model_multi (514-540).  All outputs are INTEGER counts (bit-exact,
// order independent, summable across ranks); the few float64 divisions that turn counts into EDDI / EO / F1 / AUROC
// happen on the host exactly as the reference does them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rowwise.cuh"

namespace fame {

// float32 sigmoid evaluated the way torch.sigmoid does on float32 tensors: 1 / (1 + exp(-z)) with every step
// rounded to float32 (exp correctly rounded via a float64 evaluation, IEEE add and divide, no FMA contraction).
// The reference thresholds and ranks these float32 probabilities (10_FAME.py:471-476, 516-518), so both the strict
// `>` decisions and the tie structure in the saturated tails (1 + e rounds to 1 for e < 2^-24) follow this form.
__device__ __forceinline__ float sigmoid_f32_exact(float z) {
    const float e = (float)exp(-(double)z);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
}

constexpr int kEvSlots = 8;                               // subgroup codes 0..7
constexpr int kEvConfLen = 3 * 3 * kEvSlots * 4;          // [outcome][attr][code][TP,FN,FP,TN]
constexpr int kEvTotLen = 3 * 4;                          // [outcome][TP,FN,FP,TN]
constexpr int kEvHistLen = 3 * 2 * 102;                   // [outcome][label][#thresholds strictly below p]
constexpr int kEvLen = kEvConfLen + kEvTotLen + kEvHistLen + 2;  // + patients + error flag

struct EvalCountsParams {
    const float* logits;        // [N, ld] (3 outcomes used)
    long long ld;
    const float* labels;        // [N,3]
    const long long* attr[3];   // [N] each
    double thr[3];              // decision threshold per outcome; prediction = (double)p_f32 > thr
    const double* sweep;        // [101] ascending thresholds for the F1 sweep, nullable (histogram skipped)
    unsigned long long* out;    // [kEvLen], zero-initialised
    int N;
    int logits_are_probs;       // 1: `logits` already holds float32 probabilities / 0-1 predictions
};

// One shared-memory histogram per sensitive attribute, keyed by (subgroup code, confusion cell of outcome 0, of
// outcome 1, of outcome 2) = 8 x 4 x 4 x 4 = 512 bins: a patient costs THREE shared-memory increments instead of 72
// predicated register adds (3 outcomes x 3 attributes x 8 code slots), and the kernel needs ~40 registers instead of
// 163, so 6+ CTAs per SM hide the load latency.  The 288 + 12 confusion counts are folded out of the 3 x 512 bins once
// per block.  Each thread takes 4 consecutive patients per trip through 16-byte loads (load_patient_quad, rowwise.cuh).
constexpr int kEvBins = kEvSlots * 64;

__global__ void __launch_bounds__(256)
eval_counts_kernel(const EvalCountsParams p) {
    __shared__ unsigned int hist[3][kEvBins];
    // F1-sweep histogram, one copy per warp (8 x 612 counters): the three increments of a patient then only contend with
    // the other lanes of their own warp (r01: one shared copy, 46 % of the copy bandwidth against 72 % without the sweep)
    __shared__ unsigned int sw_hist[8][kEvHistLen];
    // thresholds as float32 ROUNDED DOWN: for a float32 p and a float64 t, p > t <=> p > (largest float32 <= t), so the
    // strict float64 compare of the reference (10_FAME.py:476) is decided exactly by a float32 compare
    __shared__ float sweep_s[101];
    __shared__ unsigned int s_misc[2];   // patients, bad-code flag
    for (int i = threadIdx.x; i < 3 * kEvBins; i += blockDim.x) (&hist[0][0])[i] = 0u;
    for (int i = threadIdx.x; i < 8 * kEvHistLen; i += blockDim.x) (&sw_hist[0][0])[i] = 0u;
    if (threadIdx.x < 2) s_misc[threadIdx.x] = 0u;
    if (p.sweep != nullptr)
        for (int i = threadIdx.x; i < 101; i += blockDim.x) sweep_s[i] = __double2float_rd(p.sweep[i]);
    __syncthreads();
    unsigned int* my_sw = sw_hist[threadIdx.x >> 5];

    const int lane = threadIdx.x & 31;
    unsigned n_local = 0, bad = 0;
    const bool vec_ok = p.ld == 3 && ((reinterpret_cast<uintptr_t>(p.logits) | reinterpret_cast<uintptr_t>(p.labels) |
                                       reinterpret_cast<uintptr_t>(p.attr[0]) | reinterpret_cast<uintptr_t>(p.attr[1]) |
                                       reinterpret_cast<uintptr_t>(p.attr[2])) & 15) == 0;
    const long long per_block = (long long)blockDim.x * kPatPerThread;
    for (long long base = blockIdx.x * per_block; base < p.N; base += gridDim.x * per_block) {
        PatientQuad q;
        load_patient_quad(q, p.logits, p.ld, p.labels, p.attr, base + (long long)threadIdx.x * kPatPerThread, p.N, vec_ok);
#pragma unroll
        for (int u = 0; u < kPatPerThread; ++u) {
            if (u < q.n) {
                unsigned cells = 0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float z = q.z[u][i];
                    const float pr = p.logits_are_probs ? z : sigmoid_f32_exact(z);
                    const int y = q.y[u][i] != 0.f;
                    const int pred = (double)pr > p.thr[i];
                    cells |= (unsigned)((1 - y) * 2 + (1 - pred)) << (2 * i);   // cell: TP=0, FN=1, FP=2, TN=3
                    if (p.sweep != nullptr) {
                        // kk = number of sweep thresholds strictly below p (p > t_k <=> k < kk); thresholds ascend.
                        // Start from the bin a uniform grid would give and walk (at most a step or two).
                        int kk = min(101, max(0, (int)(pr * 100.0f)));
                        while (kk < 101 && pr > sweep_s[kk]) ++kk;
                        while (kk > 0 && !(pr > sweep_s[kk - 1])) --kk;
                        atomicAdd(&my_sw[(i * 2 + y) * 102 + kk], 1u);
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const long long c = q.code[a][u];
                    bad |= (c < 0 || c >= kEvSlots);
                    atomicAdd(&hist[a][(((int)c & (kEvSlots - 1)) << 6) | cells], 1u);
                }
                ++n_local;
            }
        }
    }
    {
        const unsigned v = __reduce_add_sync(0xffffffffu, n_local);
        const unsigned e = __reduce_or_sync(0xffffffffu, bad);
        if (lane == 0) {
            atomicAdd(&s_misc[0], v);
            if (e) atomicOr(&s_misc[1], 1u);
        }
    }
    __syncthreads();
    // fold the bins: conf[i][a][code][cell] = sum over the cells of the other two outcomes; tot[i][cell] from attr 0
    for (int idx = threadIdx.x; idx < kEvConfLen + kEvTotLen; idx += blockDim.x) {
        unsigned long long sum = 0ull;
        if (idx < kEvConfLen) {
            const int cell = idx & 3, code = (idx >> 2) & 7, ia = idx >> 5, a = ia % 3, i = ia / 3;
            for (int o = 0; o < 16; ++o) {
                // spread the 4 bits of o over the two outcome positions other than i
                const int lo = o & 3, hi = o >> 2;
                const int c0 = i == 0 ? cell : lo, c1 = i == 1 ? cell : (i == 0 ? lo : hi), c2 = i == 2 ? cell : hi;
                sum += hist[a][(code << 6) | c0 | (c1 << 2) | (c2 << 4)];
            }
        } else {
            const int t = idx - kEvConfLen, cell = t & 3, i = t >> 2;
            for (int code = 0; code < kEvSlots; ++code)
                for (int o = 0; o < 16; ++o) {
                    const int lo = o & 3, hi = o >> 2;
                    const int c0 = i == 0 ? cell : lo, c1 = i == 1 ? cell : (i == 0 ? lo : hi), c2 = i == 2 ? cell : hi;
                    sum += hist[0][(code << 6) | c0 | (c1 << 2) | (c2 << 4)];
                }
        }
        if (sum) atomicAdd(p.out + idx, sum);
    }
    if (p.sweep != nullptr)
        for (int i = threadIdx.x; i < kEvHistLen; i += blockDim.x) {
            unsigned long long v = 0ull;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sw_hist[w][i];
            if (v) atomicAdd(p.out + kEvConfLen + kEvTotLen + i, v);
        }
    if (threadIdx.x == 0) {
        if (s_misc[0]) atomicAdd(p.out + kEvLen - 2, (unsigned long long)s_misc[0]);
        if (s_misc[1]) atomicOr(p.out + kEvLen - 1, 1ull);
    }
}

// ------------------------------------------------------------------------------------------------ rank counts
// For every sample i in [i0, i1) of one outcome:  ge_pos = #{j : y_j = 1, s_j >= s_i}, ge_neg, gt_neg (strict).
//   AUROC = sum_{i pos} [ (Nneg - ge_neg_i) + (ge_neg_i - gt_neg_i) / 2 ] / (Npos * Nneg)   (ties count one half:
//           identical to sklearn's trapezoid over distinct thresholds)
//   AP    = (1 / Npos) sum_{i pos} ge_pos_i / (ge_pos_i + ge_neg_i)    (sklearn: sum_n (R_n - R_{n-1}) P_n)
// Brute force O(N^2 / ranks) with the j scores staged through shared memory: exact, sort-free, and shardable over
// i across GPUs (each rank needs all j scores = an all-gather of N floats per outcome).
struct RankParams {
    const float* scores;   // [N] probabilities of this outcome (all samples)
    const uint8_t* y;      // [N] labels 0/1
    int N, i0, i1;
    unsigned long long* auroc2;  // += sum_{i pos in range} 2 (Nneg - ge_neg) + (ge_neg - gt_neg)
    double* ap_partial;          // [gridDim.x] per-block partial sums of precision at each positive
    unsigned long long* npos_nneg;  // [2] += positives / negatives among [i0, i1)
};

constexpr int kRankTile = 2048;

__global__ void __launch_bounds__(256)
rank_counts_kernel(const RankParams p) {
    __shared__ float sj[kRankTile];
    __shared__ uint8_t yj[kRankTile];
    __shared__ double red_ap[8];
    __shared__ unsigned long long red_au[8];
    const int i = p.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < p.i1;
    const float si = live ? p.scores[i] : 0.f;
    const int yi = live ? p.y[i] : 0;
    unsigned ge_pos = 0, ge_neg = 0, gt_neg = 0;
    for (int j0 = 0; j0 < p.N; j0 += kRankTile) {
        const int n = min(kRankTile, p.N - j0);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            sj[j] = p.scores[j0 + j];
            yj[j] = p.y[j0 + j];
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < n; ++j) {
            const float s = sj[j];
            const unsigned pos = yj[j];
            const unsigned ge = s >= si, gt = s > si;
            ge_pos += ge & pos;
            ge_neg += ge & (pos ^ 1u);
            gt_neg += gt & (pos ^ 1u);
        }
    }
    // totals of negatives are needed per i: Nneg = total negatives (all j) -- recomputed from the last pass
    // by the finalizer, so emit the two tie-aware terms separately: A_i = ge_neg_i + gt_neg_i (= 2 gt + ties)
    double ap = 0.0;
    unsigned long long au = 0ull, np = 0ull, nn = 0ull;
    if (live) {
        if (yi) {
            ap = (double)ge_pos / (double)(ge_pos + ge_neg);
            au = (unsigned long long)ge_neg + (unsigned long long)gt_neg;  // subtracted from 2 Nneg per positive
            np = 1;
        } else {
            nn = 1;
        }
    }
    // block reduction (fixed order -> deterministic partials)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ap += __shfl_xor_sync(0xffffffffu, ap, o);
        au += __shfl_xor_sync(0xffffffffu, au, o);
        np += __shfl_xor_sync(0xffffffffu, np, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
    }
    if (lane == 0) {
        red_ap[warp] = ap;
        red_au[warp] = au;
        atomicAdd(p.npos_nneg, np);
        atomicAdd(p.npos_nneg + 1, nn);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        unsigned long long u = 0ull;
        for (int w = 0; w < 8; ++w) {
            a += red_ap[w];
            u += red_au[w];
        }
        p.ap_partial[blockIdx.x] = a;
        atomicAdd(p.auroc2, u);
    }
}

// ------------------------------------------------------------------------------------------------ rank counts, sorted
// The same statistics in O(N log N) for the FULL range [0, N): key = (score bits << 1) | label (scores are probabilities
// in [0, 1]: their fp32 bit patterns are non-negative and order like the values), bitonic sort of the keys, then one pass
// over the sorted keys in DESCENDING score order.  For a run of tied scores G (all keys with the same score bits), with
// cpos / cneg = positives / negatives counted up to and including the run:
//     every positive of G has  ge_pos = cpos(G), ge_neg = cneg(G), gt_neg = cneg(before G)
//     ap_sum += npos(G) * cpos(G) / (cpos(G) + cneg(G))          auroc2 += npos(G) * (cneg(G) + cneg(before G))
// Exactly the brute-force kernel's integers; the float64 AP sum runs in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
rank_keys_kernel(const float* __restrict__ scores, const uint8_t* __restrict__ y, int N, int npad,
                 uint32_t* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    keys[i] = i < N ? ((__float_as_uint(scores[i]) << 1) | (y[i] != 0 ? 1u : 0u)) : 0xffffffffu;   // padding sorts last
}

constexpr int kSortTile = 4096;      // keys per CTA in the shared-memory stages (1024 threads x 4)

// all compare-exchange stages with stride < kSortTile for the merge sizes k = k_lo .. k_hi (one launch sorts every
// 4096-key tile completely when k_lo = 2, k_hi = kSortTile; later launches finish a merge whose large strides ran in
// global memory)
__global__ void __launch_bounds__(1024)
bitonic_smem_kernel(uint32_t* __restrict__ keys, int k_lo, int k_hi, int j_hi) {
    __shared__ uint32_t sh[kSortTile];
    const int base = blockIdx.x * kSortTile;
    for (int t = threadIdx.x; t < kSortTile; t += 1024) sh[t] = keys[base + t];
    __syncthreads();
    for (int k = k_lo; k <= k_hi; k <<= 1) {
        for (int j = min(k >> 1, j_hi); j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < kSortTile / 2; t += 1024) {
                const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
                const bool up = (((base + lo) & k) == 0);
                const uint32_t a = sh[lo], b = sh[hi];
                if ((a > b) == up) {
                    sh[lo] = b;
                    sh[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < kSortTile; t += 1024) keys[base + t] = sh[t];
}

// one compare-exchange stage with stride j >= kSortTile of merge size k, in global memory
__global__ void __launch_bounds__(256)
bitonic_global_kernel(uint32_t* __restrict__ keys, int npad, int k, int j) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npad / 2) return;
    const int lo = ((t / j) * 2 * j) + (t % j), hi = lo + j;
    const bool up = ((lo & k) == 0);
    const uint32_t a = keys[lo], b = keys[hi];
    if ((a > b) == up) {
        keys[lo] = b;
        keys[hi] = a;
    }
}

// one CTA walks the sorted keys from the largest score down, 1024 positions per step
__global__ void __launch_bounds__(1024)
rank_scan_kernel(const uint32_t* __restrict__ keys, int N, uint32_t* __restrict__ cpos_g, unsigned long long* auroc2,
                 double* ap_sum, unsigned long long* npos_nneg) {
    __shared__ uint32_t wsum[32], wmax[32];
    __shared__ double red[32];
    __shared__ unsigned long long redu[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry_pos = 0, carry_start = 0;
    double ap = 0.0;
    unsigned long long au = 0ull;
    for (int d0 = 0; d0 < N; d0 += 1024) {
        const int d = d0 + threadIdx.x;                 // position in descending order <-> sorted index N - 1 - d
        const bool live = d < N;
        const uint32_t key = live ? keys[N - 1 - d] : 0u;
        const uint32_t sc = key >> 1, pos = live ? (key & 1u) : 0u;
        const bool starts = live && (d == 0 || (keys[N - d] >> 1) != sc);           // previous position has another score
        const bool ends = live && (d == N - 1 || (keys[N - 2 - d] >> 1) != sc);     // next position has another score
        // inclusive scans over the 1024 positions: positives (sum) and latest run start (max)
        uint32_t ps = pos, st = starts ? (uint32_t)d : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, ps, o), b = __shfl_up_sync(0xffffffffu, st, o);
            if (lane >= o) {
                ps += a;
                st = max(st, b);
            }
        }
        if (lane == 31) {
            wsum[warp] = ps;
            wmax[warp] = st;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t a = wsum[lane], b = wmax[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, a, o), y2 = __shfl_up_sync(0xffffffffu, b, o);
                if (lane >= o) {
                    a += x;
                    b = max(b, y2);
                }
            }
            wsum[lane] = a;
            wmax[lane] = b;
        }
        __syncthreads();
        const uint32_t cpos = carry_pos + ps + (warp > 0 ? wsum[warp - 1] : 0u);
        const uint32_t start = max(max(carry_start, st), warp > 0 ? wmax[warp - 1] : 0u);
        if (live) cpos_g[d] = cpos;
        __syncthreads();                                  // cpos of this step visible to the whole CTA (global, same block)
        if (ends) {
            const uint32_t before = start > 0 ? cpos_g[start - 1] : 0u;      // positives before the run
            const uint32_t npos_g = cpos - before;
            if (npos_g) {
                const uint32_t cneg = (uint32_t)(d + 1) - cpos, cneg_before = start - before;
                ap += (double)npos_g * ((double)cpos / (double)(d + 1));
                au += (unsigned long long)npos_g * ((unsigned long long)cneg + cneg_before);
            }
        }
        carry_pos += wsum[31];
        carry_start = max(carry_start, wmax[31]);
        __syncthreads();
    }
    // fixed-order block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ap += __shfl_xor_sync(0xffffffffu, ap, o);
        au += __shfl_xor_sync(0xffffffffu, au, o);
    }
    if (lane == 0) {
        red[warp] = ap;
        redu[warp] = au;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        unsigned long long u = 0ull;
        for (int w = 0; w < 32; ++w) {
            a += red[w];
            u += redu[w];
        }
        *ap_sum += a;
        *auroc2 += u;
        npos_nneg[0] += carry_pos;
        npos_nneg[1] += (unsigned long long)N - carry_pos;
    }
}

// fixed-order sum of the per-block partials (deterministic), accumulated into *dst
__global__ void sum_partials_kernel(const double* __restrict__ part, int n, double* __restrict__ dst) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += part[i];
        *dst += s;
    }
}

// probs[o][n] = float32 sigmoid of logits[n, o]; y8[o][n] = labels[n, o] != 0   (outcome-major for the rank kernel)
__global__ void sigmoid_probs_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ labels,
                                     float* __restrict__ probs, uint8_t* __restrict__ y8, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        probs[(long long)o * N + n] = sigmoid_f32_exact(logits[n * ld + o]);
        if (y8 != nullptr) y8[(long long)o * N + n] = labels[3ll * n + o] != 0.f;
    }
}

}  // namespace fame

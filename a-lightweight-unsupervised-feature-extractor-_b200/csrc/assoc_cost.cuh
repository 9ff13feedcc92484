// Association cost: history-bank appearance term + box / confidence terms (device functions).
//
// Appearance (Tracking.build_C_app_topk, model/mainTracking.py:141-211): for track i with a bank
// of T <= hist_max unit embeddings and detection j,
//     C_app[i][j] = 1 - mean(top-k over t of <bank_t / |bank_t|, det_j / |det_j|>),  k = min(topk, T).
// One CTA handles one track row against a tile of kTileN detections: bank rows and detection
// rows are staged (and re-normalised) in shared memory, the T x kTileN similarities are an fp32
// FFMA register-tiled product (fp32 keeps the 1e-5 parity bar; TF32/BF16 tensor-core math does
// not), and the top-k mean is a per-column selection over the T values.
// Box / confidence terms: bbox_cost, conf_cost, cal_cost of model/utils/costTool/costCard.py.
#pragma once
#include "common.cuh"

namespace b200 {
namespace cost {

constexpr int kD = B200_EMB_DIM;      // 128
constexpr int kTileN = 64;            // detections per CTA
constexpr int kThreads = 256;
constexpr int kDetStride = kD + 4;    // padded row (floats): conflict-free 16 B loads across rows
constexpr int kMaxBank = 64;          // hist_max supported by the shared-memory layout

__host__ __device__ inline int bank_cap(int T) { return T <= 32 ? 32 : 64; }
__host__ __device__ inline size_t smem_bytes(int T) {
    const int tc = bank_cap(T);
    return sizeof(float) * ((size_t)tc * kD + (size_t)kTileN * kDetStride + (size_t)tc * (kTileN + 1));
}

// row / (|row| + 1e-12) for one 128-float row handled by one warp (lane owns a float4).
__device__ __forceinline__ float4 unit_row(float4 v) {
    float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    s = warp_sum(s);
    const float n = __fadd_rn(sqrtf(s), 1e-12f);
    return make_float4(__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n), __fdiv_rn(v.w, n));
}

// Shared-memory layout of one (track row) x (detection tile) problem:
//   sBank [bank_cap(T)][128]  unit bank rows (rows >= T zero)
//   sDet  [kTileN][kDetStride] unit detection rows (rows beyond N zero)
//   sSim  [bank_cap(T)][kTileN + 1]
// sims_and_topk() turns staged sBank / sDet into C_app for column threadIdx.x (< kTileN).  Every
// thread of the kThreads-wide CTA must call it; it ends with a barrier so smem can be reused.
__device__ inline float sims_and_topk(const float* sBank, const float* sDet, float* sSim, int tc, int T, int topk,
                                      bool topk_mean) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;            // 16 x 16 threads; 2 bank rows x 4 columns each
    for (int tb = 0; tb < tc; tb += 32) {
        float acc[2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
        const float4* a0 = reinterpret_cast<const float4*>(sBank + (tb + ty) * kD);
        const float4* a1 = reinterpret_cast<const float4*>(sBank + (tb + ty + 16) * kD);
#pragma unroll 4
        for (int k = 0; k < kD / 4; ++k) {
            const float4 x0 = a0[k], x1 = a1[k];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 y = reinterpret_cast<const float4*>(sDet + (tx + 16 * q) * kDetStride)[k];
                acc[0][q] = fmaf(x0.x, y.x, acc[0][q]); acc[0][q] = fmaf(x0.y, y.y, acc[0][q]);
                acc[0][q] = fmaf(x0.z, y.z, acc[0][q]); acc[0][q] = fmaf(x0.w, y.w, acc[0][q]);
                acc[1][q] = fmaf(x1.x, y.x, acc[1][q]); acc[1][q] = fmaf(x1.y, y.y, acc[1][q]);
                acc[1][q] = fmaf(x1.z, y.z, acc[1][q]); acc[1][q] = fmaf(x1.w, y.w, acc[1][q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sSim[(tb + ty) * (kTileN + 1) + tx + 16 * q] = acc[0][q];
            sSim[(tb + ty + 16) * (kTileN + 1) + tx + 16 * q] = acc[1][q];
        }
    }
    __syncthreads();
    float result = 0.0f;
    if (tid < kTileN) {
        const float* col = sSim + tid;
        const int k = min(topk, T);
        if (!topk_mean) {
            float m = col[0];
            for (int t = 1; t < T; ++t) m = fmaxf(m, col[t * (kTileN + 1)]);
            result = __fsub_rn(1.0f, m);
        } else {
            unsigned long long taken = 0ull;
            float sum = 0.0f;
            for (int s = 0; s < k; ++s) {             // k passes of "largest not yet taken"
                float m = -__int_as_float(0x7f800000);
                int arg = 0;
                for (int t = 0; t < T; ++t) {
                    const float v = col[t * (kTileN + 1)];
                    if (!((taken >> t) & 1ull) && v > m) { m = v; arg = t; }
                }
                taken |= 1ull << arg;
                sum = __fadd_rn(sum, m);
            }
            result = __fsub_rn(1.0f, __fdiv_rn(sum, (float)k));
        }
    }
    __syncthreads();
    return result;
}

// Stages bank rows [T][128] and detection rows j0 .. j0+kTileN of det [N][128], then computes
// C_app for column j0 + threadIdx.x (valid for threadIdx.x < kTileN and j0 + threadIdx.x < N).
template <bool kNormDet>
__device__ inline float app_cost_tile(const float* __restrict__ bank, int T, const float* __restrict__ det,
                                      int N, int j0, int topk, bool topk_mean, float* smem) {
    const int tc = bank_cap(T);
    float* sBank = smem;
    float* sDet = sBank + tc * kD;
    float* sSim = sDet + kTileN * kDetStride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int t = warp; t < tc; t += kThreads / 32) {
        float4 v = make_float4(0, 0, 0, 0);
        if (t < T) v = unit_row(reinterpret_cast<const float4*>(bank + (size_t)t * kD)[lane]);
        reinterpret_cast<float4*>(sBank + t * kD)[lane] = v;
    }
    for (int j = warp; j < kTileN; j += kThreads / 32) {
        float4 v = make_float4(0, 0, 0, 0);
        if (j0 + j < N) {
            v = reinterpret_cast<const float4*>(det + (size_t)(j0 + j) * kD)[lane];
            if (kNormDet) v = unit_row(v);
        }
        reinterpret_cast<float4*>(sDet + j * kDetStride)[lane] = v;
    }
    __syncthreads();
    return sims_and_topk(sBank, sDet, sSim, tc, T, topk, topk_mean);
}

struct PairWeights {
    float w_app, w_bbox, w_conf, alpha, beta, conf_eps;
};

struct PairCost {
    float total, bbox, center, scale, conf;
};

// bbox_cost (costCard.py:141-173), conf_cost (:196-201) and the weighted sum (:264-268) for one
// (track, detection) pair, float32 with the reference's operation order and no FMA contraction.
__device__ __forceinline__ PairCost pair_cost(const float* bp, const float* bc, float cp, float cc,
                                              const PairWeights& w, float c_app) {
    PairCost o;
    const float pcx = __fmul_rn(0.5f, __fadd_rn(bp[0], bp[2])), pcy = __fmul_rn(0.5f, __fadd_rn(bp[1], bp[3]));
    const float ccx = __fmul_rn(0.5f, __fadd_rn(bc[0], bc[2])), ccy = __fmul_rn(0.5f, __fadd_rn(bc[1], bc[3]));
    const float dx = __fsub_rn(pcx, ccx), dy = __fsub_rn(pcy, ccy);
    const float dist = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const float wp = fmaxf(__fsub_rn(bp[2], bp[0]), 1.0f), hp = fmaxf(__fsub_rn(bp[3], bp[1]), 1.0f);
    const float sp = fmaxf(sqrtf(__fadd_rn(__fmul_rn(wp, wp), __fmul_rn(hp, hp))), 1.0f);
    o.center = __fdiv_rn(dist, sp);
    const float wc = fmaxf(__fsub_rn(bc[2], bc[0]), 1.0f), hc = fmaxf(__fsub_rn(bc[3], bc[1]), 1.0f);
    const float ratio = fmaxf(__fdiv_rn(__fmul_rn(wc, hc), __fmul_rn(wp, hp)), 1e-6f);
    o.scale = fabsf(logf(ratio));
    o.bbox = __fadd_rn(__fmul_rn(w.alpha, o.center), __fmul_rn(w.beta, o.scale));
    o.conf = fabsf(logf(__fdiv_rn(fmaxf(cc, w.conf_eps), fmaxf(cp, w.conf_eps))));
    o.total = __fadd_rn(__fadd_rn(__fmul_rn(w.w_app, c_app), __fmul_rn(w.w_bbox, o.bbox)),
                        __fmul_rn(w.w_conf, o.conf));
    return o;
}

}  // namespace cost
}  // namespace b200

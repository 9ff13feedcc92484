// 8-state constant-velocity Kalman filter, one thread per track (device functions).
//
// Semantics: model/utils/costTool/KalmanFilter.py:36-116 on top of filterpy 1.4.5
// predict/update (mainTracking.py:343,400).  State lives in float64 storage; `stage`
// (number of updates so far, saturating at 2) selects the arithmetic the reference's
// numpy code uses at that point of a track's life:
//   stage 0: x float32, P float32       (everything assigned as float32 at :57-99)
//   stage 1: x float32, P float64       (filterpy's float64 identity promotes P)
//   stage 2: x float64, P float64
// F = [I I; 0 I] and H = [I 0] are fixed (dt = 1), so F P F^T and H P H^T reduce to adds
// that round exactly like the dense numpy products (multiplying by exact 0/1).
#pragma once
#include "common.cuh"

namespace b200 {
namespace kf {

// bbox_xyxy_to_z, KalmanFilter.py:5-16: float64 math, float32 result.
__device__ __forceinline__ void box_to_z(const double* b, float* z) {
    const double w = fmax(1.0, b[2] - b[0]), h = fmax(1.0, b[3] - b[1]);
    z[0] = (float)(b[0] + 0.5 * w);
    z[1] = (float)(b[1] + 0.5 * h);
    z[2] = (float)(w / h);
    z[3] = (float)h;
}

// x_to_bbox_xyxy, KalmanFilter.py:19-33.
__device__ __forceinline__ void x_to_box(const double* x, double* b) {
    const double h = fmax(x[3], 1.0), a = fmax(x[2], 1e-3), w = fmax(a * h, 1.0);
    b[0] = x[0] - 0.5 * w;
    b[1] = x[1] - 0.5 * h;
    b[2] = x[0] + 0.5 * w;
    b[3] = x[1] + 0.5 * h;
}

// init_kf_from_bbox, KalmanFilter.py:74-88.
__device__ __forceinline__ void init_state(const double* box, double* x, double* P) {
    float z[4];
    box_to_z(box, z);
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = (double)z[i]; x[i + 4] = 0.0; }
#pragma unroll
    for (int i = 0; i < 64; ++i) P[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { P[i * 9] = 10.0; P[(i + 4) * 9] = 1000.0; }
}

// x <- F x ; P <- F P F^T + Q in the arithmetic type T (float or double), in place.
template <typename TX, typename TP>
__device__ __forceinline__ void predict_t(double* x, double* P, const float* q) {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = (double)((TX)x[i] + (TX)x[i + 4]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) P[i * 8 + j] = (double)((TP)P[i * 8 + j] + (TP)P[(i + 4) * 8 + j]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) P[i * 8 + j] = (double)((TP)P[i * 8 + j] + (TP)P[i * 8 + j + 4]);
#pragma unroll
    for (int i = 0; i < 8; ++i) P[i * 9] = (double)((TP)P[i * 9] + (TP)q[i]);
}

__device__ __forceinline__ void predict(double* x, double* P, int stage, const float* q) {
    if (stage == 0) predict_t<float, float>(x, P, q);
    else if (stage == 1) predict_t<float, double>(x, P, q);
    else predict_t<double, double>(x, P, q);
}

// inv(S) for a 4x4 matrix: LU with partial pivoting, then solves against the identity (what
// numpy.linalg.inv's gesv does).  Like LAPACK's getf2, pivots are inverted once and multiplied in: four
// divisions per inverse instead of twenty (a double division is a ~40-instruction dependent chain).
template <typename T>
__device__ __forceinline__ void inv4(const T* S, T* SI) {
    T a[4][4], b[4][4], rinv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[i][j] = S[i * 4 + j]; b[i][j] = (i == j) ? (T)1 : (T)0; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int p = j;
        T best = fabs(a[j][j]);
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            const T v = fabs(a[i][j]);
            if (v > best) { best = v; p = i; }
        }
#pragma unroll
        for (int i = j + 1; i < 4; ++i)
            if (p == i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    T t = a[j][k]; a[j][k] = a[i][k]; a[i][k] = t;
                    t = b[j][k]; b[j][k] = b[i][k]; b[i][k] = t;
                }
            }
        rinv[j] = (T)1 / a[j][j];
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            a[i][j] *= rinv[j];
#pragma unroll
            for (int k = j + 1; k < 4; ++k) a[i][k] -= a[i][j] * a[j][k];
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 1; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < i; ++k) b[i][c] -= a[i][k] * b[k][c];
#pragma unroll
        for (int i = 3; i >= 0; --i) {
#pragma unroll
            for (int k = i + 1; k < 4; ++k) b[i][c] -= a[i][k] * b[k][c];
            b[i][c] *= rinv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) SI[i * 4 + j] = b[i][j];
}

template <typename A, typename B> struct Promote { using type = double; };
template <> struct Promote<float, float> { using type = float; };

// filterpy KalmanFilter.update(z) with H = [I 0]; r = diag(R).  In place on x, P.
template <typename TX, typename TP>
__device__ __forceinline__ void update_t(double* x, double* P, const float* z, const float* r) {
    using TXN = typename Promote<TX, TP>::type;
    TX y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = (TX)z[i] - (TX)x[i];
    TP S[16], SI[16], K[32];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = (TP)P[i * 8 + j] + (i == j ? (TP)r[i] : (TP)0);
    inv4<TP>(S, SI);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            TP s = (TP)P[i * 8] * SI[j];
#pragma unroll
            for (int k = 1; k < 4; ++k) s = fma((TP)P[i * 8 + k], SI[k * 4 + j], s);
            K[i * 4 + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        TXN s = (TXN)K[i * 4] * (TXN)y[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) s = fma((TXN)K[i * 4 + k], (TXN)y[k], s);
        x[i] = (double)((TXN)(TX)x[i] + s);
    }
    // Joseph form in float64 (filterpy's identity is float64):
    //   P <- (I-KH) P (I-KH)^T + (K R) K^T,   I-KH = [[I-K1, 0], [-K2, I]].
    // Products are evaluated in place with k ascending; off-diagonal entries of I-KH are
    // exact negations of K, so only the four diagonal entries need a rounding of their own.
    double ikd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ikd[i] = 1.0 - (double)K[i * 4 + i];
#define B200_IK(i, k) ((i) == (k) ? ikd[(k)] : -(double)K[(i) * 4 + (k)])
    // A = (I-KH) P : rows 4..7 first (they read rows 0..3 before those are overwritten).
#pragma unroll
    for (int i = 4; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double s = B200_IK(i, 0) * P[j];
#pragma unroll
            for (int k = 1; k < 4; ++k) s = fma(B200_IK(i, k), P[k * 8 + j], s);
            P[i * 8 + j] = s + P[i * 8 + j];
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double p0 = P[j], p1 = P[8 + j], p2 = P[16 + j], p3 = P[24 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            P[i * 8 + j] = fma(B200_IK(i, 3), p3, fma(B200_IK(i, 2), p2, fma(B200_IK(i, 1), p1, B200_IK(i, 0) * p0)));
    }
    // B = A (I-KH)^T row by row, plus D = (K R) K^T evaluated in K's type.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double a0 = P[i * 8], a1 = P[i * 8 + 1], a2 = P[i * 8 + 2], a3 = P[i * 8 + 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double s = fma(a3, B200_IK(j, 3), fma(a2, B200_IK(j, 2), fma(a1, B200_IK(j, 1), a0 * B200_IK(j, 0))));
            TP d = (K[i * 4] * (TP)r[0]) * K[j * 4];
#pragma unroll
            for (int k = 1; k < 4; ++k) d = fma(K[i * 4 + k] * (TP)r[k], K[j * 4 + k], d);
            P[i * 8 + j] = (j < 4 ? s : s + P[i * 8 + j]) + (double)d;
        }
    }
#undef B200_IK
}

// Returns the new stage.
__device__ __forceinline__ int update(double* x, double* P, int stage, const float* z, const float* r) {
    if (stage == 0) update_t<float, float>(x, P, z, r);
    else if (stage == 1) update_t<float, double>(x, P, z, r);
    else update_t<double, double>(x, P, z, r);
    return stage < 2 ? stage + 1 : 2;
}

// Per-track part of gating_distance_maha (KalmanFilter.py:111-114): SI = inv(H P H^T + R + 1e-9 I).
struct Gate {
    double SI[16];
    double xs[4];
    int stage;
};

__device__ __forceinline__ void gate_prepare(const double* x, const double* P, int stage, const float* r,
                                             Gate* g) {
    g->stage = stage;
#pragma unroll
    for (int i = 0; i < 4; ++i) g->xs[i] = x[i];
    if (stage == 0) {
        float S[16], SI[16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float hph = (float)P[i * 8 + j] + (i == j ? r[i] : 0.0f);
                S[i * 4 + j] = hph + (i == j ? 1e-9f : 0.0f);
            }
        inv4<float>(S, SI);
#pragma unroll
        for (int i = 0; i < 16; ++i) g->SI[i] = (double)SI[i];
    } else {
        double S[16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double hph = P[i * 8 + j] + (i == j ? (double)r[i] : 0.0);
                S[i * 4 + j] = hph + (i == j ? (double)1e-9f : 0.0);
            }
        inv4<double>(S, g->SI);
    }
}

// Per-pair part: d2 = y^T SI y with y = z - H x (KalmanFilter.py:110-115).
__device__ __forceinline__ double gate_d2(const double* SI, const double* xs, int stage, const float* z) {
    if (stage == 0) {
        float y[4], t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = z[i] - (float)xs[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = y[0] * (float)SI[j];
#pragma unroll
            for (int i = 1; i < 4; ++i) s = fmaf(y[i], (float)SI[i * 4 + j], s);
            t[j] = s;
        }
        float d = t[0] * y[0];
#pragma unroll
        for (int j = 1; j < 4; ++j) d = fmaf(t[j], y[j], d);
        return (double)d;
    }
    double y[4], t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        y[i] = stage == 1 ? (double)(z[i] - (float)xs[i]) : (double)z[i] - xs[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double s = y[0] * SI[j];
#pragma unroll
        for (int i = 1; i < 4; ++i) s = fma(y[i], SI[i * 4 + j], s);
        t[j] = s;
    }
    double d = t[0] * y[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) d = fma(t[j], y[j], d);
    return d;
}

// Cooperative form of update_t: eight consecutive lanes of a warp share one track, lane i owning row i of
// P and of K.  Every entry is computed with the same operations in the same order as update_t (so the two
// forms agree bit for bit); what changes is who computes it, the code size (one row instead of 64 unrolled
// entries) and the register footprint.  sP (64 doubles) and sK (32 doubles) are this group's scratch in
// shared memory; `mask` is the set of lanes of the warp executing the call (all groups in it take the same
// template instantiation).  Returns the squared Mahalanobis distance of z to the UPDATED state when
// want_d2 (mainTracking.py:424), else 0.
template <typename TX, typename TP>
__device__ __forceinline__ double update_rows_t(double* gx, double* gP, const float* z, const float* r, double* sP,
                                                double* sK, int i, unsigned mask, int new_stage, bool want_d2) {
    using TXN = typename Promote<TX, TP>::type;
    double Pr[8];
    {
        const double2* p2 = reinterpret_cast<const double2*>(gP + i * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) { const double2 v = p2[k]; Pr[2 * k] = v.x; Pr[2 * k + 1] = v.y; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sP[i * 8 + j] = Pr[j];
    double xg[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xg[k] = gx[k];
    const double xi = gx[i];
    __syncwarp(mask);
    TP S[16], SI[16], Kr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) S[a * 4 + b] = (TP)sP[a * 8 + b] + (a == b ? (TP)r[a] : (TP)0);
    inv4<TP>(S, SI);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        TP acc = (TP)Pr[0] * SI[j];
#pragma unroll
        for (int k = 1; k < 4; ++k) acc = fma((TP)Pr[k], SI[k * 4 + j], acc);
        Kr[j] = acc;
    }
    TX y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = (TX)z[k] - (TX)xg[k];
    TXN sx = (TXN)Kr[0] * (TXN)y[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) sx = fma((TXN)Kr[k], (TXN)y[k], sx);
    const double xn = (double)((TXN)(TX)xi + sx);
#pragma unroll
    for (int j = 0; j < 4; ++j) sK[i * 4 + j] = (double)Kr[j];
    __syncwarp(mask);                       // P rows and K rows of the whole track are visible; x has been read
    gx[i] = xn;
    // A = (I-KH) P, row i
    double ik[4], A[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) ik[k] = (i == k) ? 1.0 - (double)Kr[k] : -(double)Kr[k];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double a = ik[0] * sP[j];
#pragma unroll
        for (int k = 1; k < 4; ++k) a = fma(ik[k], sP[k * 8 + j], a);
        A[j] = i >= 4 ? a + Pr[j] : a;
    }
    // B = A (I-KH)^T + (K R) K^T, row i
    double Pn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double ikj[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ikj[k] = (j == k) ? 1.0 - sK[j * 4 + k] : -sK[j * 4 + k];
        const double b = fma(A[3], ikj[3], fma(A[2], ikj[2], fma(A[1], ikj[1], A[0] * ikj[0])));
        TP dd = (Kr[0] * (TP)r[0]) * (TP)sK[j * 4];
#pragma unroll
        for (int k = 1; k < 4; ++k) dd = fma(Kr[k] * (TP)r[k], (TP)sK[j * 4 + k], dd);
        Pn[j] = (j < 4 ? b : b + A[j]) + (double)dd;
    }
    {
        double2* p2 = reinterpret_cast<double2*>(gP + i * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) p2[k] = make_double2(Pn[2 * k], Pn[2 * k + 1]);
    }
    if (!want_d2) return 0.0;
    __syncwarp(mask);                       // everyone is done reading the old P rows
    if (i < 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sP[i * 8 + j] = Pn[j];
        sK[i] = xn;
    }
    __syncwarp(mask);
    Gate g;
    double x4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x4[k] = sK[k];
    gate_prepare(x4, sP, new_stage, r, &g);
    return gate_d2(g.SI, g.xs, new_stage, z);
}

}  // namespace kf
}  // namespace b200

// Batched Kalman operators of the C ABI (thread per track).  The arithmetic is in kalman.cuh.
#include "kalman.cuh"

namespace b200 {
namespace {

constexpr int kKfThreads = 64;

__device__ __forceinline__ void load_state(const double* __restrict__ gx, const double* __restrict__ gP,
                                           int i, double* x, double* P) {
    const double2* px = reinterpret_cast<const double2*>(gx + (size_t)i * 8);
    const double2* pP = reinterpret_cast<const double2*>(gP + (size_t)i * 64);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const double2 v = px[k]; x[2 * k] = v.x; x[2 * k + 1] = v.y; }
#pragma unroll
    for (int k = 0; k < 32; ++k) { const double2 v = pP[k]; P[2 * k] = v.x; P[2 * k + 1] = v.y; }
}

__device__ __forceinline__ void store_state(double* __restrict__ gx, double* __restrict__ gP, int i,
                                            const double* x, const double* P) {
    double2* px = reinterpret_cast<double2*>(gx + (size_t)i * 8);
    double2* pP = reinterpret_cast<double2*>(gP + (size_t)i * 64);
#pragma unroll
    for (int k = 0; k < 4; ++k) px[k] = make_double2(x[2 * k], x[2 * k + 1]);
#pragma unroll
    for (int k = 0; k < 32; ++k) pP[k] = make_double2(P[2 * k], P[2 * k + 1]);
}

__global__ void __launch_bounds__(kKfThreads) kalman_init_kernel(const double* __restrict__ boxes, int M,
                                                                 double* gx, double* gP, uint8_t* stage) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double x[8], P[64], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = boxes[(size_t)i * 4 + k];
    kf::init_state(b, x, P);
    store_state(gx, gP, i, x, P);
    stage[i] = 0;
}

__global__ void __launch_bounds__(kKfThreads) kalman_predict_kernel(double* gx, double* gP,
                                                                    const uint8_t* __restrict__ stage, int M,
                                                                    const float* __restrict__ q_diag,
                                                                    double* __restrict__ pred_boxes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = q_diag[k];
    double x[8], P[64];
    load_state(gx, gP, i, x, P);
    kf::predict(x, P, stage[i], q);
    store_state(gx, gP, i, x, P);
    if (pred_boxes) {
        double b[4];
        kf::x_to_box(x, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) pred_boxes[(size_t)i * 4 + k] = b[k];
    }
}

__global__ void __launch_bounds__(kKfThreads) kalman_update_kernel(double* gx, double* gP, uint8_t* stage, int M,
                                                                   const int32_t* __restrict__ det_of_track,
                                                                   const double* __restrict__ boxes, int meas_is_z,
                                                                   const float* __restrict__ r_diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int j = det_of_track[i];
    if (j < 0) return;
    float r[4], z[4];
    double b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { r[k] = r_diag[k]; b[k] = boxes[(size_t)j * 4 + k]; }
    if (meas_is_z) {
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = (float)b[k];
    } else {
        kf::box_to_z(b, z);
    }
    double x[8], P[64];
    load_state(gx, gP, i, x, P);
    stage[i] = (uint8_t)kf::update(x, P, stage[i], z, r);
    store_state(gx, gP, i, x, P);
}

// One CTA per track row: the 4x4 inverse once, then threads sweep the boxes.
__global__ void __launch_bounds__(128) maha_gate_kernel(const double* __restrict__ gx, const double* __restrict__ gP,
                                                        const uint8_t* __restrict__ stage, int M,
                                                        const double* __restrict__ boxes, int N,
                                                        const float* __restrict__ r_diag, double maha_thr,
                                                        float inf_value, float* C, int ldc, double* d2, int ldd) {
    __shared__ kf::Gate g;
    const int i = blockIdx.x;
    if (threadIdx.x == 0) {
        float r[4];
        double P4[64];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = r_diag[k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) P4[a * 8 + b] = gP[(size_t)i * 64 + a * 8 + b];
        kf::gate_prepare(gx + (size_t)i * 8, P4, stage[i], r, &g);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double b[4];
        float z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] = boxes[(size_t)j * 4 + k];
        kf::box_to_z(b, z);
        const double d = kf::gate_d2(g.SI, g.xs, g.stage, z);
        if (d2) d2[(size_t)i * ldd + j] = d;
        if (C && d > maha_thr) C[(size_t)i * ldc + j] = inf_value;
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200_kalman_init(const double* boxes_xyxy, int M, double* x, double* P, uint8_t* stage,
                                void* stream) {
    B200_REQUIRE(M >= 0, "kalman_init: negative M");
    if (M == 0) return B200_OK;
    B200_REQUIRE(boxes_xyxy && x && P && stage, "kalman_init: null pointer");
    kalman_init_kernel<<<(M + kKfThreads - 1) / kKfThreads, kKfThreads, 0, as_stream(stream)>>>(boxes_xyxy, M, x, P, stage);
    return check_launch("kalman_init_kernel");
}

extern "C" int b200_kalman_predict(double* x, double* P, const uint8_t* stage, int M, const float* q_diag,
                                   double* pred_boxes_xyxy, void* stream) {
    B200_REQUIRE(M >= 0, "kalman_predict: negative M");
    if (M == 0) return B200_OK;
    B200_REQUIRE(x && P && stage && q_diag, "kalman_predict: null pointer");
    kalman_predict_kernel<<<(M + kKfThreads - 1) / kKfThreads, kKfThreads, 0, as_stream(stream)>>>(
        x, P, stage, M, q_diag, pred_boxes_xyxy);
    return check_launch("kalman_predict_kernel");
}

extern "C" int b200_kalman_update(double* x, double* P, uint8_t* stage, int M, const int32_t* det_of_track,
                                  const double* meas, int meas_is_z, const float* r_diag, void* stream) {
    B200_REQUIRE(M >= 0, "kalman_update: negative M");
    if (M == 0) return B200_OK;
    B200_REQUIRE(x && P && stage && det_of_track && meas && r_diag, "kalman_update: null pointer");
    kalman_update_kernel<<<(M + kKfThreads - 1) / kKfThreads, kKfThreads, 0, as_stream(stream)>>>(
        x, P, stage, M, det_of_track, meas, meas_is_z, r_diag);
    return check_launch("kalman_update_kernel");
}

extern "C" int b200_maha_gate(const double* x, const double* P, const uint8_t* stage, int M,
                              const double* boxes_xyxy, int N, const float* r_diag, double maha_thr,
                              float inf_value, float* C, int ldc, double* d2, int ldd, void* stream) {
    B200_REQUIRE(M >= 0 && N >= 0, "maha_gate: negative size");
    if (M == 0 || N == 0) return B200_OK;
    B200_REQUIRE(x && P && stage && boxes_xyxy && r_diag, "maha_gate: null pointer");
    B200_REQUIRE((!C || ldc >= N) && (!d2 || ldd >= N), "maha_gate: leading dimension < N");
    maha_gate_kernel<<<M, 128, 0, as_stream(stream)>>>(x, P, stage, M, boxes_xyxy, N, r_diag, maha_thr, inf_value, C,
                                                        ldc, d2, ldd);
    return check_launch("maha_gate_kernel");
}

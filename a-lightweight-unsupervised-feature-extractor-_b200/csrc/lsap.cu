// hungarian_assign on the GPU (C ABI entry point); the solver is in lsap.cuh.
#include "lsap.cuh"

namespace b200 {
namespace {

constexpr size_t kSmemMatrixBudget = 160 * 1024;

__global__ void transpose_kernel(const float* __restrict__ C, long long batch_stride, int M, int N, int ldc,
                                 float* __restrict__ T) {
    __shared__ float tile[32][33];
    const float* src = C + (size_t)blockIdx.z * batch_stride;
    float* dst = T + (size_t)blockIdx.z * M * N;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < M && j < N) ? src[(size_t)i * ldc + j] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, i = i0 + threadIdx.x;
        if (j < N && i < M) dst[(size_t)j * M + i] = tile[threadIdx.x][r];
    }
}

// One CTA per problem.  `Ct` is the [N][M] transpose (only read when M > N).
__global__ void lsap_kernel(const float* __restrict__ C, const float* __restrict__ Ct, long long batch_stride,
                            int M, int N, int ldc, double cost_max, int32_t* __restrict__ col_of_row,
                            uint8_t* __restrict__ matched, int32_t* __restrict__ status, int in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const bool tall = M > N;
    const int R = tall ? N : M, Cc = tall ? M : N;
    const float* orig = C + (size_t)b * batch_stride;
    const float* cost = tall ? Ct + (size_t)b * M * N : orig;
    const int ld = tall ? M : ldc;
    lsap::Work w = lsap::carve(smem_raw, R, Cc);
    float* stage = in_smem ? reinterpret_cast<float*>(smem_raw + lsap::work_bytes(R, Cc)) : nullptr;
    const int rc = lsap::solve_block(cost, R, Cc, ld, w, stage);
    if (tid == 0) status[b] = rc;
    int32_t* out_c = col_of_row + (size_t)b * M;
    uint8_t* out_m = matched + (size_t)b * M;
    for (int i = tid; i < M; i += nt) {
        int j = -1;
        if (rc == B200_OK) j = tall ? w.r4c[i] : w.c4r[i];
        out_c[i] = j;
        out_m[i] = (j >= 0 && (double)orig[(size_t)i * ldc + j] <= cost_max) ? 1 : 0;
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200_lsap_f32(const float* C, int batch, int64_t batch_stride, int M, int N, int ldc,
                             double cost_max, int32_t* col_of_row, uint8_t* matched, int32_t* status,
                             void* stream) {
    B200_REQUIRE(batch >= 0 && M >= 0 && N >= 0, "lsap: negative size");
    if (batch == 0) return B200_OK;
    B200_REQUIRE(status, "lsap: null status");
    cudaStream_t st = as_stream(stream);
    if (M == 0 || N == 0) {          // hung.py:19-25: nothing to solve; every row is unassigned
        B200_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * batch, st));
        if (M) {
            B200_REQUIRE(col_of_row && matched, "lsap: null output");
            B200_CUDA(cudaMemsetAsync(col_of_row, 0xff, sizeof(int32_t) * (size_t)batch * M, st));
            B200_CUDA(cudaMemsetAsync(matched, 0, (size_t)batch * M, st));
        }
        return B200_OK;
    }
    B200_REQUIRE(C && col_of_row && matched, "lsap: null pointer");
    B200_REQUIRE(ldc >= N, "lsap: ldc < N");
    const bool tall = M > N;
    const int R = tall ? N : M, Cc = tall ? M : N;
    const size_t wb = lsap::work_bytes(R, Cc);
    B200_REQUIRE(wb <= 200 * 1024, "lsap: problem %dx%d exceeds the shared-memory workspace", M, N);
    float* Ct = nullptr;
    if (tall) {
        B200_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&Ct), sizeof(float) * (size_t)batch * M * N, st));
        dim3 grid((N + 31) / 32, (M + 31) / 32, batch), block(32, 8);
        transpose_kernel<<<grid, block, 0, st>>>(C, batch_stride, M, N, ldc, Ct);
        int rc = check_launch("lsap transpose_kernel");
        if (rc != B200_OK) return rc;
    }
    const size_t mat = (size_t)R * Cc * sizeof(float);
    const int in_smem = (wb + mat <= kSmemMatrixBudget) ? 1 : 0;
    const size_t smem = wb + (in_smem ? mat : 0);
    const int nt = 256;
    static bool configured = false;
    if (!configured) {
        B200_CUDA(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured = true;
    }
    lsap_kernel<<<batch, nt, smem, st>>>(C, Ct, batch_stride, M, N, ldc, cost_max, col_of_row, matched, status,
                                         in_smem);
    int rc = check_launch("lsap_kernel");
    if (Ct) B200_CUDA(cudaFreeAsync(Ct, st));
    return rc;
}

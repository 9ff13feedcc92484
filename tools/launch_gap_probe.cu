// How long does a small kernel launched on a second stream wait before its first CTA runs while a large grid of
// small CTAs (ROI-Align-like: 64 threads, many registers, ~27 KB of shared memory, six per SM) fills the GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/launch_gap_probe tools/launch_gap_probe.cu
//   tools/build/launch_gap_probe
// Prints, per configuration (lifetime of the big grid's CTAs, size of the small kernel's CTAs, stream priorities),
// the delay between the small kernel becoming eligible (host launch after a spin kernel on its own stream has
// finished) and its first instruction, measured with %globaltimer.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int REGS_HINT>
__global__ void __launch_bounds__(64) big_kernel(int spin_ns, float* sink) {
    extern __shared__ float sm[];
    // occupy registers roughly like the ROI kernel (168 per thread) so that six CTAs fill an SM
    float acc[REGS_HINT];
#pragma unroll
    for (int i = 0; i < REGS_HINT; ++i) acc[i] = threadIdx.x * 0.001f + i;
    const unsigned long long t0 = gtimer();
    while (gtimer() - t0 < (unsigned long long)spin_ns) {
#pragma unroll
        for (int i = 0; i < REGS_HINT; ++i) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < REGS_HINT; ++i) s += acc[i];
    sm[threadIdx.x] = s;
    if (s == 12345.678f) sink[0] = sm[(threadIdx.x + 1) & 63];
}

__global__ void small_kernel(unsigned long long* first_start) {
    if (threadIdx.x == 0) atomicMin(first_start, gtimer());
}

__global__ void stamp_kernel(unsigned long long* t) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *t = gtimer();
}

int main() {
    float* sink;
    unsigned long long *d_first, *d_stamp;
    cudaMalloc(&sink, 4);
    cudaMalloc(&d_first, 8);
    cudaMalloc(&d_stamp, 8);
    auto big = big_kernel<110>;
    cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, 27 * 1024);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, big, 64, 27 * 1024);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, big);
    printf("big kernel: %d registers, %d CTAs per SM\n", fa.numRegs, per_sm);
    int lo, hi;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    for (int prio = 0; prio < 2; ++prio)
        for (int life_us : {5, 50})
            for (int small_threads : {32, 256, 1024})
                for (int small_ctas : {1, 64}) {
                    cudaStream_t sa, sb;
                    cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, lo);
                    cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, prio ? hi : lo);
                    double sum = 0;
                    const int reps = 20;
                    for (int r = 0; r < reps; ++r) {
                        unsigned long long big_val = ~0ull;
                        cudaMemcpy(d_first, &big_val, 8, cudaMemcpyHostToDevice);
                        const int total_us = 400, ctas = 148 * per_sm * (total_us / life_us);
                        big<<<ctas, 64, 27 * 1024, sa>>>(life_us * 1000, sink);
                        // on stream b: a stamp kernel (runs when it can), then the small kernel right behind it
                        stamp_kernel<<<1, 32, 0, sb>>>(d_stamp);
                        small_kernel<<<small_ctas, small_threads, 0, sb>>>(d_first);
                        cudaDeviceSynchronize();
                        unsigned long long a, b;
                        cudaMemcpy(&a, d_stamp, 8, cudaMemcpyDeviceToHost);
                        cudaMemcpy(&b, d_first, 8, cudaMemcpyDeviceToHost);
                        sum += (double)(b - a) * 1e-3;
                    }
                    printf("chain stream priority %-4s big-CTA lifetime %3d us  small kernel %4d threads x %2d CTAs : "
                           "starts %6.1f us after its predecessor on the same stream\n",
                           prio ? "high" : "same", life_us, small_threads, small_ctas, sum / reps);
                    cudaStreamDestroy(sa);
                    cudaStreamDestroy(sb);
                }
    return 0;
}

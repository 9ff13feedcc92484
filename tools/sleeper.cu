// Experiment helper: a kernel that only waits (global timer + nanosleep) with a chosen grid / block / dynamic shared
// memory, to see what a co-resident kernel of a given FOOTPRINT costs ROI Align on the other stream (tools/sleeper_probe.py).
#include <cuda_runtime.h>
__global__ void sleeper_kernel(long long ns) {
    extern __shared__ char smem[];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (ns < 0) smem[threadIdx.x] = 0;
    do {
        __nanosleep(500);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while ((long long)(t - t0) < ns);
}
extern "C" int sleeper_carveout(int percent) {      // -1 = driver default, 100 = maximum shared memory
    return (int)cudaFuncSetAttribute(sleeper_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, percent);
}
extern "C" int sleeper_launch(int grid, int block, int smem, long long ns, void* stream) {
    static int configured = 0;
    if (!configured) {
        cudaFuncSetAttribute(sleeper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured = 1;
    }
    sleeper_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(ns);
    return (int)cudaGetLastError();
}

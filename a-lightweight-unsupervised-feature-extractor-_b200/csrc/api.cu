// Library-wide pieces of the C ABI: error string, version, launch counter.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace b200 {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Stream-ordered scratch (per-ROI records and weight tables of ROI Align's two-kernel path, operand images of the
// tensor-core cost kernel) comes from a PRIVATE pool per device that keeps freed blocks cached; the application's
// default pool and its release threshold are not touched.
namespace {
struct ScratchPool {
    std::mutex mu;
    cudaMemPool_t pool = nullptr;
};
ScratchPool g_scratch[64];
}  // namespace

int scratch_pool(cudaMemPool_t* out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    ScratchPool& sp = g_scratch[dev];
    std::lock_guard<std::mutex> lock(sp.mu);
    if (!sp.pool) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        B200_CUDA(cudaMemPoolCreate(&pool, &props));
        unsigned long long keep = ~0ull;
        B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        sp.pool = pool;
    }
    *out = sp.pool;
    return B200_OK;
}
}  // namespace b200

extern "C" int b200_version(void) { return 100; }
extern "C" const char* b200_last_error(void) { return b200::g_err; }
extern "C" int64_t b200_launch_count(void) { return b200::g_launches.load(); }


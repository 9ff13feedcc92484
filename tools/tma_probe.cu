// Stand-alone probe of the TMA pieces of roi_align_tma_kernel (csrc/roi_align.cu): a 4-D box load of an NCHW
// footprint + a bulk copy on one mbarrier, checked element by element.  Variants select the box shape and how the
// tensor map reaches the kernel, one process per variant so a faulting variant does not hide the others:
//     tools/build/tma_probe <variant>      (nvcc -gencode arch=compute_100a,code=sm_100a -o tools/build/tma_probe tools/tma_probe.cu)
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box_4d(unsigned dst, const CUtensorMap* tm, int x, int y, int c, int b, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(c), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_box_2d(unsigned dst, const CUtensorMap* tm, int x, int y, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mode bit 0: tensor load, bit 1: bulk table load, bit 2: tensor map from global memory, bit 3: prefetch.tensormap
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tmap, const CUtensorMap* gmap, const float* tabs, float* outV,
                             float* outT, int x0, int y0, int c0, int b0, int box_bytes, int mode) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned sbar = sbase + 16384;
    const int lane = threadIdx.x & 31;
    const CUtensorMap* tm = (mode & 4) ? gmap : &tmap;
    if (lane == 0) {
        mbar_init(sbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        if (mode & 8) asm volatile("prefetch.tensormap [%0];\n" ::"l"(tm) : "memory");
    }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(sbar, (unsigned)(((mode & 1) ? box_bytes : 0) + ((mode & 2) ? 768 : 0)));
        if (mode & 2) bulk_load(sbase + 8192, tabs, 768, sbar);
        if (mode & 1) {
            if (mode & 16) tma_box_2d(sbase, tm, x0, y0, sbar);
            else tma_box_4d(sbase, tm, x0, y0, c0, b0, sbar);
        }
    }
    mbar_wait(sbar, 0);
    for (int i = lane; i < box_bytes / 4; i += 32) outV[i] = reinterpret_cast<const float*>(smem)[i];
    for (int i = lane; i < 192; i += 32) outT[i] = reinterpret_cast<const float*>(smem + 8192)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    // variant: box rows / box x / mode
    int BX = 4, BY = 5, mode = 3;
    switch (variant) {
        case 0: mode = 2; break;                 // bulk copy + mbarrier only
        case 1: mode = 1; break;                 // tensor load only
        case 2: mode = 3; break;                 // both (what the ROI kernel does)
        case 3: mode = 1 | 4; break;             // tensor map in global memory
        case 4: mode = 1; BX = 8; break;         // 32-byte inner box
        case 5: mode = 1; BY = 4; break;         // even row count
        case 6: mode = 3 | 8; break;             // with prefetch.tensormap
        case 7: mode = 1; BX = 16; BY = 8; break;
        case 8: mode = 1 | 16; BX = 16; BY = 8; break;   // 2-D map (W, H*C*B), in-bounds
        case 9: mode = 1; break;                 // 4-D, in-bounds coordinates
        case 10: mode = 1; break;                // 4-D, encode function from dlopen(libcuda.so.1)
        case 11: mode = 1 | 16; break;           // 2-D, 16-byte inner box
        default: break;
    }
    const int B = 3, C = 70, H = 34, W = 60;
    std::vector<float> h((size_t)B * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.25f;
    float *d = nullptr, *tabs = nullptr, *outV = nullptr, *outT = nullptr;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> ht(192);
    for (int i = 0; i < 192; ++i) ht[i] = 1000.0f + i;
    CK(cudaMalloc(&tabs, 768));
    CK(cudaMemcpy(tabs, ht.data(), 768, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&outV, 16384));
    CK(cudaMalloc(&outT, 768));
    CK(cudaMemset(outV, 0, 16384));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("FAIL no cuTensorMapEncodeTiled\n"); return 2; }
    if (variant == 10) {
        void* h = dlopen("libcuda.so.1", RTLD_NOW);
        void* f2 = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr;
        printf("dlsym cuTensorMapEncodeTiled = %p (entry point %p)\n", f2, fn);
        if (f2) fn = f2;
    }
    CUtensorMap tmap;
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)BX, (cuuint32_t)BY, 32u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    memset(&tmap, 0, sizeof(tmap));
    CUresult r;
    if (mode & 16) {
        const cuuint64_t d2[2] = {(cuuint64_t)W, (cuuint64_t)H * C * B};
        const cuuint64_t s2[1] = {(cuuint64_t)W * 4};
        r = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, d2, s2, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else
    r = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    {
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&tmap);
        printf("tensor map:");
        for (int i = 0; i < 16; ++i) printf(" %016llx", w[i]);
        printf("\n");
    }
    if (r != CUDA_SUCCESS) { printf("FAIL encode rc=%d\n", (int)r); return 2; }
    CUtensorMap* gmap = nullptr;
    CK(cudaMalloc(&gmap, sizeof(CUtensorMap)));
    CK(cudaMemcpy(gmap, &tmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    int x0 = 58, y0 = 31, c0 = 64, b0 = 1;                       // crosses the right / bottom edge and the channel count
    if (variant >= 8) { x0 = 8; y0 = 8; c0 = 0; b0 = 0; }
    if (argc > 5) { x0 = atoi(argv[2]); y0 = atoi(argv[3]); c0 = atoi(argv[4]); b0 = atoi(argv[5]); }
    printf("coords x %d y %d c %d b %d\n", x0, y0, c0, b0);
    const int box_bytes = (mode & 16) ? BX * BY * 4 : BX * BY * 32 * 4;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    probe_kernel<<<1, 32, 32768>>>(tmap, gmap, tabs, outV, outT, x0, y0, c0, b0, box_bytes, mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> v(box_bytes / 4), t(192);
    CK(cudaMemcpy(v.data(), outV, box_bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t.data(), outT, 768, cudaMemcpyDeviceToHost));
    int bad = 0;
    if (mode & 16) {
        for (int y = 0; y < BY; ++y)
            for (int x = 0; x < BX; ++x)
                if (v[y * BX + x] != h[(size_t)(y0 + y) * W + x0 + x]) ++bad;
    } else if (mode & 1)
        for (int c = 0; c < 32; ++c)
            for (int y = 0; y < BY; ++y)
                for (int x = 0; x < BX; ++x) {
                    const int gc = c0 + c, gy = y0 + y, gx = x0 + x;
                    const float want = (gc < C && gy < H && gx < W) ? h[(((size_t)b0 * C + gc) * H + gy) * W + gx] : 0.0f;
                    if (v[(c * BY + y) * BX + x] != want) ++bad;
                }
    if (mode & 2)
        for (int i = 0; i < 192; ++i)
            if (t[i] != ht[i]) ++bad;
    printf("variant %d (box %dx%dx32, mode %d): %s (%d mismatches)\n", variant, BX, BY, mode, bad ? "WRONG" : "ok", bad);
    return bad ? 1 : 0;
}

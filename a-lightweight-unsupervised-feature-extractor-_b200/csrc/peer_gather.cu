// Result-table exchange between the GPUs of one node over NVLink peer memory.
//
// The path shards across streams only, so the one piece of inter-GPU traffic is optional: the consumer of
// tracking.py:329 wants every stream's matches in one place.  A NCCL all-gather does that, but its kernel sits on an
// SM and spins until the slowest rank arrives, and with a one-CTA footprint it moves the tables at a fraction of a
// link's bandwidth (8 GPUs, 16 frames per call: 15 MB in ~0.6 ms, all of it exposed at the end of a short run).  Here every
// rank owns a receive ring that its peers map through CUDA IPC; a rank PUSHES its tables into the same slot of every
// peer's ring with plain 16-byte stores over NVLink / NVSwitch and then raises a per-(slot, source) flag.  Nobody waits
// for anybody while producing.  A consumer that wants sequence number q waits (one small kernel) until the flags of q
// from all ranks are up, copies the slot out and acknowledges it to the producers, which only look at the
// acknowledgement when they are about to overwrite that slot n_slots sequence numbers later.
#include <string.h>

#include <new>

#include "common.cuh"

namespace b200 {
namespace peer {

constexpr int kMaxWorld = 16;
constexpr int kCtasPerPeer = 4;
constexpr int kThreads = 256;

struct Ring {                       // addresses of ONE rank's ring as seen from this process
    char* data;                     // [n_slots][world][bytes_per_rank]
    unsigned long long* flag;       // [n_slots][world]  sequence number + 1 of the tables in that cell
    unsigned long long* ack;        // [world]           ack[r] = 1 + the last sequence number rank r has consumed from ME
};

struct Params {
    Ring ring[kMaxWorld];           // ring[p] = rank p's ring (ring[rank] is local memory)
    int rank, world, n_slots;
    long long bytes_per_rank;
    unsigned int* arrive;           // local: [world] CTAs of the current push that have finished their chunk for peer p
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid = world * kCtasPerPeer: CTA (p, c) stores chunk c of `src` into rank p's ring, cell (slot, my rank).
__global__ void __launch_bounds__(kThreads) push_kernel(Params P, const char* __restrict__ src, long long bytes,
                                                        long long seq) {
    const int p = blockIdx.x / kCtasPerPeer, c = blockIdx.x % kCtasPerPeer;
    const int slot = (int)(seq % P.n_slots);
    const Ring& local = P.ring[P.rank];
    // The cell is free once rank p has consumed what I put there n_slots pushes ago (p tells me through MY ack array).
    if (seq >= P.n_slots && threadIdx.x == 0) {
        const unsigned long long need = (unsigned long long)(seq - P.n_slots + 1);
        while (ld_acquire_sys(local.ack + p) < need) __nanosleep(200);
    }
    __syncthreads();
    char* dst = P.ring[p].data + ((size_t)slot * P.world + P.rank) * (size_t)P.bytes_per_rank;
    const long long n16 = bytes / 16;
    const long long per = (n16 + kCtasPerPeer - 1) / kCtasPerPeer;
    const long long lo = c * per, hi = lo + per < n16 ? lo + per : n16;
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (long long i = lo + threadIdx.x; i < hi; i += kThreads) d4[i] = s4[i];
    if (c == kCtasPerPeer - 1)
        for (long long i = n16 * 16 + threadIdx.x; i < bytes; i += kThreads) dst[i] = src[i];
    __threadfence_system();                       // my stores are visible system-wide before the arrival is counted
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(P.arrive + p, 1u);
        if (prev == kCtasPerPeer - 1) {           // last chunk for this peer: raise the flag in ITS memory
            P.arrive[p] = 0;
            __threadfence_system();
            st_release_sys(P.ring[p].flag + (size_t)slot * P.world + P.rank, (unsigned long long)(seq + 1));
        }
    }
}

// Waits until every rank's tables of sequence number `seq` are in my ring, copies [world][bytes] to `dst` and tells the
// producers that the slot may be overwritten.  grid = a few CTAs; each waits on its own.
__global__ void __launch_bounds__(kThreads) collect_kernel(Params P, char* __restrict__ dst, long long bytes, long long seq,
                                                           unsigned int* done_ctas) {
    const int slot = (int)(seq % P.n_slots);
    const Ring& local = P.ring[P.rank];
    if (threadIdx.x < P.world) {
        const unsigned long long* f = local.flag + (size_t)slot * P.world + threadIdx.x;
        while (ld_acquire_sys(f) < (unsigned long long)(seq + 1)) __nanosleep(200);
    }
    __syncthreads();
    const long long n16 = bytes / 16;                                  // callers keep bytes a multiple of 16
    const long long total = n16 * P.world;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long r = i / n16, k = i - r * n16;
        const int4* s4 = reinterpret_cast<const int4*>(local.data + ((size_t)slot * P.world + r) * (size_t)P.bytes_per_rank);
        reinterpret_cast<int4*>(dst + (size_t)r * bytes)[k] = s4[k];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done_ctas, 1u);
        if (prev == gridDim.x - 1) {                                   // every CTA has read the slot: acknowledge
            *done_ctas = 0;
            for (int p = 0; p < P.world; ++p) st_release_sys(P.ring[p].ack + P.rank, (unsigned long long)(seq + 1));
        }
    }
}

}  // namespace peer
}  // namespace b200

using namespace b200;

struct b200_peer_gather {
    peer::Params P;
    void* local = nullptr;          // my ring: data | flags | acks | arrive | done counter
    void* mapped[peer::kMaxWorld] = {};
    size_t total_bytes = 0;
    bool connected = false;
};

namespace {
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

void carve(char* base, int world, int n_slots, long long bytes_per_rank, peer::Ring* r) {
    const size_t data = align256((size_t)n_slots * world * (size_t)bytes_per_rank);
    const size_t flags = align256(sizeof(unsigned long long) * (size_t)n_slots * world);
    r->data = base;
    r->flag = reinterpret_cast<unsigned long long*>(base + data);
    r->ack = reinterpret_cast<unsigned long long*>(base + data + flags);
}
}  // namespace

extern "C" int b200_peer_gather_create(b200_peer_gather** out, int rank, int world, int64_t bytes_per_rank, int n_slots) {
    B200_REQUIRE(out, "peer_gather_create: null pointer");
    B200_REQUIRE(world >= 1 && world <= peer::kMaxWorld && rank >= 0 && rank < world, "peer_gather_create: rank %d of %d", rank, world);
    B200_REQUIRE(bytes_per_rank > 0 && bytes_per_rank % 16 == 0 && n_slots >= 2 && n_slots <= 64,
                 "peer_gather_create: bytes_per_rank must be a positive multiple of 16 and 2 <= n_slots <= 64");
    b200_peer_gather* g = new (std::nothrow) b200_peer_gather();
    B200_REQUIRE(g, "peer_gather_create: out of host memory");
    const size_t data = align256((size_t)n_slots * world * (size_t)bytes_per_rank);
    const size_t flags = align256(sizeof(unsigned long long) * (size_t)n_slots * world);
    const size_t acks = align256(sizeof(unsigned long long) * world);
    const size_t ctrs = align256(sizeof(unsigned) * (world + 1));
    g->total_bytes = data + flags + acks + ctrs;
    cudaError_t e = cudaMalloc(&g->local, g->total_bytes);
    if (e == cudaSuccess) e = cudaMemset(g->local, 0, g->total_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(g->local);
        delete g;
        return fail(B200_ECUDA, "peer_gather_create: %s", cudaGetErrorString(e));
    }
    memset(&g->P, 0, sizeof(g->P));
    g->P.rank = rank; g->P.world = world; g->P.n_slots = n_slots; g->P.bytes_per_rank = bytes_per_rank;
    char* base = static_cast<char*>(g->local);
    carve(base, world, n_slots, bytes_per_rank, &g->P.ring[rank]);
    g->P.arrive = reinterpret_cast<unsigned*>(base + data + flags + acks);
    g->connected = world == 1;
    *out = g;
    return B200_OK;
}

extern "C" int b200_peer_gather_handle(b200_peer_gather* g, void* handle_host) {
    B200_REQUIRE(g && handle_host, "peer_gather_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    B200_CUDA(cudaIpcGetMemHandle(&h, g->local));
    memcpy(handle_host, &h, sizeof(h));
    return B200_OK;
}

extern "C" int b200_peer_gather_connect(b200_peer_gather* g, const void* handles_host) {
    B200_REQUIRE(g && handles_host, "peer_gather_connect: null pointer");
    const peer::Params& P = g->P;
    for (int p = 0; p < P.world; ++p) {
        if (p == P.rank || g->mapped[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles_host) + (size_t)p * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_ECUDA, "peer_gather_connect: cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
        }
        g->mapped[p] = ptr;
        carve(static_cast<char*>(ptr), P.world, P.n_slots, P.bytes_per_rank, &g->P.ring[p]);
    }
    g->connected = true;
    return B200_OK;
}

extern "C" int b200_peer_gather_push(b200_peer_gather* g, const void* src, int64_t bytes, int64_t seq, void* stream) {
    B200_REQUIRE(g && src, "peer_gather_push: null pointer");
    B200_REQUIRE(g->connected, "peer_gather_push: not connected");
    B200_REQUIRE(bytes > 0 && bytes <= g->P.bytes_per_rank && bytes % 16 == 0 && seq >= 0 &&
                 (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                 "peer_gather_push: bytes must be a multiple of 16 within the slot, src 16-byte aligned");
    peer::push_kernel<<<g->P.world * peer::kCtasPerPeer, peer::kThreads, 0, as_stream(stream)>>>(
        g->P, static_cast<const char*>(src), bytes, seq);
    return check_launch("peer push_kernel");
}

extern "C" int b200_peer_gather_collect(b200_peer_gather* g, void* dst, int64_t bytes, int64_t seq, void* stream) {
    B200_REQUIRE(g && dst, "peer_gather_collect: null pointer");
    B200_REQUIRE(g->connected, "peer_gather_collect: not connected");
    B200_REQUIRE(bytes > 0 && bytes <= g->P.bytes_per_rank && bytes % 16 == 0 && seq >= 0 &&
                 (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "peer_gather_collect: bytes must be a multiple of 16 within the slot, dst 16-byte aligned");
    peer::collect_kernel<<<8, peer::kThreads, 0, as_stream(stream)>>>(g->P, static_cast<char*>(dst), bytes, seq,
                                                                       g->P.arrive + g->P.world);
    return check_launch("peer collect_kernel");
}

extern "C" void b200_peer_gather_destroy(b200_peer_gather* g) {
    if (!g) return;
    cudaDeviceSynchronize();
    for (int p = 0; p < peer::kMaxWorld; ++p)
        if (g->mapped[p]) cudaIpcCloseMemHandle(g->mapped[p]);
    cudaFree(g->local);
    delete g;
}

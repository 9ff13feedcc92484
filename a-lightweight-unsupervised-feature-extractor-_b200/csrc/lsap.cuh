// Rectangular linear-sum assignment on the GPU: shortest augmenting paths with float64 duals,
// one problem per group of cooperating threads (device functions).
//
// Follows the algorithm behind scipy.optimize.linear_sum_assignment, which the reference calls
// at model/utils/costTool/hung.py:28 (Crouse 2016, no initialisation phase), step for step:
// the same descending initial column list, the same swap-removal of scanned columns, the same
// tie rule (first minimum in list order unless a later minimum is an unassigned column) and the
// same expression ((minv + c) - u[i]) - v[j] for the reduced cost, so the result -- including
// which optimum is returned under ties -- is the one scipy returns.  What is parallel is the
// scan of the remaining columns inside one Dijkstra step: each thread owns list positions
// tid, tid+nt, ...; the (value, tie rule) argmin is an associative reduction over them.
#pragma once
#include "common.cuh"

namespace b200 {
namespace lsap {

constexpr int kMaxThreads = 256;

struct Work {              // all arrays in shared memory, sized for R rows / Cc columns
    double* u;             // [R]
    double* v;             // [Cc]
    double* dist;          // [Cc]
    int* pred;             // [Cc]
    int* c4r;              // [R]
    int* r4c;              // [Cc]
    int* todo;             // [Cc]
    int* seen_rows;        // [R]   rows scanned in the current search (SR)
    int* seen_cols;        // [Cc]  columns scanned in the current search (SC)
    double* red_val;       // [kMaxThreads / 32] cross-warp reduction scratch
    int* red_it;           // [kMaxThreads / 32]
    int* red_un;           // [kMaxThreads / 32]
    int* flag;             // [1]
};

__host__ __device__ inline size_t work_bytes(int R, int Cc) {
    const size_t d = (size_t)R + 2 * (size_t)Cc + kMaxThreads / 32;
    const size_t i = 2 * (size_t)R + 4 * (size_t)Cc + 2 * (kMaxThreads / 32) + 2;
    return d * 8 + ((i * 4 + 7) & ~(size_t)7);
}

__device__ __forceinline__ Work carve(unsigned char* base, int R, int Cc) {
    Work w;
    double* d = reinterpret_cast<double*>(base);
    w.u = d; d += R;
    w.v = d; d += Cc;
    w.dist = d; d += Cc;
    w.red_val = d; d += kMaxThreads / 32;
    int* p = reinterpret_cast<int*>(d);
    w.pred = p; p += Cc;
    w.c4r = p; p += R;
    w.r4c = p; p += Cc;
    w.todo = p; p += Cc;
    w.seen_rows = p; p += R;
    w.seen_cols = p; p += Cc;
    w.red_it = p; p += kMaxThreads / 32;
    w.red_un = p; p += kMaxThreads / 32;
    w.flag = p;
    return w;
}

// Barrier among the `nt` cooperating threads (named barrier 1 when they span several warps).
__device__ __forceinline__ void group_sync(int nt) {
    if (nt <= 32) __syncwarp();
    else asm volatile("bar.sync 1, %0;" ::"r"(nt) : "memory");
}

struct Cand {
    double val;
    int it;      // position in the todo list
    int un;      // 1 if that column is unassigned
};

// scipy's selection rule as an associative combine.
__device__ __forceinline__ Cand better(const Cand& a, const Cand& b) {
    if (a.it < 0) return b;
    if (b.it < 0) return a;
    if (a.val < b.val) return a;
    if (b.val < a.val) return b;
    if (a.un != b.un) return a.un ? a : b;
    if (a.un) return a.it > b.it ? a : b;      // last unassigned minimum in list order
    return a.it < b.it ? a : b;                // else the first minimum
}

__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand d;
        d.val = __shfl_xor_sync(0xffffffffu, c.val, o);
        d.it = __shfl_xor_sync(0xffffffffu, c.it, o);
        d.un = __shfl_xor_sync(0xffffffffu, c.un, o);
        c = better(c, d);
    }
    return c;
}

// Solves the R x Cc problem (R <= Cc) whose row i is cost[i * ld + 0..Cc).  `tid` in [0, nt),
// nt a multiple of 32 (<= kMaxThreads); all nt threads must call.  On return w.c4r / w.r4c hold
// the assignment.  Returns B200_OK, B200_ENUMERIC or B200_EINFEASIBLE (same on every thread).
__device__ inline int solve(const float* cost, int R, int Cc, int ld, const Work& w, int tid, int nt) {
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    // NaN / -inf check (scipy: "matrix contains invalid numeric entries").
    if (tid == 0) *w.flag = 0;
    group_sync(nt);
    int bad = 0;
    for (int i = 0; i < R; ++i)
        for (int j = tid; j < Cc; j += nt) {
            const float c = cost[(size_t)i * ld + j];
            if (c != c || c == -__int_as_float(0x7f800000)) bad = 1;
        }
    if (bad) *w.flag = 1;
    for (int i = tid; i < R; i += nt) { w.u[i] = 0.0; w.c4r[i] = -1; }
    for (int j = tid; j < Cc; j += nt) { w.v[j] = 0.0; w.r4c[j] = -1; w.pred[j] = -1; }
    group_sync(nt);
    if (*w.flag) return B200_ENUMERIC;

    for (int cur = 0; cur < R; ++cur) {
        for (int t = tid; t < Cc; t += nt) {
            w.todo[t] = Cc - 1 - t;
            w.dist[t] = kInf;
        }
        group_sync(nt);
        int i = cur, n_todo = Cc, n_sr = 0, n_sc = 0, sink = -1;
        double minv = 0.0;
        while (sink < 0) {
            if (tid == 0) w.seen_rows[n_sr] = i;
            ++n_sr;
            const double ui = w.u[i];
            const float* crow = cost + (size_t)i * ld;
            Cand best;
            best.val = kInf; best.it = -1; best.un = 0;
            for (int it = tid; it < n_todo; it += nt) {
                const int j = w.todo[it];
                const double r = ((minv + (double)crow[j]) - ui) - w.v[j];
                double d = w.dist[j];
                if (r < d) {
                    d = r;
                    w.dist[j] = r;
                    w.pred[j] = i;
                }
                const int un = w.r4c[j] < 0;
                if (best.it < 0 || d < best.val || (d == best.val && un)) {
                    best.val = d; best.it = it; best.un = un;
                }
            }
            best = warp_best(best);
            if (nt > 32) {
                const int wid = tid >> 5;
                if ((tid & 31) == 0) { w.red_val[wid] = best.val; w.red_it[wid] = best.it; w.red_un[wid] = best.un; }
                group_sync(nt);
                Cand c;
                c.val = kInf; c.it = -1; c.un = 0;
                if ((tid & 31) < (nt >> 5)) {
                    c.val = w.red_val[tid & 31]; c.it = w.red_it[tid & 31]; c.un = w.red_un[tid & 31];
                }
                best = warp_best(c);
            }
            minv = best.val;
            if (best.it < 0 || minv == kInf) return B200_EINFEASIBLE;      // uniform across the group
            const int j = w.todo[best.it];
            const int rj = w.r4c[j];
            group_sync(nt);                    // everyone has read todo[best.it] / red_* before they change
            if (tid == 0) {
                w.seen_cols[n_sc] = j;
                w.todo[best.it] = w.todo[n_todo - 1];
            }
            ++n_sc;
            --n_todo;
            if (rj < 0) sink = j; else i = rj;
            group_sync(nt);
        }
        // dual update (scipy: u[cur] += minv; u[i] += minv - dist[c4r[i]]; v[j] -= minv - dist[j])
        for (int s = tid; s < n_sr; s += nt) {
            const int r = w.seen_rows[s];
            if (r == cur) w.u[r] += minv;
            else w.u[r] += minv - w.dist[w.c4r[r]];
        }
        for (int s = tid; s < n_sc; s += nt) {
            const int j = w.seen_cols[s];
            w.v[j] -= minv - w.dist[j];
        }
        group_sync(nt);
        if (tid == 0) {                       // augment along the predecessor chain
            int j = sink;
            for (;;) {
                const int r = w.pred[j];
                w.r4c[j] = r;
                const int prev = w.c4r[r];
                w.c4r[r] = j;
                j = prev;
                if (r == cur) break;
            }
        }
        group_sync(nt);
    }
    return B200_OK;
}

}  // namespace lsap
}  // namespace b200

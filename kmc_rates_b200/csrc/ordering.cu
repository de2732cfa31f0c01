// Level-structure sweeps of the nested-dissection ordering on the GPU (see ordering.h, symbolic.h: SweepAccel).
//
// A cluster of eight CTAs walks the whole subgraph level by level.  Every CTA keeps the visited set as a bitmap in its
// shared memory (n / 8 bytes), so the ~19 of 20 adjacency entries that lead to a visited node cost one shared-memory
// read.  The sweep reports the serial breadth-first queue order exactly: the serial sweep appends a node at its first
// sighting, i.e. at the smallest (position of the parent in its level, position in the parent's row) among the entries
// that lead to it, so every entry (p, j) of the frontier that leads to an unvisited node v claims it with
// atomicMin(key[v], p << 18 | j), the claim that stands is the winner, and winners are placed at (exclusive scan of the
// winners per parent) + (rank inside the parent's row).  See sw_run for the phases.
// The kernel gives up (status 2) on structures thinner than ~24 nodes per level: a level costs it ~10 us whatever its
// size, a host core ~2.6 ns per adjacency entry.
#include "ordering.h"

#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <vector>

namespace ngt {
namespace {

constexpr int SW_THREADS = 512;        // per CTA: 128 registers each, a whole row of <= 32 entries in flight
constexpr int SW_CL = 8;               // CTAs of the cluster that walks one subgraph
constexpr int SW_FCAP = 16384;         // largest level (its nodes' row extents live in global memory)
constexpr int SW_LP = SW_FCAP / SW_CL; // largest slice of a level one CTA takes
constexpr int SW_LT = 4096;            // most nodes one CTA's slice may win for the next level
constexpr int SW_ROW = 32;             // row entries a thread keeps in flight
constexpr int SW_JBITS = 18;           // claim = position of the parent in its level << 18 | position in the parent's row
constexpr int SW_SMEM_LIMIT = 227 * 1024;

struct SweepOut {
    int count;    // nodes reached by the last sweep
    int nlev;     // its number of levels
    int status;   // 0 ok, 1 not connected, 2 levels too thin, 3 a level exceeded the arrays, 4 internal
    int root;     // start node of the last sweep
#ifdef NGT_SW_CLOCKS
    long long t[12];  // cycles of thread 0 of CTA 0 per piece of a level (see SW_T), [11] = levels
#endif
};
#ifdef NGT_SW_CLOCKS
#define SW_T(i) do { if (threadIdx.x == 0 && c == 0) { const long long c_ = clock64(); a.out->t[i] += c_ - t_last; t_last = c_; } } while (0)
#else
#define SW_T(i) do { } while (0)
#endif

struct SweepArgs {
    int n;
    const int *ptr;
    const int *adj;
    int start;
    unsigned *key;
    int *lvl;
    int *visit;
    SweepOut *out;
    int nw;       // bitmap words
    int *hr0;     // 2 x SW_FCAP: row start of the nodes of the current / next level (by level parity)
    int *hdeg;    // 2 x SW_FCAP: their degrees
};

struct SweepSmem {
    unsigned *bits;        // every CTA keeps the whole visited set
    int *lr0, *ldeg;       // row start and degree of this CTA's slice of the level
    int *off;              // winners per parent of the slice, then their exclusive prefix sums
    int *touched;          // the nodes this CTA's slice won, in the order they were found
    unsigned *tkey;        // per touched node: parent (index in the slice) << 19 | rank among the parent's winners
    int *wsum;             // 33
    int *ctr;              // 0: touched nodes
    int *tot;              // 2 x SW_CL x 2: per level parity and CTA: nodes won, overflow flag (written by the peers)
    unsigned long long *best;
};

__device__ __forceinline__ bool sw_seen(const unsigned *bits, int v) { return (bits[v >> 5] >> (v & 31)) & 1u; }
// all threads of all CTAs of the cluster; orders their global-memory accesses (release / acquire at cluster scope)
__device__ __forceinline__ void sw_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the same shared-memory variable of CTA `cta` of the cluster
__device__ __forceinline__ void sw_peer_store(int *mine, int cta, int value) {
    const unsigned local = (unsigned)__cvta_generic_to_shared(mine);
    unsigned remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(value) : "memory");
}
__device__ __forceinline__ int sw_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return (int)r;
}

// exclusive prefix sums of x[0..n) in place; returns the total (uniform over the CTA)
__device__ int sw_scan(int *x, int n, int *wsum) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + SW_THREADS - 1) / SW_THREADS;
    const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
    int s = 0;
    for (int i = i0; i < i1; ++i) s += x[i];
    int inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < SW_THREADS / 32 ? wsum[lane] : 0;
        int winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += y;
        }
        wsum[lane] = winc - w;
        if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    int run = wsum[warp] + inc - s;
    for (int i = i0; i < i1; ++i) {
        const int c = x[i];
        x[i] = run;
        run += c;
    }
    const int total = wsum[32];
    __syncthreads();
    return total;
}

// up to 32 consecutive entries of one row, all in flight: claims ...
__device__ __forceinline__ void sw_claim_chunk(const SweepArgs &a, const SweepSmem &s, const int *src, int nvalid,
                                               unsigned kbase) {
    int v[SW_ROW];
#pragma unroll
    for (int t = 0; t < SW_ROW; ++t) v[t] = t < nvalid ? __ldg(src + t) : -1;
#pragma unroll
    for (int t = 0; t < SW_ROW; ++t)
        if (v[t] >= 0 && !sw_seen(s.bits, v[t])) atomicMin(a.key + v[t], kbase + (unsigned)t);
}
// ... and winners (cnt: winners of this parent so far = rank of the next one)
__device__ __forceinline__ void sw_win_chunk(const SweepArgs &a, const SweepSmem &s, const int *src, int nvalid,
                                             unsigned kbase, int q, int &cnt) {
    int v[SW_ROW];
    unsigned k[SW_ROW];
#pragma unroll
    for (int t = 0; t < SW_ROW; ++t) {
        v[t] = t < nvalid ? __ldg(src + t) : -1;
        if (v[t] >= 0 && sw_seen(s.bits, v[t])) v[t] = -1;
    }
#pragma unroll
    for (int t = 0; t < SW_ROW; ++t) k[t] = v[t] >= 0 ? __ldcg(a.key + v[t]) : 0u;
#pragma unroll
    for (int t = 0; t < SW_ROW; ++t)
        if (v[t] >= 0 && k[t] == kbase + (unsigned)t) {
            const int i = atomicAdd(s.ctr, 1);
            if (i < SW_LT) {
                s.touched[i] = v[t];
                s.tkey[i] = ((unsigned)q << 19) | (unsigned)cnt;
            }
            ++cnt;
        }
}

// One sweep from `start` by the whole cluster; every return value is uniform over the cluster.  A level's nodes (the
// "parents") are split into SW_CL contiguous slices, one per CTA, one thread per parent, up to 32 row entries in flight:
//   A  an entry (p, j) that leads to an unvisited node v claims it: atomicMin(key[v], p << 18 | j), no return value;
//   B  the same walk again: the entry whose claim stands is the winner.  A thread meets its parent's winners in row
//      order, so the rank of a winner inside its parent is a running count; winners go on the CTA's `touched` list;
//   C  exclusive scan of the winners per parent inside the slice; the slices' totals are exchanged through distributed
//      shared memory, which gives every CTA the offset of its slice in the next level;
//   D  one thread per touched node: position = slice offset + parent offset + rank; level and row extent to global
//      memory;
//   E  every CTA sets the bits of the whole new level in its own copy of the visited set.
// Three cluster barriers per level.  What bounds a level on one SM is the ~1.3 cycles its memory pipe needs per
// scattered 32-bit atomic or load (about seven claims and seven re-reads per new node: 20 us per level of ~700 nodes,
// no better than a host core); a warp or eight lanes per parent cost 5-10x the instructions, returning 64-bit atomics
// twice the pipe time.  Eight SMs share that work.
__device__ int sw_run(const SweepArgs &a, const SweepSmem &s, int start, int *count, int *nlev, int *last_base,
                      int *last_size) {
    const int tid = threadIdx.x;
    const int c = sw_cluster_rank();
#ifdef NGT_SW_CLOCKS
    long long t_last = clock64();
#endif
    for (int i = tid; i < a.nw; i += SW_THREADS) s.bits[i] = 0u;
    for (int i = c * SW_THREADS + tid; i < a.n; i += SW_CL * SW_THREADS) a.key[i] = ~0u;
    if (tid == 0) {
        if (c == 0) {
            a.visit[0] = start;
            a.lvl[start] = 0;
            const int r0 = __ldg(a.ptr + start);
            a.hr0[0] = r0;
            a.hdeg[0] = __ldg(a.ptr + start + 1) - r0;
        }
        s.ctr[0] = 0;
    }
    __syncthreads();
    if (tid == 0) s.bits[start >> 5] |= 1u << (start & 31);
    sw_cluster_sync();
    int base = 0, fsz = 1, level = 0, visited = 1;
    SW_T(0);
    for (;;) {
        const int par = level & 1;
        const int *chr0 = a.hr0 + par * SW_FCAP, *chdeg = a.hdeg + par * SW_FCAP;
        int *nhr0 = a.hr0 + (par ^ 1) * SW_FCAP, *nhdeg = a.hdeg + (par ^ 1) * SW_FCAP;
        const int lo = (int)((long long)fsz * c / SW_CL), hi = (int)((long long)fsz * (c + 1) / SW_CL);
        const int np = hi - lo;
        // ---- A: claims
        for (int q = tid; q < np; q += SW_THREADS) {
            const int r0 = __ldcg(chr0 + lo + q), dg = __ldcg(chdeg + lo + q);
            s.lr0[q] = r0;
            s.ldeg[q] = dg;
            const unsigned kp = (unsigned)(lo + q) << SW_JBITS;
            for (int j = 0; j < dg; j += SW_ROW)
                sw_claim_chunk(a, s, a.adj + r0 + j, min(dg - j, SW_ROW), kp + (unsigned)j);
        }
        SW_T(3);
        sw_cluster_sync();
        SW_T(4);
        // ---- B: winners, in row order per parent
        for (int q = tid; q < np; q += SW_THREADS) {
            const int r0 = s.lr0[q], dg = s.ldeg[q];
            const unsigned kp = (unsigned)(lo + q) << SW_JBITS;
            int cnt = 0;
            for (int j = 0; j < dg; j += SW_ROW)
                sw_win_chunk(a, s, a.adj + r0 + j, min(dg - j, SW_ROW), kp + (unsigned)j, q, cnt);
            s.off[q] = cnt;
        }
        __syncthreads();
        SW_T(5);
        // ---- C: offsets of the parents inside the slice, of the slice inside the next level
        const int nt = s.ctr[0];
        const int slice_total = sw_scan(s.off, np, s.wsum);
        if (tid < SW_CL) {                                   // every CTA learns every slice's total (distributed shared memory)
            sw_peer_store(s.tot + (par * SW_CL + c) * 2, tid, nt);
            sw_peer_store(s.tot + (par * SW_CL + c) * 2 + 1, tid, (nt > SW_LT || slice_total != nt) ? 1 : 0);
        }
        if (tid == 0) s.ctr[0] = 0;                          // for the next level's B
        SW_T(6);
        sw_cluster_sync();
        int slice_base = 0, total = 0, bad = 0;
#pragma unroll
        for (int cc = 0; cc < SW_CL; ++cc) {
            const int t = s.tot[(par * SW_CL + cc) * 2];
            bad |= s.tot[(par * SW_CL + cc) * 2 + 1];
            if (cc < c) slice_base += t;
            total += t;
        }
        SW_T(7);
        if (bad || total > SW_FCAP) return 3;                // a level (or what one slice of it wins) exceeds the arrays
        // ---- D: the next level, in queue order
        const int nbase = base + fsz;
        for (int i = tid; i < nt; i += SW_THREADS) {
            const unsigned tk = s.tkey[i];
            const int slot = slice_base + s.off[tk >> 19] + (int)(tk & ((1u << 19) - 1u));
            const int v = s.touched[i];
            const int q0 = __ldg(a.ptr + v), q1 = __ldg(a.ptr + v + 1);
            a.visit[nbase + slot] = v;
            a.lvl[v] = level + 1;
            nhr0[slot] = q0;
            nhdeg[slot] = q1 - q0;
        }
        SW_T(8);
        sw_cluster_sync();
        SW_T(9);
        // ---- E: the whole new level into this CTA's visited set
        for (int i = tid; i < total; i += SW_THREADS) {
            const int v = __ldcg(a.visit + nbase + i);
            atomicOr(s.bits + (v >> 5), 1u << (v & 31));
        }
        __syncthreads();
        SW_T(10);
#ifdef NGT_SW_CLOCKS
        if (tid == 0 && c == 0) a.out->t[11] += 1;
#endif
        if (total == 0) break;
        base = nbase;
        fsz = total;
        ++level;
        visited += total;
        if (level >= 64 && visited < 24 * level) return 2;
    }
    *count = visited;
    *nlev = level + 1;
    *last_base = base;
    *last_size = fsz;
    return visited == a.n ? 0 : 1;
}

__global__ void __cluster_dims__(SW_CL, 1, 1) __launch_bounds__(SW_THREADS, 1) nd_sweep_kernel(SweepArgs a) {
    extern __shared__ __align__(16) unsigned char sw_raw[];
    SweepSmem s;
    s.best = reinterpret_cast<unsigned long long *>(sw_raw);
    s.ctr = reinterpret_cast<int *>(sw_raw + 8);
    s.wsum = reinterpret_cast<int *>(sw_raw + 32);
    s.tot = s.wsum + 36;
    s.lr0 = s.tot + 4 * SW_CL;
    s.ldeg = s.lr0 + SW_LP;
    s.off = s.ldeg + SW_LP;
    s.touched = s.off + SW_LP;
    s.tkey = reinterpret_cast<unsigned *>(s.touched + SW_LT);
    s.bits = s.tkey + SW_LT;
    const int tid = threadIdx.x;
    const int c = sw_cluster_rank();
#ifdef NGT_SW_CLOCKS
    if (tid == 0 && c == 0) for (int i = 0; i < 12; ++i) a.out->t[i] = 0;
#endif
    int count = 0, nlev = 0, lb = 0, ls = 0;
    int root = a.start < 0 ? 0 : a.start;
    int status = sw_run(a, s, root, &count, &nlev, &lb, &ls);
    if (a.start < 0 && status == 0) {
        // minimum-degree node of the last level, the latest visited among equals (George & Liu's pseudo-peripheral
        // step); every CTA works it out for itself
        if (tid == 0) *s.best = ~0ull;
        __syncthreads();
        for (int i = tid; i < ls; i += SW_THREADS) {
            const int pos = lb + i;
            const int v = __ldcg(a.visit + pos);
            const int deg = __ldg(a.ptr + v + 1) - __ldg(a.ptr + v);
            atomicMin(s.best, ((unsigned long long)(unsigned)deg << 32) | (0xffffffffu - (unsigned)pos));
        }
        __syncthreads();
        const int best = __ldcg(a.visit + (int)(0xffffffffu - (unsigned)(*s.best & 0xffffffffull)));
        sw_cluster_sync();                                   // nobody reads the first sweep's order any more
        if (best != 0) {
            root = best;
            status = sw_run(a, s, root, &count, &nlev, &lb, &ls);
        }
    }
    if (tid == 0 && c == 0) {
        a.out->count = count;
        a.out->nlev = nlev;
        a.out->status = status;
        a.out->root = root;
    }
}

std::atomic<long long> g_sw_h2d{0}, g_sw_d2h{0};

struct SweepCtx {
    cudaStream_t st = nullptr;
    int *d_ptr = nullptr, *d_adj = nullptr, *d_lvl = nullptr, *d_visit = nullptr;
    unsigned *d_key = nullptr;
    SweepOut *d_out = nullptr, *h_out = nullptr;
    int *d_hdr = nullptr;   // row extents of two levels
    size_t cap_n = 0, cap_adj = 0;
    void free_buffers() {
        cudaFree(d_ptr); cudaFree(d_adj); cudaFree(d_lvl); cudaFree(d_visit); cudaFree(d_key);
        d_ptr = d_adj = d_lvl = d_visit = nullptr;
        d_key = nullptr;
        cap_n = cap_adj = 0;
    }
    bool reserve(size_t n, size_t nadj) {
        if (n > cap_n) {
            cudaFree(d_ptr); cudaFree(d_lvl); cudaFree(d_visit); cudaFree(d_key);
            d_ptr = d_lvl = d_visit = nullptr;
            d_key = nullptr;
            cap_n = 0;
            const size_t c = n + n / 8 + 64;
            if (cudaMalloc(&d_ptr, (c + 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&d_lvl, c * sizeof(int)) != cudaSuccess ||
                cudaMalloc(&d_visit, c * sizeof(int)) != cudaSuccess ||
                cudaMalloc(&d_key, c * sizeof(unsigned)) != cudaSuccess)
                return false;
            cap_n = c;
        }
        if (nadj > cap_adj) {
            cudaFree(d_adj);
            d_adj = nullptr;
            cap_adj = 0;
            const size_t c = nadj + nadj / 8 + 64;
            if (cudaMalloc(&d_adj, c * sizeof(int)) != cudaSuccess) return false;
            cap_adj = c;
        }
        return true;
    }
};

class GpuSweeps : public SweepAccel {
public:
    explicit GpuSweeps(int device) : device_(device) {
        if (const char *e = std::getenv("NGT_B200_GPU_SWEEP_MIN")) min_nodes = std::max(1, std::atoi(e));
    }
    ~GpuSweeps() override = default;   // contexts live as long as the process (the CUDA context may be gone at exit)

    int sweep(int32_t n, const int32_t *ptr, const int32_t *adj, int64_t nadj, int32_t start, int32_t *lvl,
              int32_t *visit, int32_t *nlev) override {
        if (broken_.load(std::memory_order_relaxed) || n < 2 || nadj <= 0 || nadj > 0x7fffffff) return -1;
        const int nw = (n + 31) / 32;
        if (max_degree(n, ptr) >= (1 << SW_JBITS) - 1) return -1;   // claims hold 18 bits of row position
        const size_t smem = 32 + (36 + 4 * SW_CL) * 4 + (3 * (size_t)SW_LP + 2 * (size_t)SW_LT) * 4 + (size_t)nw * 4;
        if (smem > (size_t)SW_SMEM_LIMIT) return -1;      // the visited set does not fit in shared memory
        SweepCtx *c = acquire();
        if (!c) return 0;
        bool ok = false;
        do {
            if (cudaSetDevice(device_) != cudaSuccess) break;
            if (!c->st) {
                if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) break;
                if (cudaMalloc(&c->d_out, sizeof(SweepOut)) != cudaSuccess) break;
                if (cudaMalloc(&c->d_hdr, 4 * (size_t)SW_FCAP * sizeof(int)) != cudaSuccess) break;
                if (cudaMallocHost(&c->h_out, sizeof(SweepOut)) != cudaSuccess) break;
            }
            if (!attr_set_.load(std::memory_order_acquire)) {
                if (cudaFuncSetAttribute(nd_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SMEM_LIMIT) != cudaSuccess) break;
                attr_set_.store(true, std::memory_order_release);
            }
            if (!c->reserve((size_t)n, (size_t)nadj)) break;
            static const bool trace = std::getenv("NGT_B200_TIMING") != nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            if (cudaMemcpyAsync(c->d_ptr, ptr, ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice, c->st) != cudaSuccess) break;
            if (cudaMemcpyAsync(c->d_adj, adj, (size_t)nadj * sizeof(int), cudaMemcpyHostToDevice, c->st) != cudaSuccess) break;
            g_sw_h2d += (long long)(((size_t)n + 1 + (size_t)nadj) * sizeof(int));
            uploads.fetch_add(1, std::memory_order_release);   // the engine holds its own bulk copies back until here
            auto t1 = t0;
            if (trace) { cudaStreamSynchronize(c->st); t1 = std::chrono::steady_clock::now(); }
            SweepArgs a;
            a.n = n;
            a.ptr = c->d_ptr;
            a.adj = c->d_adj;
            a.start = start;
            a.key = c->d_key;
            a.lvl = c->d_lvl;
            a.visit = c->d_visit;
            a.out = c->d_out;
            a.nw = nw;
            a.hr0 = c->d_hdr;
            a.hdeg = c->d_hdr + 2 * SW_FCAP;
            nd_sweep_kernel<<<SW_CL, SW_THREADS, smem, c->st>>>(a);
            if (cudaGetLastError() != cudaSuccess) break;
            if (cudaMemcpyAsync(c->h_out, c->d_out, sizeof(SweepOut), cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
            if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
            g_sw_d2h += (long long)sizeof(SweepOut);
            const auto t2 = std::chrono::steady_clock::now();
            if (c->h_out->status != 0) {                     // a property of the graph, not an error: the host sweeps
                release(c);
                return -1;
            }
            if (cudaMemcpyAsync(lvl, c->d_lvl, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
            if (cudaMemcpyAsync(visit, c->d_visit, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
            if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
            g_sw_d2h += (long long)(2 * (size_t)n * sizeof(int));
            *nlev = c->h_out->nlev;
            ok = true;
            if (trace) {
                auto ms = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) {
                    return std::chrono::duration<double, std::milli>(y - x).count();
                };
                std::fprintf(stderr, "[ngt timing]     device sweep n %d start %d: %d levels, upload %.2f ms, kernel %.2f ms, download %.2f ms\n",
                             n, start, c->h_out->nlev, ms(t0, t1), ms(t1, t2), ms(t2, std::chrono::steady_clock::now()));
#ifdef NGT_SW_CLOCKS
                const long long *t = c->h_out->t;
                const double L = (double)std::max(1LL, t[11]);
                std::fprintf(stderr, "[ngt timing]       cycles per level over %lld levels: claims %.0f  barrier %.0f  winners %.0f  "
                             "scan+publish %.0f  barrier %.0f  place %.0f  barrier %.0f  visited set %.0f   (init %lld)\n",
                             t[11], t[3] / L, t[4] / L, t[5] / L, t[6] / L, t[7] / L, t[8] / L, t[9] / L, t[10] / L, t[0]);
#endif
            }
        } while (false);
        if (!ok) {
            // a CUDA error here must not take the analysis down: the host sweeps, and the engine's own device
            // initialisation reports what is wrong with the device
            cudaGetLastError();
            broken_.store(true, std::memory_order_relaxed);
        }
        release(c);
        return ok ? 1 : -1;
    }

    void release_buffers() {
        std::lock_guard<std::mutex> lk(mu_);
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(device_);
        for (auto &c : all_)
            if (!c->busy) c->ctx.free_buffers();
        cudaSetDevice(cur);
    }
    int device() const { return device_; }

private:
    static int max_degree(int32_t n, const int32_t *ptr) {
        int m = 0;
        for (int32_t i = 0; i < n; ++i) m = std::max(m, ptr[i + 1] - ptr[i]);
        return m;
    }
    struct Slot {
        SweepCtx ctx;
        bool busy = false;
    };
    SweepCtx *acquire() {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto &c : all_)
            if (!c->busy) { c->busy = true; return &c->ctx; }
        if (all_.size() >= 4) return nullptr;
        all_.emplace_back(new Slot());
        all_.back()->busy = true;
        return &all_.back()->ctx;
    }
    void release(SweepCtx *x) {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto &c : all_)
            if (&c->ctx == x) c->busy = false;
    }
    int device_;
    std::mutex mu_;
    std::vector<std::unique_ptr<Slot> > all_;
    std::atomic<bool> broken_{false};
    std::atomic<bool> attr_set_{false};
};

std::mutex g_sw_mu;
std::vector<GpuSweeps *> g_sw_all;   // one per device, never destroyed (see ~GpuSweeps)

}  // namespace

SweepAccel *gpu_sweeps(int device) {
    static const bool off = std::getenv("NGT_B200_NO_GPU_SWEEPS") != nullptr;
    if (off || device < 0) return nullptr;
    std::lock_guard<std::mutex> lk(g_sw_mu);
    for (GpuSweeps *g : g_sw_all)
        if (g->device() == device) return g;
    g_sw_all.push_back(new GpuSweeps(device));
    return g_sw_all.back();
}

void release_gpu_sweeps() {
    std::lock_guard<std::mutex> lk(g_sw_mu);
    for (GpuSweeps *g : g_sw_all) g->release_buffers();
}

long long gpu_sweeps_h2d_bytes() { return g_sw_h2d.load(); }
long long gpu_sweeps_d2h_bytes() { return g_sw_d2h.load(); }

}  // namespace ngt

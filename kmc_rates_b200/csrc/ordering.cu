// Level-structure sweeps of the nested-dissection ordering on the GPU (see ordering.h, symbolic.h: SweepAccel).
//
// One CTA of 1024 threads walks the whole subgraph level by level.  The visited set is a bitmap in shared memory
// (n / 8 bytes), so the ~19 of 20 adjacency entries that lead to a visited node cost one shared-memory read.  The
// sweep reports the serial breadth-first queue order exactly: the serial sweep appends a node at its first sighting,
// i.e. at the smallest (position of the parent in its level, position in the parent's row) among the entries that
// lead to it, so every entry (p, j) of the frontier that leads to an unvisited node v claims it with
// atomicMin(key[v], p << 32 | j), the claim that stands is the winner, and winners are placed at (exclusive scan of
// the winners per parent) + (rank inside the parent's row).  See sw_run for the phases.
// A level costs a few dependent L2 round trips plus ~20 warp instructions per frontier node, against ~2.6 ns per
// adjacency entry on a host core; the kernel gives up (status 2) on structures thinner than ~24 nodes per level.
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

constexpr int SW_THREADS = 1024;
constexpr int SW_FCAP_MAX = 7168;      // largest level the per-parent arrays are sized for (28 bytes per node, < 2^13)
constexpr int SW_SMEM_LIMIT = 227 * 1024;

struct SweepOut {
    int count;    // nodes reached by the last sweep
    int nlev;     // its number of levels
    int status;   // 0 ok, 1 not connected, 2 levels too thin, 3 a level exceeded the shared-memory arrays, 4 internal
    int root;     // start node of the last sweep
#ifdef NGT_SW_CLOCKS
    long long t[6];   // cycles of thread 0 in: init, A, B1, scan, B3, levels
#endif
};
#ifdef NGT_SW_CLOCKS
#define SW_T(i) do { if (threadIdx.x == 0) { const long long c_ = clock64(); a.out->t[i] += c_ - t_last; t_last = c_; } } while (0)
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
    int fcap;     // capacity of the per-parent counters
};

struct SweepSmem {
    unsigned *bits;
    int *r0, *deg;         // row start and degree of the current / next level's nodes (2 x fcap each, double buffered)
    int *off;              // winners per parent, then their exclusive prefix sums
    int *touched;          // the nodes of the next level, in the order their winners were found
    unsigned *tkey;        // per touched node: parent << SW_JBITS | rank among the parent's winners
    int *wsum;             // 33
    int *ctr;              // 0: touched nodes
    unsigned long long *best;
};
constexpr int SW_JBITS = 19;           // claim = parent position << 19 | position in the parent's row

__device__ __forceinline__ bool sw_seen(const unsigned *bits, int v) { return (bits[v >> 5] >> (v & 31)) & 1u; }

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
        const int w = wsum[lane];
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

// One sweep from `start`; every return value is uniform over the CTA.  One thread per node of the level (the
// "parent") walks that node's row, four entries in flight:
//   A  an entry (p, j) that leads to an unvisited node v claims it: atomicMin(key[v], p << 19 | j), no return value;
//   B  the same walk again: the entry whose claim stands is the winner.  A thread meets its parent's winners in row
//      order, so the rank of a winner inside its parent is a running count; winners go on the `touched` list;
//   C  exclusive scan of the winners per parent: every parent's offset in the next level;
//   D  one thread per touched node: position = offset of its parent + rank; level, visited bit, row extent.
// What bounds a level is the ~1.3 cycles the SM's memory pipe needs per scattered 32-bit atomic or load (about seven
// claims and seven re-reads per new node), then the issue slots of 32 warps; a warp or eight lanes per parent cost
// 5-10x the instructions, returning (64-bit) atomics twice the pipe time.
__device__ int sw_run(const SweepArgs &a, const SweepSmem &s, int start, int *count, int *nlev, int *last_base,
                      int *last_size) {
    const int tid = threadIdx.x;
#ifdef NGT_SW_CLOCKS
    long long t_last = clock64();
#endif
    for (int i = tid; i < a.nw; i += SW_THREADS) s.bits[i] = 0u;
    for (int i = tid; i < a.n; i += SW_THREADS) a.key[i] = ~0u;
    if (tid == 0) {
        a.visit[0] = start;
        a.lvl[start] = 0;
        const int r0 = __ldg(a.ptr + start);
        s.r0[0] = r0;
        s.deg[0] = __ldg(a.ptr + start + 1) - r0;
        s.ctr[0] = 0;
    }
    __syncthreads();
    if (tid == 0) s.bits[start >> 5] |= 1u << (start & 31);
    __syncthreads();
    int base = 0, fsz = 1, level = 0, visited = 1;
    SW_T(0);
    for (;;) {
        const int flip = (level & 1) ? a.fcap : 0;
        const int *cr0 = s.r0 + flip, *cdeg = s.deg + flip;
        int *nr0 = s.r0 + (a.fcap - flip), *ndeg = s.deg + (a.fcap - flip);
        // ---- A: claims
        for (int p = tid; p < fsz; p += SW_THREADS) {
            const int dg = cdeg[p];
            const int *row = a.adj + cr0[p];
            const unsigned kp = (unsigned)p << SW_JBITS;
            for (int j = 0; j < dg; j += 4) {
                int v[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) v[t] = j + t < dg ? __ldg(row + j + t) : -1;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (v[t] >= 0 && !sw_seen(s.bits, v[t])) atomicMin(a.key + v[t], kp | (unsigned)(j + t));
            }
        }
        __syncthreads();
        SW_T(1);
        // ---- B: winners, in row order per parent
        for (int p = tid; p < fsz; p += SW_THREADS) {
            const int dg = cdeg[p];
            const int *row = a.adj + cr0[p];
            const unsigned kp = (unsigned)p << SW_JBITS;
            int cnt = 0;
            for (int j = 0; j < dg; j += 4) {
                int v[4];
                unsigned k[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    v[t] = j + t < dg ? __ldg(row + j + t) : -1;
                    if (v[t] >= 0 && sw_seen(s.bits, v[t])) v[t] = -1;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) k[t] = v[t] >= 0 ? __ldcg(a.key + v[t]) : 0u;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (v[t] >= 0 && k[t] == (kp | (unsigned)(j + t))) {
                        const int i = atomicAdd(s.ctr, 1);
                        if (i < a.fcap) {
                            s.touched[i] = v[t];
                            s.tkey[i] = kp | (unsigned)cnt;
                        }
                        ++cnt;
                    }
            }
            s.off[p] = cnt;
        }
        __syncthreads();
        SW_T(2);
        const int nt = s.ctr[0];
        if (nt > a.fcap) return 3;                           // the next level exceeds the shared-memory arrays
        // ---- C: offsets of the parents in the next level
        const int total = sw_scan(s.off, fsz, s.wsum);
        if (tid == 0) s.ctr[0] = 0;                          // for the next level's B (two barriers away)
        SW_T(3);
        if (total != nt) return 4;                           // cannot happen: every claimed node has one winner
        // ---- D: the next level, in queue order
        const int nbase = base + fsz;
        for (int i = tid; i < nt; i += SW_THREADS) {
            const unsigned tk = s.tkey[i];
            const int slot = s.off[tk >> SW_JBITS] + (int)(tk & ((1u << SW_JBITS) - 1u));
            const int v = s.touched[i];
            const int q0 = __ldg(a.ptr + v), q1 = __ldg(a.ptr + v + 1);
            a.visit[nbase + slot] = v;
            a.lvl[v] = level + 1;
            nr0[slot] = q0;
            ndeg[slot] = q1 - q0;
            atomicOr(s.bits + (v >> 5), 1u << (v & 31));
        }
        __syncthreads();
        SW_T(4);
#ifdef NGT_SW_CLOCKS
        if (tid == 0) a.out->t[5] += 1;
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

__global__ void __launch_bounds__(SW_THREADS, 1) nd_sweep_kernel(SweepArgs a) {
    extern __shared__ __align__(16) unsigned char sw_raw[];
    SweepSmem s;
    s.best = reinterpret_cast<unsigned long long *>(sw_raw);
    s.ctr = reinterpret_cast<int *>(sw_raw + 8);
    s.wsum = reinterpret_cast<int *>(sw_raw + 32);
    s.bits = reinterpret_cast<unsigned *>(s.wsum + 36);
    s.r0 = reinterpret_cast<int *>(s.bits + a.nw);
    s.deg = s.r0 + 2 * a.fcap;
    s.off = s.deg + 2 * a.fcap;
    s.touched = s.off + a.fcap;
    s.tkey = reinterpret_cast<unsigned *>(s.touched + a.fcap);
    const int tid = threadIdx.x;
#ifdef NGT_SW_CLOCKS
    if (tid == 0) for (int i = 0; i < 6; ++i) a.out->t[i] = 0;
#endif
    int count = 0, nlev = 0, lb = 0, ls = 0;
    int root = a.start < 0 ? 0 : a.start;
    int status = sw_run(a, s, root, &count, &nlev, &lb, &ls);
    if (a.start < 0 && status == 0) {
        // minimum-degree node of the last level, the latest visited among equals (George & Liu's pseudo-peripheral step)
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
        __syncthreads();
        if (best != 0) {
            root = best;
            status = sw_run(a, s, root, &count, &nlev, &lb, &ls);
        }
    }
    if (tid == 0) {
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

    bool sweep(int32_t n, const int32_t *ptr, const int32_t *adj, int64_t nadj, int32_t start, int32_t *lvl,
               int32_t *visit, int32_t *nlev) override {
        if (broken_.load(std::memory_order_relaxed) || n < 2 || nadj <= 0 || nadj > 0x7fffffff) return false;
        const int nw = (n + 31) / 32;
        if (max_degree(n, ptr) >= (1 << 19)) return false;   // claims hold 19 bits of row position
        const long long fixed = 32 + 36 * 4 + (long long)nw * 4;   // scalars, scan scratch, bitmap
        long long fcap = (SW_SMEM_LIMIT - fixed) / 28;
        if (fcap < 1024) return false;                       // the bitmap does not fit beside the per-parent arrays
        if (fcap > SW_FCAP_MAX) fcap = SW_FCAP_MAX;
        // levels of a d-dimensional structure hold ~n^(1-1/d) nodes; what shared memory the arrays leave becomes L1,
        // which then holds the rows of a level between the four-entry steps of their threads
        const long long want = std::max(2048LL, (long long)(10.0 * std::sqrt((double)n)));
        if (!big_levels_.load(std::memory_order_relaxed) && fcap > want) fcap = want;
        const size_t smem = (size_t)(fixed + fcap * 28);
        SweepCtx *c = acquire();
        if (!c) return false;
        bool ok = false;
        do {
            if (cudaSetDevice(device_) != cudaSuccess) break;
            if (!c->st) {
                if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) break;
                if (cudaMalloc(&c->d_out, sizeof(SweepOut)) != cudaSuccess) break;
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
            a.fcap = (int)fcap;
            nd_sweep_kernel<<<1, SW_THREADS, smem, c->st>>>(a);
            if (cudaGetLastError() != cudaSuccess) break;
            if (cudaMemcpyAsync(c->h_out, c->d_out, sizeof(SweepOut), cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
            if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
            g_sw_d2h += (long long)sizeof(SweepOut);
            const auto t2 = std::chrono::steady_clock::now();
            if (c->h_out->status != 0) {                     // a property of the graph, not an error: the host sweeps
                if (c->h_out->status == 3) big_levels_.store(true, std::memory_order_relaxed);   // next time: all the room
                release(c);
                return false;
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
                std::fprintf(stderr, "[ngt timing]       cycles: init %lld  claims %lld  winners %lld  scan %lld  append %lld  over %lld levels\n",
                             t[0], t[1], t[2], t[3], t[4], t[5]);
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
        return ok;
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
    std::atomic<bool> broken_{false}, big_levels_{false};
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

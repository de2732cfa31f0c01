// C-ABI implementation (include/ngt_b200.h): handle, ingest, schedule of kernel launches, results.
//
// Host orchestration mirrors what the reference's NGT object does (reference source/kmc_rates/ngt.hpp):
//   ctor (ngt.hpp:105-181)            -> ngt_create*: ids, A/B, symbolic analysis, device buffers
//   phase_one (ngt.hpp:338-354)       -> Engine::phase_one: ingest kernel + per-level assembly + blocked elimination
//   phase_two (ngt.hpp:362-431)       -> Engine::phase_two: one task front per a, same elimination kernels
//   rate getters (ngt.hpp:444-513)    -> Engine::finish_rates (host arithmetic on |A|+|B| numbers)
//   committors (ngt.hpp:518-630)      -> Engine::backsub (back-substitution over the kept factors)
// There is no CPU compute path: every numeric step is a CUDA kernel from kernels.cu.
#include "../../include/ngt_b200.h"
#include "kernels.cuh"
#include "ordering.h"
#include "symbolic.h"

#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};
struct InvalidError : std::runtime_error {
    explicit InvalidError(const std::string &m) : std::runtime_error(m) {}
};
struct StateError : std::runtime_error {
    explicit StateError(const std::string &m) : std::runtime_error(m) {}
};
struct NumericError : std::runtime_error {
    explicit NumericError(const std::string &m) : std::runtime_error(m) {}
};

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +      \
                            std::to_string(__LINE__) + ")");                                              \
    } while (0)

// Process-wide cache of device allocations.  cudaMalloc / cudaFree of the multi-hundred-MB front pools cost up to
// ~100 ms each (cudaFree also synchronises the device); a sweep that creates and destroys handles would pay that per
// network.  Freed blocks are kept (up to a cap) and handed back to the next handle on the same device.
// under NGT_B200_TIMING: runtime calls that stall (a cudaMalloc that has to map new memory, a copy behind other work)
struct SlowCall {
    const char *what;
    size_t bytes;
    std::chrono::steady_clock::time_point t0;
    static bool on() {
        static const bool v = std::getenv("NGT_B200_TIMING") != nullptr;
        return v;
    }
    SlowCall(const char *w, size_t b) : what(w), bytes(b) { if (on()) t0 = std::chrono::steady_clock::now(); }
    ~SlowCall() {
        if (!on()) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 2.0) std::fprintf(stderr, "[ngt timing]     slow call: %s of %zu bytes took %.2f ms\n", what, bytes, ms);
    }
};

struct DevCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void *> blocks;   // (device, bytes) -> pointer
    size_t cached = 0;
    static constexpr size_t CAP = (size_t)24 << 30;

    // Requests are rounded up to size classes (four per octave below 256 MB, 2 MB steps above) so that handles of
    // similar networks find each other's blocks: a cudaMalloc of a size the process has not used before occasionally
    // takes 20-90 ms on these boxes (seen for 2 MB blocks), which tripled the cost of a 5 ms small-network build.
    static size_t size_class(size_t bytes) {
        if (bytes <= ((size_t)64 << 10)) return (size_t)64 << 10;
        if (bytes >= ((size_t)256 << 20)) return (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        size_t top = (size_t)1 << 16;
        while ((top << 1) <= bytes) top <<= 1;
        const size_t step = top >> 2;
        return (bytes + step - 1) / step * step;
    }
    void *get(size_t bytes, size_t &cap) {
        bytes = size_class(bytes);
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = blocks.lower_bound(std::make_pair(dev, bytes));
            if (it != blocks.end() && it->first.first == dev && it->first.second <= bytes + bytes / 2 + (1u << 20)) {
                void *p = it->second;
                cap = it->first.second;
                cached -= cap;
                blocks.erase(it);
                return p;
            }
        }
        void *p = nullptr;
        SlowCall sc("cudaMalloc", bytes);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            trim();   // give cached memory back and retry once
            e = cudaMalloc(&p, bytes);
        }
        if (e != cudaSuccess)
            throw CudaError(std::string("cudaMalloc(") + std::to_string(bytes) + " bytes): " + cudaGetErrorString(e));
        cap = bytes;
        return p;
    }
    void put(void *p, size_t cap) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (cached + cap > CAP) {
            SlowCall sc("cudaFree (cache full)", cap);
            cudaFree(p);
            return;
        }
        blocks.emplace(std::make_pair(dev, cap), p);
        cached += cap;
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        SlowCall sc("cache trim", cached);
        for (auto &kv : blocks) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
        }
        cudaSetDevice(cur);
        blocks.clear();
        cached = 0;
    }
};
DevCache g_cache;

// bytes moved between host and device by this library (bench.py reports them per end-to-end step)
std::atomic<long long> g_h2d_bytes{0}, g_d2h_bytes{0};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; cap = o.cap;
            o.p = nullptr; o.n = o.cap = 0;
        }
        return *this;
    }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T *>(g_cache.get(count * sizeof(T), cap));
    }
    void upload(const std::vector<T> &v) {
        alloc(v.size());
        if (!v.empty()) {
            SlowCall sc("cudaMemcpy (upload)", v.size() * sizeof(T));
            CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
            g_h2d_bytes += (long long)(v.size() * sizeof(T));
        }
    }
    void release() {
        if (p) g_cache.put(p, cap);
        p = nullptr;
        n = cap = 0;
    }
    ~DevBuf() { release(); }
};


// ---------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time.  The library takes no link-time dependency on it: the host process (torch, MPI, ...) has
// normally loaded libnccl.so.2 already and dlopen returns that copy.  Only the handful of entry points the sharded
// paths need are declared (ABI stable since NCCL 2.0).
// ---------------------------------------------------------------------------------------------------------------
struct NcclApi {
    typedef struct ncclComm *comm_t;
    struct unique_id { char internal[128]; };
    static constexpr int F64 = 8, SUM = 0;
    int (*GetUniqueId)(unique_id *) = nullptr;
    int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*CommSplit)(comm_t, int, int, comm_t *, void *) = nullptr;   // NCCL >= 2.18; optional
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *lib = nullptr;
    std::string why;

    static NcclApi &get() {
        static NcclApi api = [] {
            NcclApi a;
            const char *names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char *nm : names) {
                a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
                if (a.lib) break;
            }
            if (!a.lib) { a.why = std::string("libnccl.so.2 not found: ") + dlerror(); return a; }
            auto sym = [&](const char *n) { void *p = dlsym(a.lib, n); if (!p) a.why = std::string("NCCL symbol missing: ") + n; return p; };
            a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
            a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
            a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
            a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
            a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
            a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
            a.CommSplit = reinterpret_cast<decltype(a.CommSplit)>(dlsym(a.lib, "ncclCommSplit"));
            a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
            a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
            a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
            return a;
        }();
        if (!api.lib || !api.why.empty()) throw CudaError("NCCL unavailable: " + api.why);
        return api;
    }
};
#define NK(call)                                                                                         \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != 0)                                                                                     \
            throw CudaError(std::string(#call) + ": " + NcclApi::get().GetErrorString(r_) + " (" __FILE__ ":" + \
                            std::to_string(__LINE__) + ")");                                             \
    } while (0)

// one communicator per process and GPU; the collectives run on the engine's own streams
struct Comm {
    NcclApi::comm_t comm = nullptr;
    // a clone of `comm` (ncclCommSplit) for the bulk panel broadcasts of the multi-GPU dense loop: they run on a third
    // stream beside the owner's factorisation, and a communicator of their own keeps them out of the issue order of the
    // small collectives on the main stream.  Null (older NCCL, world of one): everything goes through `comm`, in order.
    NcclApi::comm_t comm2 = nullptr;
    int rank = 0, world = 1, device = 0;
    ~Comm() {
        if (comm2) NcclApi::get().CommDestroy(comm2);
        if (comm) NcclApi::get().CommDestroy(comm);
    }
};

using ngt::DFront;
using ngt::Symbolic;

// One group of fronts eliminated by the same launches.
struct HostBatch {
    std::vector<int> ids;     // front ids
    std::vector<int> f, k;    // their order and pivot count (for grid sizing)
    std::vector<int> fr;      // row counts of rectangular fronts (empty = square)
    size_t dev_off = 0;       // offset into the device id array
    int max_k = 0;
};

struct Engine {
    ngt_options opt{};
    Symbolic S;
    bool dense = false;
    int nbatch = 1;
    int64_t n = 0, nnz = 0;
    std::vector<int64_t> A, B;
    std::vector<double> weights;
    bool have_weights = false;
    int device = 0;
    cudaStream_t stream = nullptr;

    // device: structure
    DevBuf<DFront> d_fronts;
    std::vector<DFront> h_fronts, h_p2fronts;   // host mirrors (single-front launches pass the descriptor by value)
    DevBuf<int> d_ids;              // batch ids (phase 1), then extend-add rounds
    DevBuf<int> d_row_child, d_row_crow, d_rel;                 // assembly: per-parent-row contribution lists, relative maps
    DevBuf<long long> d_row_base, d_arow_ptr, d_rel_off;
    DevBuf<int> d_child_split;   // assembly: per child, where its update columns cross the parent's column quarters
    DevBuf<int> d_fidx;
    DevBuf<long long> d_idx_off;
    std::vector<HostBatch> batches;
    struct Round { int level; size_t dev_off; int count; int max_m; };
    std::vector<Round> rounds;      // assembly launches: the parents of each tree level (max_m = largest parent order)
    // device: values
    DevBuf<double> d_pool;
    long long pool_stride = 0;
    DevBuf<double> d_wbuf, d_rsbuf;
    int maxfb = 1, maxchunks = 1;
    DevBuf<int> d_err;
    DevBuf<double> d_tau0;
    // sparse ingest
    DevBuf<long long> d_row_ptr, d_dest, d_tau_dest, d_dstnodes;
    DevBuf<int> d_row_node, d_korig;
    DevBuf<double> d_k;
    int n_rows = 0;
    // dense ingest
    DevBuf<double> d_K_owned;
    const double *d_K = nullptr;
    DevBuf<int> d_pos, d_node_at;
    // batches of small complete networks: one CTA per network does ingest + phase 1 (kernels.cu, dense_network_kernel);
    // the cross-check GEMM paths and large or single networks keep the multi-kernel path
    bool use_dense_net() const {
        static const bool off = std::getenv("NGT_B200_NO_DENSE_NET") != nullptr;
        return !off && dense && S.nI > 0 && opt.gemm_path == 0 && n <= ngt::DENSE_NET_MAX_N && (nbatch >= 8 || n <= 160);
    }
    // phase 2
    DevBuf<double> d_p2pool;
    long long p2_stride = 0;
    DevBuf<DFront> d_p2fronts;
    DevBuf<int> d_p2ids;
    DevBuf<double> d_om, d_ftau;    // nbatch x m
    // back-substitution
    DevBuf<double> d_x;

    // results (host)
    int m = 0, root_ld = 0;
    std::vector<double> root;       // nbatch x m x root_ld
    std::vector<double> tau0;       // nbatch x n
    std::vector<double> final_om, final_tau;   // nbatch x m
    std::vector<double> rates;      // nbatch x 4
    std::vector<double> xq;         // nbatch x npos x 2
    bool have_phase1 = false, have_phase2 = false, have_backsub = false, values_loaded = false;

    bool use_lookahead = std::getenv("NGT_B200_NO_LOOKAHEAD") == nullptr;
    // 2-D tensor maps over the single front of a dense handle (A box 16 x 128, B box 16 x 16, 128-byte swizzle)
    CUtensorMap tmapA{}, tmapB{};
    bool have_tmaps = false;
    bool use_tmap = true;     // tensor-map (TMA) fed kernel for the large trailing updates of a dense handle
    void make_tensor_maps() {
        have_tmaps = false;
        if (!dense || nbatch != 1 || S.nI <= 0) return;
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess)
            return;
        const ngt::Front &F0 = S.fronts[0];
        void *base = d_pool.p + F0.off;
        const cuuint64_t gdim[2] = {(cuuint64_t)F0.ld, (cuuint64_t)F0.f};
        const cuuint64_t gstride[1] = {(cuuint64_t)F0.ld * sizeof(double)};
        const cuuint32_t estr[2] = {1, 1};
        const cuuint32_t boxA[2] = {16, 128}, boxB[2] = {16, 16};
        EncodeFn enc = reinterpret_cast<EncodeFn>(fn);
        const CUresult ra = enc(&tmapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, boxA, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const CUresult rb = enc(&tmapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, boxB, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        have_tmaps = (ra == CUDA_SUCCESS && rb == CUDA_SUCCESS);
    }

    // the tensor-map (TMA) fed kernel needs K to be a multiple of the 16-pivot stage and to start on one, and tile
    // columns starting on an even index (rows start 64-byte aligned by construction: ld % 8 == 0)
    static bool tma_eligible(const HostBatch &b, size_t i, int j, int mode) {
        ngt::Region R;
        if (!ngt::get_region(b.f[i], b.fr.empty() ? 0 : b.fr[i], b.k[i], j, mode, R)) return false;
        const int tk = ngt::update_tile_k();
        return (R.k1 - R.k0) % tk == 0 && R.k0 % tk == 0 && R.c0 % 2 == 0 && (R.k1 - R.k0) >= 4 * tk;
    }
    bool use_graph = std::getenv("NGT_B200_NO_GRAPH") == nullptr;
    bool use_cluster = std::getenv("NGT_B200_NO_CLUSTER") == nullptr;
    bool capturing = false, last_step_graph = false;
    bool p1_timed = false, p2_timed = false;   // the phase events of the last pass were recorded (not a graph replay)
    long long steps_done = 0;
    cudaGraphExec_t graph_exec = nullptr;
    const double *graph_dK = nullptr;
    bool d2_pending = false;
    bool wslot_per_step = false;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t sync_ev[2] = {nullptr, nullptr};
    cudaEvent_t sync_event(int i) { return sync_ev[i]; }
    bool profile_schur = false;
    // which launches the profiled pass brackets with events: Schur updates (the roofline kernel) unless the
    // experiment switch NGT_B200_PROFILE_KIND says "assemble" or "panel"
    int profile_kind = [] {
        const char *e = std::getenv("NGT_B200_PROFILE_KIND");
        return !e ? 0 : (std::string(e) == "assemble" ? 1 : (std::string(e) == "panel" ? 2 : 0));
    }();
    std::vector<cudaEvent_t> prof_ev;   // pairs around every update launch when profile_schur
    size_t prof_used = 0;
    cudaEvent_t prof_event() {
        if (prof_used == prof_ev.size()) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            prof_ev.push_back(e);
        }
        return prof_ev[prof_used++];
    }

    ngt_stats st{};
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    ~Engine() {
        if (stream) cudaStreamSynchronize(stream);
        if (stream2) cudaStreamSynchronize(stream2);
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        for (auto &e : prof_ev) cudaEventDestroy(e);
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        for (auto &e : sync_ev)
            if (e) cudaEventDestroy(e);
        if (bulk_gate) cudaEventDestroy(bulk_gate);
        if (ev_packed) cudaEventDestroy(ev_packed);
        if (ev_rows) cudaEventDestroy(ev_rows);
        if (stream3) { cudaStreamSynchronize(stream3); cudaStreamDestroy(stream3); }
        if (stream2) cudaStreamDestroy(stream2);
        if (stream) cudaStreamDestroy(stream);
    }

    ngt::LaunchCtx ctx(double *pool, long long stride, const DFront *fr) {
        ngt::LaunchCtx c;
        c.pool = pool;
        c.pool_stride = stride;
        c.nbatch = nbatch;
        c.fronts = fr;
        c.wbuf = d_wbuf.p;
        c.maxfb = maxfb;
        c.rsbuf = d_rsbuf.p;
        c.maxchunks = maxchunks;
        c.errflag = d_err.p;
        c.stream = stream;
        c.gemm_path = opt.gemm_path;
        c.ko = ngt::KOpt{0, 0, 0, ngt::FUSE_COLS};
        return c;
    }

    // ---------------------------------------------------------------------------------------------
    void init_device() {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CudaError(std::string("no CUDA device: ") + cudaGetErrorString(e) +
                            " — kmc_rates_b200 has no CPU fallback");
        device = opt.device;
        if (device < 0 || device >= ndev) throw InvalidError("device ordinal out of range");
        CK(cudaSetDevice(device));
        // the panel chain is latency critical: its stream outranks the look-ahead stream of big trailing updates
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi));
        CK(cudaStreamCreateWithPriority(&stream2, cudaStreamNonBlocking, prio_lo));
        for (auto &x : sync_ev) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&bulk_gate, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_packed, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ev_rows, cudaEventDisableTiming));
        CK(cudaStreamCreateWithPriority(&stream3, cudaStreamNonBlocking, prio_hi));
        for (auto &x : ev) CK(cudaEventCreate(&x));
        d_err.alloc(1);
        CK(cudaMemset(d_err.p, 0, sizeof(int)));
        ngt::init_kernel_attributes();
    }

    void validate_groups() {
        // A may be empty: then every node outside B is eliminated and the back-substitution yields the mean first
        // passage time of every node into B (the linear-solve view, reference rates_linalg.py:246-285)
        if (B.empty()) throw InvalidError("B must contain at least one node");
        std::vector<char> seen(n, 0);
        for (int64_t a : A) {
            if (a < 0 || a >= n) throw InvalidError("an element in the reactant set (A) is not in the graph");
            if (seen[a]) throw InvalidError("duplicate node in A");
            seen[a] = 1;
        }
        for (int64_t b : B) {
            if (b < 0 || b >= n) throw InvalidError("an element in the product set (B) is not in the graph");
            if (seen[b] == 1) throw InvalidError("A and B have at least one node in common");
            if (seen[b] == 2) throw InvalidError("duplicate node in B");
            seen[b] = 2;
        }
    }

    static std::vector<HostBatch> make_batches(const Symbolic &S, std::vector<int> &flat) {
        std::vector<HostBatch> out;
        for (const auto &b : S.batches) {
            HostBatch hb;
            hb.dev_off = flat.size();
            for (int id : b.fronts) {
                hb.ids.push_back(id);
                hb.f.push_back(S.fronts[id].f);
                hb.k.push_back(S.fronts[id].k);
                hb.max_k = std::max(hb.max_k, S.fronts[id].k);
                flat.push_back(id);
            }
            out.push_back(hb);
        }
        return out;
    }

    void upload_structure() {
        const int nf = (int)S.fronts.size();
        std::vector<DFront> hf(nf);
        for (int i = 0; i < nf; ++i) {
            hf[i].off = S.fronts[i].off;
            hf[i].f = S.fronts[i].f;
            hf[i].k = S.fronts[i].k;
            hf[i].ld = S.fronts[i].ld;
            hf[i].parent = S.fronts[i].parent;
        }
        d_fronts.upload(hf);
        h_fronts = hf;
        std::vector<int> flat;
        batches = make_batches(S, flat);
        for (auto &b : batches) maxfb = std::max(maxfb, (int)b.ids.size());

        // assembly: per tree level the parents that have children (one launch per level), and for every parent row the
        // list of child rows that add into it
        rounds.clear();
        {
            std::vector<std::vector<int> > per_level(S.n_levels);
            for (int p = 0; p < nf; ++p)
                if (S.child_ptr[p + 1] > S.child_ptr[p]) per_level[S.fronts[p].level].push_back(p);
            for (int l = 0; l < S.n_levels; ++l) {
                auto &r = per_level[l];
                for (size_t s0 = 0; s0 < r.size(); s0 += 32768) {      // gridDim.y limit
                    Round rd;
                    rd.level = l;
                    rd.dev_off = flat.size();
                    rd.count = (int)std::min<size_t>(32768, r.size() - s0);
                    rd.max_m = 0;                                       // here: largest parent order
                    for (int i = 0; i < rd.count; ++i) {
                        rd.max_m = std::max(rd.max_m, S.fronts[r[s0 + i]].f);
                        flat.push_back(r[s0 + i]);
                    }
                    rounds.push_back(rd);
                }
            }
            // per parent row: the (child, child update row) pairs that contribute to it, children in their fixed order
            std::vector<long long> row_base(nf + 1, 0);
            for (int p = 0; p < nf; ++p) row_base[p + 1] = row_base[p] + (S.child_ptr[p + 1] > S.child_ptr[p] ? S.fronts[p].f : 0);
            std::vector<long long> row_ptr(row_base[nf] + 1, 0);
            // every parent owns a contiguous range of row_ptr, so parents are counted and filled independently
            unsigned hw = ngt::host_threads();
            const int nthr = nf < 512 ? 1 : (int)std::max(1u, std::min(hw ? hw : 1u, 8u));
            auto over_parents = [&](auto &&body) {
                std::vector<std::thread> th;
                for (int t = 1; t < nthr; ++t)
                    th.emplace_back([&, t] { for (int p = nf * t / nthr; p < nf * (t + 1) / nthr; ++p) body(p); });
                for (int p = 0; p < nf / nthr; ++p) body(p);
                for (auto &x : th) x.join();
            };
            over_parents([&](int p) {
                for (long long ci = S.child_ptr[p]; ci < S.child_ptr[p + 1]; ++ci) {
                    const int c = S.child_ids[ci];
                    for (int jx = 0; jx < S.fronts[c].m; ++jx) row_ptr[row_base[p] + S.rel[S.rel_off[c] + jx] + 1]++;
                }
            });
            for (size_t i = 0; i + 1 < row_ptr.size(); ++i) row_ptr[i + 1] += row_ptr[i];
            std::vector<int> row_child(row_ptr.back()), row_crow(row_ptr.back());
            over_parents([&](int p) {
                if (S.child_ptr[p + 1] == S.child_ptr[p]) return;
                std::vector<long long> fill(row_ptr.begin() + row_base[p], row_ptr.begin() + row_base[p] + S.fronts[p].f);
                for (long long ci = S.child_ptr[p]; ci < S.child_ptr[p + 1]; ++ci) {
                    const int c = S.child_ids[ci];
                    for (int jx = 0; jx < S.fronts[c].m; ++jx) {
                        const long long w = fill[S.rel[S.rel_off[c] + jx]]++;
                        row_child[w] = c;
                        row_crow[w] = jx;
                    }
                }
            });
            d_row_base.upload(row_base);
            d_arow_ptr.upload(row_ptr);
            d_row_child.upload(row_child);
            d_row_crow.upload(row_crow);
            d_rel.upload(S.rel);
            std::vector<long long> ro(S.rel_off.begin(), S.rel_off.end());
            d_rel_off.upload(ro);
            // column split points of the children of parents whose rows are shared by ASM_WPR warps (see kernels.cuh)
            std::vector<int> split((size_t)nf * (ngt::ASM_WPR - 1), 0);
            for (const Round &rd : rounds) {
                if (!ngt::assemble_splits_rows(rd.count, nbatch, rd.max_m)) continue;
                for (int i = 0; i < rd.count; ++i) {
                    const int p = flat[rd.dev_off + i];
                    const int q = (S.fronts[p].f + ngt::ASM_WPR - 1) / ngt::ASM_WPR;
                    for (long long ci = S.child_ptr[p]; ci < S.child_ptr[p + 1]; ++ci) {
                        const int c = S.child_ids[ci];
                        const int32_t *rl = S.rel.data() + S.rel_off[c];
                        for (int part = 1; part < ngt::ASM_WPR; ++part)
                            split[(size_t)c * (ngt::ASM_WPR - 1) + part - 1] =
                                (int)(std::lower_bound(rl, rl + S.fronts[c].m, part * q) - rl);
                    }
                }
            }
            d_child_split.upload(split);
        }
        d_ids.upload(flat);
        d_fidx.upload(S.fidx);
        {
            std::vector<long long> io(nf);
            for (int i = 0; i < nf; ++i) io[i] = S.fronts[i].idx_off;
            d_idx_off.upload(io);
        }
        m = (int)(S.nA + S.nB);
        root_ld = S.fronts[S.root_id].ld;
        pool_stride = S.pool_doubles;
        d_pool.alloc((size_t)pool_stride * nbatch);
        d_tau0.alloc((size_t)n * nbatch);
        CK(cudaMemset(d_tau0.p, 0, (size_t)n * nbatch * sizeof(double)));

        build_phase2_structure();
        d_wbuf.alloc((size_t)nbatch * maxfb * ngt::NB_INNER * ngt::NB_INNER);
        maxchunks = std::max(ngt::PANEL_CLUSTER, (S.max_front + ngt::RS_CHUNK - 1) / ngt::RS_CHUNK);
        d_rsbuf.alloc((size_t)nbatch * maxfb * maxchunks * ngt::NB_INNER);
        d_om.alloc((size_t)nbatch * m);
        d_ftau.alloc((size_t)nbatch * m);

        st.n_nodes = n;
        st.n_rates = nnz;
        st.n_eliminated = S.nI;
        st.n_fronts = nf;
        st.n_levels = S.n_levels;
        st.max_front = S.max_front;
        st.pool_doubles = pool_stride * nbatch;
        st.alg_flops = S.alg_flops * nbatch;
        st.alg_bytes = S.alg_bytes * nbatch;
    }

    // Phase-2 plan: per group a binary tree of eliminations (kernels.cu, "phase 2").  Level d holds the fronts whose
    // source is a front of level d-1 (level 0: the reduced root block); both groups share the launches of a level.
    // Sharded runs (rank, nranks) own the members [lo, hi) of every group given by a contiguous split and build only
    // the fronts whose kept set meets their range.
    struct P2Level { size_t task0; int ntask; int max_f; std::vector<HostBatch> batches; };
    std::vector<P2Level> p2levels;
    int p2_nleaf = 0, p2_rank = -1, p2_nranks = 0;
    DevBuf<ngt::P2Task> d_p2tasks;
    DevBuf<int> d_p2leaf_front, d_p2leaf_out;

    static void shard_span(int n, int rank, int nranks, int &lo, int &hi) {
        const int base = n / nranks, extra = n % nranks;
        lo = rank * base + std::min(rank, extra);
        hi = lo + base + (rank < extra ? 1 : 0);
    }

    void build_phase2_structure(int rank = 0, int nranks = 1) {
        if (rank == p2_rank && nranks == p2_nranks) return;
        if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }   // the captured step used the old plan
        struct Node { int group, lo, hi, front; };   // kept members [lo, hi) of `group`, front that produced them
        std::vector<DFront> pf;
        std::vector<ngt::P2Task> tasks;
        std::vector<int> flat, leaf_front, leaf_out;
        p2levels.clear();
        long long off = 0;
        const int ngs[2] = {(int)S.nA, (int)S.nB};
        const int g0s[2] = {0, (int)S.nA};
        int own_lo[2], own_hi[2];
        for (int g = 0; g < 2; ++g) shard_span(ngs[g], rank, nranks, own_lo[g], own_hi[g]);
        auto add_front = [&](P2Level &L, int group, int src, int src_k, int npk, int t_off, int t_len) {
            DFront F{};
            F.f = npk + 1;
            F.k = npk - t_len;
            F.ld = ((F.f + 1 + 7) / 8) * 8;
            F.off = off;
            F.parent = -1;
            off += (long long)F.f * F.ld;
            off = ((off + 15) / 16) * 16;
            ngt::P2Task t{};
            t.src = src; t.src_k = src_k; t.npk = npk; t.t_off = t_off; t.t_len = t_len;
            t.o0 = group == 0 ? (int)S.nA : 0;
            t.no = group == 0 ? (int)S.nB : (int)S.nA;
            const int id = (int)pf.size();
            pf.push_back(F);
            tasks.push_back(t);
            L.ntask++;
            L.max_f = std::max(L.max_f, F.f);
            if (F.k > 0) {
                if (L.batches.empty() || L.batches.back().ids.size() == 32768) {
                    L.batches.emplace_back();
                    L.batches.back().dev_off = flat.size();
                }
                HostBatch &hb = L.batches.back();
                hb.ids.push_back(id); hb.f.push_back(F.f); hb.k.push_back(F.k);
                hb.max_k = std::max(hb.max_k, F.k);
                flat.push_back(id);
            }
            return id;
        };
        // level 0: every group as a whole, read from the root block (a group of one member is its own leaf)
        std::vector<Node> cur, nxt;
        for (int g = 0; g < 2; ++g)
            if (ngs[g] > 0 && own_hi[g] > own_lo[g]) cur.push_back(Node{g, 0, ngs[g], -1});
        while (!cur.empty()) {
            P2Level L{tasks.size(), 0, 0, {}};
            nxt.clear();
            for (const Node &nd : cur) {
                const int h = nd.hi - nd.lo;
                const int src_k = nd.front < 0 ? g0s[nd.group] : pf[nd.front].k;
                auto child = [&](int lo, int hi) {
                    if (hi <= own_lo[nd.group] || lo >= own_hi[nd.group]) return;   // another rank's members
                    const int id = add_front(L, nd.group, nd.front, src_k, h, lo - nd.lo, hi - lo);
                    if (hi - lo == 1) { leaf_front.push_back(id); leaf_out.push_back(g0s[nd.group] + lo); }
                    else nxt.push_back(Node{nd.group, lo, hi, id});
                };
                if (h == 1) child(nd.lo, nd.hi);           // only at level 0: a group with a single member
                else {
                    const int mid = nd.lo + (h + 1) / 2;
                    child(nd.lo, mid);
                    child(mid, nd.hi);
                }
            }
            if (L.ntask > 65535) throw InvalidError("A or B too large for phase 2 (more than 65535 fronts in one tree level)");
            if (L.ntask > 0) p2levels.push_back(std::move(L));
            cur.swap(nxt);
        }
        for (auto &L : p2levels)
            for (auto &b : L.batches) maxfb = std::max(maxfb, (int)b.ids.size());
        p2_nleaf = (int)leaf_front.size();
        p2_stride = off;
        d_p2fronts.upload(pf);
        h_p2fronts = pf;
        d_p2tasks.upload(tasks);
        d_p2ids.upload(flat);
        d_p2leaf_front.upload(leaf_front);
        d_p2leaf_out.upload(leaf_out);
        d_p2pool.alloc((size_t)std::max<long long>(p2_stride, 1) * nbatch);
        // wbuf / rsbuf are sized by the largest batch of either phase
        if (d_wbuf.n < (size_t)nbatch * maxfb * ngt::NB_INNER * ngt::NB_INNER)
            d_wbuf.alloc((size_t)nbatch * maxfb * ngt::NB_INNER * ngt::NB_INNER);
        if (d_rsbuf.p && d_rsbuf.n < (size_t)nbatch * maxfb * maxchunks * ngt::NB_INNER)
            d_rsbuf.alloc((size_t)nbatch * maxfb * maxchunks * ngt::NB_INNER);
        p2_rank = rank;
        p2_nranks = nranks;
    }

    // ---------------------------------------------------------------------------------------------
    // blocked elimination of a batch of fronts (panel -> row panel -> Schur updates)
    // ---------------------------------------------------------------------------------------------
    static int max_tiles(const HostBatch &b, int j, int mode, int tm, int tn, double &flops) {
        int mx = 0;
        ngt::Region R;
        for (size_t i = 0; i < b.ids.size(); ++i) {
            if (!ngt::get_region(b.f[i], b.fr.empty() ? 0 : b.fr[i], b.k[i], j, mode, R)) continue;
            const int t = ((R.r1 - R.r0 + tm - 1) / tm) * ((R.c1 - R.c0 + tn - 1) / tn);
            mx = std::max(mx, t);
            flops += 2.0 * (R.r1 - R.r0) * (double)(R.c1 - R.c0) * (R.k1 - R.k0);
        }
        return mx;
    }

    // (Processing the networks of a batched handle in L2-sized chunks was measured slower on config 5 -- 14.1 vs 12.1 ms,
    // each launch then holds a single wave of latency-bound CTAs -- and is gone; batches of small complete networks now
    // take the whole-network kernel instead.)
    void eliminate(const ngt::LaunchCtx &c0, const HostBatch &b, const int *dev_ids) { eliminate_chunk(c0, b, dev_ids); }

    void eliminate_chunk(const ngt::LaunchCtx &c, const HostBatch &b, const int *dev_ids) {
        const int nbatch = c.nbatch;   // shadows the member: flops and look-ahead sizing are per chunk
        const int nfb = (int)b.ids.size();
        const int tm = ngt::update_tile_m(opt.gemm_path), tn = ngt::update_tile_n(opt.gemm_path);
        const int *ids = dev_ids + b.dev_off;
        // a launch that covers exactly one front carries its descriptor in the kernel parameters
        const DFront *f0 = c.f0;
        if (!f0 && nfb == 1) {
            if (c.fronts == d_fronts.p && (size_t)b.ids[0] < h_fronts.size()) f0 = &h_fronts[b.ids[0]];
            else if (c.fronts == d_p2fronts.p && (size_t)b.ids[0] < h_p2fronts.size()) f0 = &h_p2fronts[b.ids[0]];
        }
        for (int j = 0; j < b.max_k; j += ngt::NB_INNER) {
            ngt::LaunchCtx cj = c;
            cj.f0 = f0;
            if (wslot_per_step) cj.ko.wslot = (j % ngt::NB_OUTER) / ngt::NB_INNER;   // keep every W of the outer block
            int max_cols = 0;   // columns right of the panel incl. the tau column, widest live front
            for (int i = 0; i < nfb; ++i)
                if (b.k[i] > j) max_cols = std::max(max_cols, b.f[i] + 1 - std::min(j + ngt::NB_INNER, b.k[i]));
            // a handful of fronts (top of the assembly tree): clusters of CTAs per front do the row sums, the panel and
            // the row panel in one launch
            cj.panel_cluster = use_cluster && !wslot_per_step && cj.ko.own_G <= 1 &&
                               (long long)nfb * nbatch * ngt::PANEL_CLUSTER <= 148 && max_cols <= ngt::PANEL_CLUSTER_MAX_COLS;
            const bool pe_panel = profile_schur && profile_kind == 2;
            if (pe_panel) CK(cudaEventRecord(prof_event(), stream));
            if (!cj.panel_cluster && max_cols - 1 > ngt::RS_SPLIT) {
                ngt::launch_rowsum(cj, ids, nfb, j, max_cols - 1);
                st.n_launches++;
            }
            ngt::launch_panel(cj, ids, nfb, j);
            st.n_launches++;
            if (!cj.panel_cluster && max_cols > cj.ko.fuse_cols) {
                ngt::launch_rowpanel(cj, ids, nfb, j, max_cols);
                st.n_launches++;
            }
            if (pe_panel) { CK(cudaEventRecord(prof_event(), stream)); st.reserved[0] += 1; }
            double fl = 0, dummy = 0;
            int ta = max_tiles(b, j, ngt::MODE_A, tm, tn, fl);
            int tb = max_tiles(b, j, ngt::MODE_B, tm, tn, fl);
            int td1a = max_tiles(b, j, ngt::MODE_D1A, tm, tn, fl);
            int td1b = max_tiles(b, j, ngt::MODE_D1B, tm, tn, fl);
            int td2 = max_tiles(b, j, ngt::MODE_D2, tm, tn, fl);
            st.schur_flops += fl * nbatch;
            // launches that would fill less than one CTA slot per SM use half-height tiles (twice the CTAs): these
            // thin updates are bound by each SM's L2 port, not by the tensor pipe
            long long live_fronts = 0;
            for (int i = 0; i < nfb; ++i) live_fronts += (b.k[i] > j);
            auto half_tiles = [&](int tiles_full) {
                return opt.gemm_path == 0 && tiles_full > 0 && (long long)tiles_full * live_fronts * nbatch < 148;
            };
            ngt::LaunchCtx c_ab = cj, c_d1 = cj;
            if (half_tiles(ta + tb)) {
                c_ab.tile_m = 64;
                ta = max_tiles(b, j, ngt::MODE_A, 64, tn, dummy);
                tb = max_tiles(b, j, ngt::MODE_B, 64, tn, dummy);
            }
            if (half_tiles(td1a + td1b)) {
                c_d1.tile_m = 64;
                td1a = max_tiles(b, j, ngt::MODE_D1A, 64, tn, dummy);
                td1b = max_tiles(b, j, ngt::MODE_D1B, 64, tn, dummy);
            }
            const bool d2_half = half_tiles(td2);
            if (d2_half) td2 = max_tiles(b, j, ngt::MODE_D2, 64, tn, dummy);
            ngt::LaunchCtx c2 = cj;
            if (d2_half) c2.tile_m = 64;
            bool d2_async = false;
            if (td2 > 0 && !d2_half) {
                long long live = 0;
                for (int i = 0; i < nfb; ++i) live += (b.k[i] > j);
                // look-ahead: a trailing update big enough to keep the GPU busy runs on the second stream while
                // the next outer block's panels proceed on the main stream
                d2_async = use_lookahead && (long long)td2 * live * nbatch >= 148;
            }
            auto timed = [&](cudaStream_t sx, auto &&fn) {
                const bool pe = profile_schur && profile_kind == 0;
                if (pe) CK(cudaEventRecord(prof_event(), sx));
                fn();
                if (pe) CK(cudaEventRecord(prof_event(), sx));
                st.n_launches++;
                if (profile_kind == 0) st.reserved[0] += 1;   // Schur-update launches
            };
            if (ta + tb > 0) timed(stream, [&] { ngt::launch_update(c_ab, ids, nfb, j, ngt::MODE_AB, ta + tb, ta); });
            if (td1a + td1b + td2 > 0) {
                if (d2_async) {
                    // D2 needs the finished panels of this outer block ...
                    CK(cudaEventRecord(sync_event(0), stream));
                    CK(cudaStreamWaitEvent(stream2, sync_event(0), 0));
                }
                if (td1a + td1b > 0) {
                    // ... D1 touches the region the previous D2 was writing
                    if (d2_pending) { CK(cudaStreamWaitEvent(stream, sync_event(1), 0)); d2_pending = false; }
                    timed(stream, [&] { ngt::launch_update(c_d1, ids, nfb, j, ngt::MODE_D1, td1a + td1b, td1a); });
                }
                if (td2 > 0) {
                    const bool via_tmap = !d2_half && use_tmap && have_tmaps && opt.gemm_path == 0 && c2.gemm_path == 0 && nfb == 1 &&
                                          c.fronts == d_fronts.p && b.ids[0] == 0 && tma_eligible(b, 0, j, ngt::MODE_D2);
                    if (d2_async) {
                        c2.stream = stream2;
                        timed(stream2, [&] {
                            if (via_tmap) ngt::launch_update_tmap(c2, ids, j, ngt::MODE_D2, td2, &tmapA, &tmapB);
                            else ngt::launch_update(c2, ids, nfb, j, ngt::MODE_D2, td2, 0);
                        });
                        CK(cudaEventRecord(sync_event(1), stream2));
                        d2_pending = true;
                    } else {
                        if (d2_pending) { CK(cudaStreamWaitEvent(stream, sync_event(1), 0)); d2_pending = false; }
                        timed(stream, [&] {
                            if (via_tmap) ngt::launch_update_tmap(c2, ids, j, ngt::MODE_D2, td2, &tmapA, &tmapB);
                            else ngt::launch_update(c2, ids, nfb, j, ngt::MODE_D2, td2, 0);
                        });
                    }
                }
            }
        }
        // everything after this batch (assembly into parents, next level) reads what D2 wrote
        if (d2_pending) { CK(cudaStreamWaitEvent(stream, sync_event(1), 0)); d2_pending = false; }
    }

    void check_errflag(const char *where) {
        int flag = 0;
        CK(cudaMemcpyAsync(&flag, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        if (flag) {
            CK(cudaMemset(d_err.p, 0, sizeof(int)));
            if (flag & 4) throw std::logic_error(std::string(where) + ": a rate has no slot in its front");
            if (flag & 2) throw NumericError(std::string(where) + ": a node has no outgoing rate (tau = 1/0)");
            throw NumericError(std::string(where) +
                               ": a pivot 1-P_xx was zero or not finite (node cannot reach the rest of the graph)");
        }
    }

    void ingest() {
        CK(cudaMemsetAsync(d_pool.p, 0, (size_t)pool_stride * nbatch * sizeof(double), stream));
        if (dense) {
            const int f0 = S.nI > 0 ? 0 : S.root_id;
            DFront F0{}, R{};
            auto cv = [&](int id) {
                DFront d{};
                d.off = S.fronts[id].off; d.f = S.fronts[id].f; d.k = S.fronts[id].k; d.ld = S.fronts[id].ld;
                d.parent = S.fronts[id].parent;
                return d;
            };
            F0 = cv(f0);
            R = cv(S.root_id);
            ngt::launch_ingest_dense(d_pool.p, pool_stride, nbatch, d_K, (int)n, d_pos.p, (int)S.nI, F0, R, d_tau0.p,
                                     d_err.p, stream);
        } else {
            ngt::launch_ingest_sparse(d_pool.p, pool_stride, nbatch, d_k.p, nnz, n_rows, d_row_ptr.p, d_row_node.p,
                                      d_korig.p, d_dest.p, d_tau_dest.p, d_tau0.p, n, d_err.p, stream);
        }
        st.n_launches += 1;
    }

    void phase_one() {
        if (!values_loaded) throw StateError("no rate constants loaded");
        if (!capturing) last_step_graph = false;
        CK(cudaSetDevice(device));
        st.n_launches = 0;
        st.schur_flops = 0;
        st.reserved[0] = 0;
        prof_used = 0;
        if (!capturing) CK(cudaEventRecord(ev[0], stream));
        if (use_dense_net()) {
            // ingest + elimination of every intermediate + the reduced graph in ONE launch, one CTA per network
            if (!capturing) CK(cudaEventRecord(ev[1], stream));
            ngt::launch_dense_network(d_K, (int)n, nbatch, d_node_at.p, (int)S.nI, d_pool.p, pool_stride, h_fronts[0],
                                      h_fronts[S.root_id], d_tau0.p, d_err.p, stream);
            st.n_launches++;
            CK(cudaGetLastError());
            if (!capturing) CK(cudaEventRecord(ev[2], stream));
            p1_timed = !capturing;
            have_phase1 = true;
            have_phase2 = false;
            have_backsub = false;
            return;
        }
        ingest();
        if (!capturing) CK(cudaEventRecord(ev[1], stream));
        ngt::LaunchCtx c = ctx(d_pool.p, pool_stride, d_fronts.p);
        size_t ri = 0, bi = 0;
        for (int l = 0; l < S.n_levels; ++l) {
            for (; ri < rounds.size() && rounds[ri].level == l; ++ri) {
                const bool pe = profile_schur && profile_kind == 1;
                if (pe) CK(cudaEventRecord(prof_event(), stream));
                ngt::launch_assemble(c, d_ids.p + rounds[ri].dev_off, rounds[ri].count, rounds[ri].max_m, d_row_base.p,
                                     d_arow_ptr.p, d_row_child.p, d_row_crow.p, d_rel.p, d_rel_off.p, d_child_split.p);
                if (pe) { CK(cudaEventRecord(prof_event(), stream)); st.reserved[0] += 1; }
                st.n_launches++;
            }
            for (; bi < batches.size() && S.batches[bi].level == l; ++bi) eliminate(c, batches[bi], d_ids.p);
        }
        CK(cudaGetLastError());     // a bad launch configuration must not pass silently
        if (!capturing) CK(cudaEventRecord(ev[2], stream));
        p1_timed = !capturing;
        have_phase1 = true;
        have_phase2 = false;
        have_backsub = false;
    }

    void download_root() {
        const DFront R{S.fronts[S.root_id].off, S.fronts[S.root_id].f, 0, S.fronts[S.root_id].ld, -1};
        root.resize((size_t)nbatch * m * root_ld);
        const size_t blk = (size_t)m * root_ld;
        if (nbatch <= 8) {   // (a 2-D copy's pitch is limited to 2 GiB; single networks can exceed it)
            for (int b = 0; b < nbatch; ++b)
                CK(cudaMemcpyAsync(root.data() + b * blk, d_pool.p + (size_t)b * pool_stride + R.off,
                                   blk * sizeof(double), cudaMemcpyDeviceToHost, stream));
        } else {
            CK(cudaMemcpy2DAsync(root.data(), blk * sizeof(double), d_pool.p + R.off,
                                 (size_t)pool_stride * sizeof(double), blk * sizeof(double), nbatch,
                                 cudaMemcpyDeviceToHost, stream));
        }
        tau0.resize((size_t)nbatch * n);
        CK(cudaMemcpyAsync(tau0.data(), d_tau0.p, tau0.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        g_d2h_bytes += (long long)((root.size() + tau0.size()) * sizeof(double));
    }

    void phase_two(int rank = 0, int nranks = 1) {
        phase_two_enqueue(rank, nranks);
        phase_two_finish(nranks);
    }

    void phase_two_enqueue(int rank = 0, int nranks = 1) {
        if (!have_phase1) throw StateError("phase_two called before phase_one");
        CK(cudaSetDevice(device));
        if (!capturing) build_phase2_structure(rank, nranks);
        if (!capturing) CK(cudaEventRecord(ev[3], stream));
        const DFront R{S.fronts[S.root_id].off, S.fronts[S.root_id].f, 0, S.fronts[S.root_id].ld, -1};
        ngt::LaunchCtx c1 = ctx(d_pool.p, pool_stride, d_fronts.p);
        ngt::LaunchCtx c2 = ctx(d_p2pool.p, p2_stride, d_p2fronts.p);
        CK(cudaMemsetAsync(d_p2pool.p, 0, (size_t)std::max<long long>(p2_stride, 1) * nbatch * sizeof(double), stream));
        // members owned by other ranks report 0, so that an all-reduce(sum) merges the shards
        CK(cudaMemsetAsync(d_om.p, 0, (size_t)nbatch * m * sizeof(double), stream));
        CK(cudaMemsetAsync(d_ftau.p, 0, (size_t)nbatch * m * sizeof(double), stream));
        for (const P2Level &L : p2levels) {
            ngt::launch_phase2_gather(c1, R, d_p2pool.p, p2_stride, d_p2fronts.p, d_p2tasks.p, (int)L.task0, L.ntask, L.max_f);
            st.n_launches++;
            for (const auto &b : L.batches) eliminate(c2, b, d_p2ids.p);
        }
        ngt::launch_phase2_collect(c2, d_p2pool.p, p2_stride, d_p2fronts.p, d_p2leaf_front.p, d_p2leaf_out.p, p2_nleaf,
                                   d_om.p, d_ftau.p, m);
        st.n_launches++;
        CK(cudaGetLastError());
        if (!capturing) CK(cudaEventRecord(ev[4], stream));
        p2_timed = !capturing;
    }

    // One whole numeric step (ingest + phase 1 + phase 2 kernels).  The launch sequence depends only on the
    // symbolic structure, so it is captured once into a CUDA graph (both streams, their event dependencies and the
    // memsets included) and replayed: ~400 dependent small launches cost one graph launch.
    void step_enqueue() {
        // A single step is cheaper launched directly (capture + instantiation cost a few ms); steps dominated by
        // a few long tensor-bound launches (one huge front) gain nothing from a graph and keep their stream priorities.
        if (!use_graph || profile_schur || steps_done < 1 || S.max_front > 4096) {
            last_step_graph = false;
            phase_one();
            phase_two_enqueue();
            ++steps_done;
            return;
        }
        CK(cudaSetDevice(device));
        build_phase2_structure(0, 1);      // (destroys a graph captured with a sharded phase-2 plan)
        if (!graph_exec || graph_dK != d_K) {
            if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
            cudaGraph_t g = nullptr;
            capturing = true;
            CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed));
            try {
                phase_one();
                phase_two_enqueue();
            } catch (...) {
                cudaStreamEndCapture(stream, &g);
                if (g) cudaGraphDestroy(g);
                capturing = false;
                throw;
            }
            capturing = false;
            CK(cudaStreamEndCapture(stream, &g));
            cudaError_t ie = cudaGraphInstantiate(&graph_exec, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) { graph_exec = nullptr; throw CudaError(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
            graph_dK = d_K;
        }
        CK(cudaGraphLaunch(graph_exec, stream));
        last_step_graph = true;
        p1_timed = p2_timed = false;
        have_phase1 = true;
        have_phase2 = false;
        have_backsub = false;
    }

    void phase_two_finish(int nranks = 1) {
        final_om.resize((size_t)nbatch * m);
        final_tau.resize((size_t)nbatch * m);
        CK(cudaMemcpyAsync(final_om.data(), d_om.p, final_om.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(final_tau.data(), d_ftau.p, final_tau.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        g_d2h_bytes += (long long)(2 * final_om.size() * sizeof(double));
        download_root();
        check_errflag("phase two");
        float ms = 0;
        if (p1_timed) {
            cudaEventElapsedTime(&ms, ev[0], ev[1]); st.ms_ingest = ms;
            cudaEventElapsedTime(&ms, ev[1], ev[2]); st.ms_phase1 = ms;
        }
        if (p2_timed) { cudaEventElapsedTime(&ms, ev[3], ev[4]); st.ms_phase2 = ms; }
        have_phase2 = true;
        if (nranks == 1) finish_rates();
    }

    // reference ngt.hpp:444-513
    void finish_rates() {
        rates.assign((size_t)nbatch * 4, 0.0);
        const int nA = (int)S.nA, nB = (int)S.nB;
        for (int b = 0; b < nbatch; ++b) {
            const double *R = root.data() + (size_t)b * m * root_ld;
            const double *om = final_om.data() + (size_t)b * m, *ft = final_tau.data() + (size_t)b * m;
            const double *t0 = tau0.data() + (size_t)b * n;
            for (int g = 0; g < 2; ++g) {
                const int g0 = g == 0 ? 0 : nA, ng = g == 0 ? nA : nB, o0 = g == 0 ? nA : 0, no = g == 0 ? nB : nA;
                const std::vector<int64_t> &grp = g == 0 ? A : B;
                double s = 0, ss = 0, norm = 0;
                for (int i = 0; i < ng; ++i) {
                    const double w = have_weights ? weights[grp[i]] : 1.0;
                    s += w * om[g0 + i] / ft[g0 + i];
                    double PaB = 0;
                    for (int c = 0; c < no; ++c) PaB += R[(size_t)(g0 + i) * root_ld + o0 + c];
                    ss += w * PaB / t0[grp[i]];
                    norm += w;
                }
                rates[(size_t)b * 4 + g] = s / norm;
                rates[(size_t)b * 4 + 2 + g] = ss / norm;
            }
        }
    }

    // ---------------------------------------------------------------------------------------------
    // Single dense network over several GPUs: 1-D block-cyclic columns (block = NB_OUTER), panel broadcast.
    // Two drivers share the stages below: (1) dist_phase_one(comm) enqueues the whole block loop on the engine's
    // streams with NCCL called directly -- no host synchronisation between blocks; (2) the staged C ABI
    // (ngt_dist_rowsum / factor / panel / apply), synchronous per call, for callers that bring their own
    // collectives (the emulated-rank tests, kmc_rates_b200/distributed.py::dense_phase_one_emulated).
    // ---------------------------------------------------------------------------------------------
    int dist_G = 0, dist_g = 0;
    static constexpr int LD_PAN = ngt::NB_OUTER + 8;
    static constexpr int W_DOUBLES = (ngt::NB_OUTER / ngt::NB_INNER) * ngt::NB_INNER * ngt::NB_INNER;
    DevBuf<double> d_pan, d_spart, d_tauglob, d_rootc;
    size_t pan_stride = 0;
    double *pan_buf(int J) { return d_pan.p + (size_t)((J / ngt::NB_OUTER) & 1) * pan_stride; }
    // descriptors of every outer block, uploaded once: [2 blk] the rectangular panel front, [2 blk + 1] the big matrix
    // with that block's panel buffer as the A operand of its Schur updates
    DevBuf<DFront> d_dfronts;
    std::vector<DFront> h_dfronts;
    DevBuf<int> d_dids;

    ngt::KOpt own() const { return ngt::KOpt{dist_G, dist_g, 0, 0}; }

    void dist_begin(int g, int G) {
        if (!dense || nbatch != 1) throw InvalidError("the multi-GPU dense path needs one complete network (ngt_create_dense*)");
        if (G < 1 || g < 0 || g >= G) throw InvalidError("bad rank / world size");
        if (S.nI <= 0) throw InvalidError("no intermediate nodes to eliminate");
        if (!values_loaded) throw StateError("no rate constants loaded");
        CK(cudaSetDevice(device));
        dist_G = G;
        dist_g = g;
        const ngt::Front &F0 = S.fronts[0];
        pan_stride = (size_t)W_DOUBLES + (size_t)F0.f * LD_PAN;
        d_pan.alloc(2 * pan_stride);   // double buffered: block J+1's panel arrives while block J's update still reads J's
        d_spart.alloc(ngt::NB_OUTER);
        d_tauglob.alloc((size_t)n);
        d_rootc.alloc((size_t)m * m);
        const int nblk = (int)((S.nI + ngt::NB_OUTER - 1) / ngt::NB_OUTER);
        h_dfronts.assign((size_t)2 * nblk, DFront{});
        std::vector<int> iota((size_t)2 * nblk);
        for (int b = 0; b < nblk; ++b) {
            const int J = b * ngt::NB_OUTER, kb = block_end(J) - J, fr = F0.f - J;
            double *pan = pan_buf(J) + W_DOUBLES;
            DFront &P = h_dfronts[2 * b];
            P.off = pan - d_pool.p;          // the kernels address fronts relative to the pool base
            P.f = kb + 1;                    // kb block columns + S; column kb+1 is tau
            P.k = kb;
            P.ld = LD_PAN;
            P.parent = -1;
            P.fr = fr;
            DFront &B = h_dfronts[2 * b + 1];
            B.off = F0.off; B.f = F0.f; B.k = F0.k; B.ld = F0.ld; B.parent = -1; B.fr = 0;
            B.a_ld = LD_PAN;
            B.a_off = (pan - d_pool.p) - (long long)J * LD_PAN - J;   // A(r, p) = pan[(r - J) * LD_PAN + (p - J)]
            iota[2 * b] = 2 * b;
            iota[2 * b + 1] = 2 * b + 1;
        }
        d_dfronts.upload(h_dfronts);
        d_dids.upload(iota);
        d2_pending = false;
        bulk_J = -1;                   // (a previous loop may have been abandoned on an error)
        if ((int)(ngt::NB_OUTER / ngt::NB_INNER) > maxfb) {
            // wbuf grows (one W per inner panel of an outer block): a step graph captured earlier has the old pointer
            // and stride baked in
            if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
            maxfb = ngt::NB_OUTER / ngt::NB_INNER;
            d_wbuf.alloc((size_t)nbatch * maxfb * ngt::NB_INNER * ngt::NB_INNER);
        }
        st.n_launches = 0;
        st.schur_flops = 0;
        ingest();
        ngt::launch_dist_tau_init(d_tau0.p, d_pos.p, (int)n, d_tauglob.p, stream);
        CK(cudaStreamSynchronize(stream));
        check_errflag("ingest");
        have_phase1 = have_phase2 = have_backsub = false;
    }

    void dist_check(int J) const {
        if (dist_G < 1) throw StateError("ngt_dist_begin has not been called");
        if (J < 0 || J >= S.nI || J % ngt::NB_OUTER) throw InvalidError("J must be a multiple of the outer block inside the pivot range");
    }
    int block_end(int J) const { return (int)std::min<int64_t>(J + ngt::NB_OUTER, S.nI); }

    // partial mass of the block rows into this rank's columns right of the block -> d_spart (to be all-reduced)
    void dist_rowsum_enqueue(int J) {
        const ngt::Front &F0 = S.fronts[0];
        const int JE = block_end(J);
        CK(cudaMemsetAsync(d_spart.p, 0, ngt::NB_OUTER * sizeof(double), stream));
        ngt::launch_dist_rowsum(d_pool.p + F0.off, F0.ld, F0.f, J, JE, own(), d_spart.p, stream);
        st.n_launches++;
    }

    // owner: build the panel front from its own columns and factor it with the ordinary front kernels.
    // stage 0 = everything (staged API); the NCCL loop runs 1 = pack, 2 = factor the square block, and the column-strip
    // replay separately on every rank (dist_colstrip_enqueue)
    void dist_colstrip_enqueue(int J) {
        const int kb = block_end(J) - J, fr = S.fronts[0].f - J;
        ngt::launch_dist_colstrip_slab(pan_buf(J) + W_DOUBLES, LD_PAN, kb, fr, stream);
        st.n_launches++;
    }
    void dist_factor_enqueue(int J, int stage = 0) {
        const ngt::Front &F0 = S.fronts[0];
        const int JE = block_end(J), kb = JE - J, fr = F0.f - J;
        const int blk = J / ngt::NB_OUTER;
        // Only the block's own kb rows are factored with the front kernels (a square front); the rows below take the
        // block's column operations afterwards in one slab launch.  (The cross-check GEMM paths factor the whole
        // rectangular panel right-looking instead: 8 thin update launches over all rows, the round-1 schedule.)
        const bool square = opt.gemm_path == 0;
        DFront P = h_dfronts[2 * blk];
        if (square) P.fr = kb;
        double *pan = pan_buf(J) + W_DOUBLES;
        if (stage != 2) ngt::launch_dist_pack(d_pool.p + F0.off, F0.ld, J, kb, fr, pan, LD_PAN, d_spart.p, d_tauglob.p, stream);
        if (stage == 1) return;
        HostBatch hb;
        hb.ids = {2 * blk}; hb.f = {P.f}; hb.k = {P.k}; hb.fr = {P.fr}; hb.max_k = kb; hb.dev_off = 2 * blk;
        ngt::LaunchCtx c = ctx(d_pool.p, pool_stride, d_dfronts.p);
        c.f0 = &P;
        wslot_per_step = true;
        const bool la = use_lookahead;
        use_lookahead = false;
        // the panel buffer is disjoint from the region a pending bulk update (D2 of the previous block, second stream)
        // writes: eliminate_chunk must not make the main stream wait for it after the factorisation
        const bool pend = d2_pending;
        d2_pending = false;
        try {
            eliminate_chunk(c, hb, d_dids.p);
        } catch (...) {
            wslot_per_step = false;
            use_lookahead = la;
            d2_pending = pend;
            throw;
        }
        wslot_per_step = false;
        use_lookahead = la;
        d2_pending = pend;
        if (square && stage == 0) dist_colstrip_enqueue(J);
        CK(cudaMemcpyAsync(pan_buf(J), d_wbuf.p, W_DOUBLES * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    }

    // what travels: [W of the block's inner panels | panel front rows J..f-1]
    void dist_panel(int J, double **ptr, int64_t *cnt) {
        dist_check(J);
        *ptr = pan_buf(J);
        *cnt = (int64_t)W_DOUBLES + (int64_t)(S.fronts[0].f - J) * LD_PAN;
    }

    // everybody: replay the block's row operations on the own columns, then the trailing update with the panel as A.
    // Look-ahead: only the strips the NEXT block needs (its columns, its rows) are updated on the main stream; the
    // rest of the trailing update runs on the second stream and overlaps the next block's row sums, all-reduce,
    // panel factorisation and broadcast.
    // defer_bulk: do not launch the bulk of the trailing update (D2) yet -- dist_bulk_enqueue(J) does it later (the rank
    // that factors the NEXT block holds its own bulk update back until that factorisation is in the queue)
    void dist_apply_enqueue(int J, bool defer_bulk = false) {
        const ngt::Front &F0 = S.fronts[0];
        const int JE = block_end(J), kb = JE - J, fr = F0.f - J;
        const int blk = J / ngt::NB_OUTER;
        const double *pan = pan_buf(J) + W_DOUBLES;
        CK(cudaMemcpyAsync(d_wbuf.p, pan_buf(J), W_DOUBLES * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        ngt::launch_dist_tau(pan, LD_PAN, kb, fr, J, d_tauglob.p, stream);
        const DFront &B = h_dfronts[2 * blk + 1];
        HostBatch hb;
        hb.ids = {2 * blk + 1}; hb.f = {B.f}; hb.k = {B.k}; hb.max_k = B.k; hb.dev_off = 0;
        const int tm = ngt::update_tile_m(opt.gemm_path), tn = ngt::update_tile_n(opt.gemm_path);
        ngt::LaunchCtx c = ctx(d_pool.p, pool_stride, d_dfronts.p);
        c.ko = own();
        c.f0 = &B;
        const int *ids = d_dids.p + 2 * blk + 1;
        double fl = 0;
        if (opt.gemm_path == 0) {
            // all inner panels of the block replayed on the own columns in one launch
            ngt::launch_dist_apply_slab(d_pool.p + F0.off, F0.ld, F0.f, J, kb, pan, LD_PAN, d_wbuf.p, own(), stream);
            for (int j = J; j < JE; j += ngt::NB_INNER) (void)max_tiles(hb, j, ngt::MODE_B, tm, tn, fl);   // flop count only
            st.n_launches += 1;
        } else {
            for (int j = J; j < JE; j += ngt::NB_INNER) {
                c.ko.wslot = (j - J) / ngt::NB_INNER;
                const int je = std::min(j + ngt::NB_INNER, JE);
                ngt::launch_rowpanel(c, ids, 1, j, F0.f + 1 - je);
                const int tb = max_tiles(hb, j, ngt::MODE_B, tm, tn, fl);
                if (tb > 0) ngt::launch_update(c, ids, 1, j, ngt::MODE_B, tb, 0);
                st.n_launches += 2;
            }
        }
        const int jl = J + ((kb - 1) / ngt::NB_INNER) * ngt::NB_INNER;
        const int t1a = max_tiles(hb, jl, ngt::MODE_D1A, tm, tn, fl);
        const int t1b = max_tiles(hb, jl, ngt::MODE_D1B, tm, tn, fl);
        const int t2 = max_tiles(hb, jl, ngt::MODE_D2, tm, tn, fl);
        // the strips (and everything after them) touch the region the previous block's D2 was writing
        if (d2_pending) { CK(cudaStreamWaitEvent(stream, sync_event(1), 0)); d2_pending = false; }
        if (t1a + t1b > 0) { ngt::launch_update(c, ids, 1, jl, ngt::MODE_D1, t1a + t1b, t1a); st.n_launches++; }
        st.schur_flops += fl / dist_G;   // region flops; a rank performs its 1/G share of the columns
        bulk_J = -1;
        if (t2 > 0) {
            if (use_lookahead) {
                CK(cudaEventRecord(sync_event(0), stream));     // "the strips of block J are updated"
                bulk_J = J;
                bulk_tiles = t2;
                if (!defer_bulk) dist_bulk_enqueue();
            } else {
                ngt::launch_update(c, ids, 1, jl, ngt::MODE_D2, t2, 0);
                st.n_launches++;
            }
        }
    }

    // the bulk (D2) of block bulk_J's trailing update, on the second stream; after_main: it must also wait for what the
    // main stream holds at this moment (the owner of the next block: its panel factorisation -- both would otherwise
    // fight for the same SMs and the factorisation, which every other rank waits for, would lose)
    int bulk_J = -1, bulk_tiles = 0;
    void dist_bulk_enqueue(bool after_main = false) {
        if (bulk_J < 0) return;
        const int J = bulk_J, blk = J / ngt::NB_OUTER;
        const int JE = block_end(J), kb = JE - J;
        const int jl = J + ((kb - 1) / ngt::NB_INNER) * ngt::NB_INNER;
        const DFront &B = h_dfronts[2 * blk + 1];
        ngt::LaunchCtx c2 = ctx(d_pool.p, pool_stride, d_dfronts.p);
        c2.ko = own();
        c2.f0 = &B;
        c2.stream = stream2;
        CK(cudaStreamWaitEvent(stream2, sync_event(0), 0));
        if (after_main) {
            CK(cudaEventRecord(bulk_gate, stream));
            CK(cudaStreamWaitEvent(stream2, bulk_gate, 0));
        }
        ngt::launch_update(c2, d_dids.p + 2 * blk + 1, 1, jl, ngt::MODE_D2, bulk_tiles, 0);
        CK(cudaEventRecord(sync_event(1), stream2));
        d2_pending = true;
        st.n_launches++;
        bulk_J = -1;
    }
    cudaEvent_t bulk_gate = nullptr;
    cudaStream_t stream3 = nullptr;                  // bulk panel broadcasts of the multi-GPU dense loop
    cudaEvent_t ev_packed = nullptr, ev_rows = nullptr;

    void dist_root_contrib_enqueue() {
        const ngt::Front &F0 = S.fronts[0];
        if (d2_pending) { CK(cudaStreamWaitEvent(stream, sync_event(1), 0)); d2_pending = false; }
        ngt::launch_dist_root_contrib(d_pool.p + F0.off, F0.ld, F0.f, (int)S.nI, m, own(), d_rootc.p, stream);
    }

    void dist_finish() {
        if (dist_G < 1) throw StateError("ngt_dist_begin has not been called");
        const ngt::Front &R = S.fronts[S.root_id];
        ngt::launch_dist_root_finish(d_pool.p + R.off, R.ld, m, d_rootc.p, d_tauglob.p, (int)S.nI, stream);
        CK(cudaStreamSynchronize(stream));
        check_errflag("distributed phase one");
        have_phase1 = true;
        have_phase2 = have_backsub = false;
        last_step_graph = true;   // no per-phase events were recorded for this pass
        p1_timed = false;
    }

    // staged, synchronous variants (C ABI ngt_dist_*)
    double *dist_rowsum(int J) {
        dist_check(J);
        dist_rowsum_enqueue(J);
        CK(cudaStreamSynchronize(stream));
        return d_spart.p;
    }
    void dist_factor(int J) {
        dist_check(J);
        dist_factor_enqueue(J);
        CK(cudaStreamSynchronize(stream));
    }
    void dist_apply(int J) {
        dist_check(J);
        dist_apply_enqueue(J);
        CK(cudaStreamSynchronize(stream));
    }
    double *dist_root_contrib() {
        if (dist_G < 1) throw StateError("ngt_dist_begin has not been called");
        dist_root_contrib_enqueue();
        CK(cudaStreamSynchronize(stream));
        return d_rootc.p;
    }

    // The whole phase 1 of one dense network over the ranks of `cm`: every stage of every block is enqueued on the
    // engine's streams, the three collectives per block (row-sum all-reduce, panel broadcast; root all-reduce at the
    // end) are NCCL calls on the same stream, so the GPU never waits for the host between blocks.
    // ms (optional): device time of the loop from CUDA events on the main stream.
    void dist_phase_one(Comm &cm, double *ms) {
        NcclApi &nc = NcclApi::get();
        if (cm.device != device) throw InvalidError("communicator and handle live on different devices");
        CK(cudaSetDevice(device));
        CK(cudaEventRecord(ev[0], stream));     // the timed region includes the ingest kernel, like a single-GPU step
        dist_begin(cm.rank, cm.world);
        const int block = ngt::NB_OUTER;
        // NGT_B200_DIST_TRACE=1: CUDA events between the stages of every block on the main stream; the per-stage sums
        // are printed to stderr afterwards (where does a block's time go at this world size?)
        static const bool trace = std::getenv("NGT_B200_DIST_TRACE") != nullptr;
        std::vector<cudaEvent_t> tev;
        auto mark = [&] {
            if (!trace) return;
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            CK(cudaEventRecord(e, stream));
            tev.push_back(e);
        };
        for (int J = 0; J < (int)S.nI; J += block) {
            mark();
            dist_rowsum_enqueue(J);
            mark();
            NK(nc.AllReduce(d_spart.p, d_spart.p, (size_t)block, NcclApi::F64, NcclApi::SUM, cm.comm, stream));
            mark();
            const int owner = (J / block) % cm.world;
            const int kb = block_end(J) - J;
            const bool split_bcast = opt.gemm_path == 0 && cm.world > 1;
            double *p = nullptr;
            int64_t cnt = 0;
            dist_panel(J, &p, &cnt);
            if (split_bcast) {
                // The rows below the block (the bulk of the panel, up to 34 MB) travel RAW, on a third stream, while the
                // owner factors the square block; afterwards only the factored block rows + W (0.6 MB) are broadcast
                // and every rank replays the block's column operations on its copy of the rows below (39 us, instead
                // of 120 us of broadcast behind 230 us of factorisation on everybody's critical path).
                const size_t small_cnt = (size_t)W_DOUBLES + (size_t)kb * LD_PAN;
                if (cm.rank == owner) dist_factor_enqueue(J, 1);                       // pack
                CK(cudaEventRecord(ev_packed, stream));
                CK(cudaStreamWaitEvent(stream3, ev_packed, 0));
                if ((size_t)cnt > small_cnt)
                    NK(nc.Broadcast(p + small_cnt, p + small_cnt, (size_t)cnt - small_cnt, NcclApi::F64, owner,
                                    cm.comm2 ? cm.comm2 : cm.comm, stream3));
                CK(cudaEventRecord(ev_rows, stream3));
                if (cm.rank == owner) dist_factor_enqueue(J, 2);                       // factor the square block
                dist_bulk_enqueue(true);      // the previous block's bulk update, if this rank held it back (see below)
                mark();
                NK(nc.Broadcast(p, p, small_cnt, NcclApi::F64, owner, cm.comm, stream));
                CK(cudaStreamWaitEvent(stream, ev_rows, 0));
                dist_colstrip_enqueue(J);
            } else {
                if (cm.rank == owner) dist_factor_enqueue(J);
                dist_bulk_enqueue(true);
                mark();
                NK(nc.Broadcast(p, p, (size_t)cnt, NcclApi::F64, owner, cm.comm, stream));
            }
            mark();
            // the rank that factors the next block keeps its bulk update out of that factorisation's way
            const bool next_owner = cm.world > 1 && J + block < (int)S.nI && cm.rank == ((J / block) + 1) % cm.world;
            dist_apply_enqueue(J, next_owner);
            mark();
        }
        dist_bulk_enqueue();
        dist_root_contrib_enqueue();
        NK(nc.AllReduce(d_rootc.p, d_rootc.p, (size_t)m * m, NcclApi::F64, NcclApi::SUM, cm.comm, stream));
        CK(cudaEventRecord(ev[2], stream));
        CK(cudaGetLastError());
        dist_finish();
        if (ms) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, ev[0], ev[2]));
            *ms = t;
        }
        if (trace) {
            const char *names[5] = {"row sums", "all-reduce (incl. waiting for the slowest rank)", "panel factorisation (owner only)",
                                    "panel broadcast (incl. waiting for the owner)", "row operations + strips (incl. waiting for the bulk update)"};
            double sum[5] = {0, 0, 0, 0, 0};
            for (size_t i = 0; i + 5 < tev.size() + 1 && i + 5 <= tev.size() - 1 + 1; i += 6)
                for (int k = 0; k < 5 && i + k + 1 < tev.size(); ++k) {
                    float t = 0;
                    cudaEventElapsedTime(&t, tev[i + k], tev[i + k + 1]);
                    sum[k] += t;
                }
            std::fprintf(stderr, "[ngt dist trace] rank %d of %d, %d blocks:", cm.rank, cm.world, (int)(tev.size() / 6));
            for (int k = 0; k < 5; ++k) std::fprintf(stderr, "  %s %.2f ms;", names[k], sum[k]);
            std::fprintf(stderr, "\n");
            for (auto &e : tev) cudaEventDestroy(e);
        }
    }

    // Sharded phase 2 over the ranks of `cm`: optional NCCL broadcast of rank 0's reduced block (so that every rank
    // reduces bit-identical input), this rank's branches of the elimination tree, one in-place all-reduce(sum) of the
    // (1-P'_aa, tau'_a) vectors on the engine's stream, then the rates on every rank.
    void phase_two_sharded(Comm &cm, bool broadcast_root, double *ms) {
        NcclApi &nc = NcclApi::get();
        if (!have_phase1) throw StateError("phase_two called before phase_one");
        if (cm.device != device) throw InvalidError("communicator and handle live on different devices");
        CK(cudaSetDevice(device));
        cudaEvent_t e0 = prof_event(), e1 = prof_event();
        CK(cudaEventRecord(e0, stream));
        if (broadcast_root && cm.world > 1) {
            double *blk = d_pool.p + S.fronts[S.root_id].off;
            for (int b = 0; b < nbatch; ++b)
                NK(nc.Broadcast(blk + (size_t)b * pool_stride, blk + (size_t)b * pool_stride, (size_t)m * root_ld,
                                NcclApi::F64, 0, cm.comm, stream));
        }
        phase_two_enqueue(cm.rank, cm.world);
        if (cm.world > 1) {
            NK(nc.AllReduce(d_om.p, d_om.p, (size_t)nbatch * m, NcclApi::F64, NcclApi::SUM, cm.comm, stream));
            NK(nc.AllReduce(d_ftau.p, d_ftau.p, (size_t)nbatch * m, NcclApi::F64, NcclApi::SUM, cm.comm, stream));
        }
        CK(cudaEventRecord(e1, stream));
        phase_two_finish(1);      // downloads the merged vectors and forms the rates
        if (ms) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, e0, e1));
            *ms = t;
        }
    }

    // committors and MFPTs of every eliminated node by back-substitution (top-down over the assembly tree)
    void backsub() {
        if (!have_phase1) throw StateError("committors requested before phase one");
        CK(cudaSetDevice(device));
        const int64_t npos = S.nI + S.nA + S.nB;
        std::vector<double> x0((size_t)npos * 2, 0.0);
        for (int64_t i = 0; i < S.nB; ++i) x0[(size_t)(S.nI + S.nA + i) * 2] = 1.0;
        d_x.alloc((size_t)nbatch * npos * 2);
        for (int b = 0; b < nbatch; ++b)
            CK(cudaMemcpyAsync(d_x.p + (size_t)b * npos * 2, x0.data(), x0.size() * sizeof(double),
                               cudaMemcpyHostToDevice, stream));
        ngt::LaunchCtx c = ctx(d_pool.p, pool_stride, d_fronts.p);
        for (size_t bi = batches.size(); bi-- > 0;) {
            const HostBatch &b = batches[bi];
            const int last = ((b.max_k - 1) / ngt::NB_INNER) * ngt::NB_INNER;
            for (int j = last; j >= 0; j -= ngt::NB_INNER) {
                ngt::launch_backsub_step(c, d_ids.p + b.dev_off, (int)b.ids.size(), d_fidx.p, d_idx_off.p, d_x.p,
                                         npos * 2, j);
                st.n_launches++;
            }
        }
        xq.resize((size_t)nbatch * npos * 2);
        CK(cudaMemcpyAsync(xq.data(), d_x.p, xq.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        g_h2d_bytes += (long long)(nbatch * x0.size() * sizeof(double));
        g_d2h_bytes += (long long)(xq.size() * sizeof(double));
        have_backsub = true;
    }
};

// ---------------------------------------------------------------------------------------------------
// construction helpers
// ---------------------------------------------------------------------------------------------------
Engine *make_engine(const ngt_options *opt, int64_t nbatch) {
    Engine *E = new Engine();
    if (opt) E->opt = *opt;
    if (nbatch < 1 || nbatch > 65535) { delete E; throw InvalidError("nbatch must be in 1..65535 (one grid dimension per network)"); }
    E->nbatch = (int)nbatch;
    return E;
}

// Host buffers of the sparse analysis, kept between calls (one set per process, handed to one build at a time): on a
// fresh allocation the first-touch page faults of these ~150 MB cost more than the passes that fill them.
struct HostWorkspace {
    ngt::EdgeIndex ix;
    ngt::Adjacency g;
    ngt::IndexScratch is;
    std::vector<int> row_node, korig;
    std::vector<long long> row_ptr, tau_dest, dest;
    std::vector<std::vector<int32_t> > marks;
    std::vector<int32_t> row_cnt;
    std::vector<char> reach;
    std::vector<int32_t> stack;
    size_t bytes() const {
        size_t b = ix.outptr.capacity() * 8 + ix.ent.capacity() * 4 + ix.present.capacity() + g.ptr.capacity() * 8 +
                   g.adj.capacity() * 4 + is.tmp.capacity() * 4 + (is.aptr.capacity() + is.ucnt.capacity()) * 8 +
                   row_node.capacity() * 4 + korig.capacity() * 4 + (row_ptr.capacity() + tau_dest.capacity() + dest.capacity()) * 8;
        for (auto &v : is.c_out) b += v.capacity() * 4;
        for (auto &v : is.c_adj) b += v.capacity() * 4;
        for (auto &v : is.seen) b += v.capacity() * 4;
        for (auto &v : is.w_out) b += v.capacity() * 8;
        for (auto &v : is.w_adj) b += v.capacity() * 8;
        for (auto &v : marks) b += v.capacity() * 4;
        return b;
    }
};
// a piece of work that runs beside the symbolic analysis; its exception (if any) is re-thrown where the work used to be
struct SideTask {
    std::thread th;
    std::exception_ptr err;
    template <class F>
    void start(F f) {
        th = std::thread([this, f] {
            try { f(); } catch (...) { err = std::current_exception(); }
        });
    }
    void finish() {
        if (th.joinable()) th.join();
        if (err) {
            std::exception_ptr e = err;
            err = nullptr;
            std::rethrow_exception(e);
        }
    }
    ~SideTask() { if (th.joinable()) th.join(); }
};

static std::mutex g_ws_mu;
static std::unique_ptr<HostWorkspace> g_ws_cached;
constexpr size_t HOST_WS_KEEP_BYTES = (size_t)1 << 30;   // larger workspaces are released after use

struct WorkspaceLease {
    std::unique_ptr<HostWorkspace> ws;
    WorkspaceLease() {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        ws = std::move(g_ws_cached);
        if (!ws) ws.reset(new HostWorkspace());
    }
    ~WorkspaceLease() {
        if (ws && ws->bytes() <= HOST_WS_KEEP_BYTES) {
            std::lock_guard<std::mutex> lk(g_ws_mu);
            if (!g_ws_cached) g_ws_cached = std::move(ws);
        }
    }
};
void trim_host_workspace() {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    g_ws_cached.reset();
}

void build_sparse(Engine *E, int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, const double *k,
                  int64_t nA, const int64_t *A, int64_t nB, const int64_t *B) {
    auto t0 = std::chrono::steady_clock::now();
    ngt::Timer tm;
    if (n <= 0 || nnz < 0) throw InvalidError("empty network");
    if (n > 2000000000LL || nnz > 2000000000LL) throw InvalidError("too many nodes or rates (32-bit entry indices)");
    E->n = n;
    E->nnz = nnz;
    E->dense = false;
    E->A.assign(A, A + nA);
    E->B.assign(B, B + nB);
    E->validate_groups();
    WorkspaceLease lease;
    HostWorkspace &W = *lease.ws;
    ngt::EdgeIndex &ix = W.ix;
    ngt::Adjacency &g = W.g;
    ngt::index_edges(n, nnz, src, dst, ix, g, &W.is);
    if (ix.range_error) throw InvalidError("rate endpoint out of range");
    std::vector<char> &present = ix.present;
    std::vector<int64_t> &outdeg = ix.outptr;     // row pointers of the CSR-by-source entry lists
    for (int64_t a : E->A) if (!present[a]) throw InvalidError("an element in the reactant set (A) is not in the graph");
    for (int64_t b : E->B) if (!present[b]) throw InvalidError("an element in the product set (B) is not in the graph");
    bool all_present = true;
    for (int64_t u = 0; u < n; ++u) {
        if (present[u] && outdeg[u + 1] == outdeg[u]) throw InvalidError("node " + std::to_string(u) + " has no outgoing rate");
        all_present = all_present && present[u];
    }
    tm.lap("index edges (validate, CSR buckets, adjacency)");
    // Beside the analysis (which leaves cores idle while it walks the top of the dissection tree): (1) the device is
    // initialised and the rate constants are uploaded, (2) the entries that survive de-duplication are counted per
    // source node ("last wins", std::map semantics of ngt.pyx:76-81) for the rate -> slot table built afterwards.
    const std::vector<int32_t> &ent = ix.ent;
    auto dedup_row = [&](int64_t u, std::vector<int32_t> &mark, auto &&emit) {
        // later duplicates win: walk the row backwards and keep the first sighting of every destination
        for (int64_t q = outdeg[u + 1]; q-- > outdeg[u];) {
            const int32_t e = ent[q];
            const int64_t v = dst[e];
            if (mark[v] == (int32_t)u) continue;
            mark[v] = (int32_t)u;
            emit(e);
        }
    };
    // the sweeps of the largest subgraphs of the dissection run on the device (ordering.cu); the first of them is on
    // the critical path of the whole build, so the bulk copies below wait until its subgraph has gone up (they have
    // ~30 ms of analysis to hide behind, and two pageable copies at once halve each other's rate)
    ngt::SweepAccel *accel = E->opt.host_sweeps ? nullptr : ngt::gpu_sweeps(E->opt.device);
    const int32_t sweep_min = E->opt.sweep_min_nodes > 0 ? E->opt.sweep_min_nodes : (accel ? accel->min_nodes : 0);
    const int uploads_before = accel ? accel->uploads.load(std::memory_order_acquire) : 0;
    std::atomic<bool> analysis_done{false};
    SideTask device_task, count_task;
    device_task.start([&] {
        E->init_device();
        E->d_k.alloc((size_t)nnz * E->nbatch);
        if (accel && n >= sweep_min) {
            const auto t_gate = std::chrono::steady_clock::now();
            while (accel->uploads.load(std::memory_order_acquire) == uploads_before && !analysis_done.load(std::memory_order_acquire) &&
                   std::chrono::steady_clock::now() - t_gate < std::chrono::milliseconds(8))
                std::this_thread::sleep_for(std::chrono::microseconds(100));
        }
        if (k) {
            CK(cudaMemcpy(E->d_k.p, k, (size_t)nnz * E->nbatch * sizeof(double), cudaMemcpyHostToDevice));
            g_h2d_bytes += (long long)((size_t)nnz * E->nbatch * sizeof(double));
            E->values_loaded = true;
        }
        // Every node named by a rate (the usual case): the CSR-by-source edge index doubles as the ingest kernel's row
        // structure and the rate -> slot table is built on the device once the analysis is known, so the index and
        // the destination ids go up now, beside the ordering.  (Networks with absent ids, duplicate rates or pruned
        // components take the host-built table below, which overwrites these.)
        if (all_present) {
            static_assert(sizeof(long long) == sizeof(int64_t), "row pointers are uploaded as they are");
            E->d_row_ptr.alloc((size_t)n + 1);
            CK(cudaMemcpy(E->d_row_ptr.p, outdeg.data(), ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
            E->d_korig.alloc((size_t)nnz);
            CK(cudaMemcpy(E->d_korig.p, ix.ent.data(), (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice));
            E->d_dstnodes.alloc((size_t)nnz);
            CK(cudaMemcpy(E->d_dstnodes.p, dst, (size_t)nnz * sizeof(int64_t), cudaMemcpyHostToDevice));
            g_h2d_bytes += (long long)(((size_t)n + 1) * 8 + (size_t)nnz * 12);
        }
    });
    std::vector<int32_t> &row_cnt = W.row_cnt;
    row_cnt.assign(n, 0);
    count_task.start([&] {
        const int nthr = n < 4096 ? 1 : 4;
        if ((int)W.marks.size() < nthr) W.marks.resize(nthr);
        std::vector<std::thread> th;
        for (int t = 0; t < nthr; ++t)
            th.emplace_back([&, t] {
                std::vector<int32_t> &mark = W.marks[t];
                mark.assign(n, -1);
                for (int64_t u = n * t / nthr; u < n * (t + 1) / nthr; ++u) {
                    int32_t c = 0;
                    dedup_row(u, mark, [&](int32_t) { ++c; });
                    row_cnt[u] = c;
                }
            });
        for (auto &x : th) x.join();
    });
    // Nodes with no (undirected) path to A ∪ B cannot influence any result; like the reference's Python front end
    // (_kmc_rates.py:414-425) they are dropped before elimination.  Their committors read NaN.  Almost every network
    // is connected, so the reachability sweep runs on its own thread beside the analysis, which is only repeated on
    // the pruned node set when the sweep did find something to drop.
    std::vector<char> &reach = W.reach;
    reach.assign(n, 0);
    std::thread sweep([&] {
        std::vector<int32_t> &stack = W.stack;
        stack.clear();
        for (int64_t a : E->A) if (!reach[a]) { reach[a] = 1; stack.push_back((int32_t)a); }
        for (int64_t b : E->B) if (!reach[b]) { reach[b] = 1; stack.push_back((int32_t)b); }
        while (!stack.empty()) {
            const int32_t u = stack.back();
            stack.pop_back();
            for (int64_t e = g.ptr[u]; e < g.ptr[u + 1]; ++e) {
                const int32_t v = g.adj[e];
                if (!reach[v]) { reach[v] = 1; stack.push_back(v); }
            }
        }
    });
    try {
        ngt::analyse_sparse(g, present, E->A, E->B, E->opt.leaf_size, E->S, accel, sweep_min);
    } catch (...) {
        analysis_done.store(true, std::memory_order_release);
        sweep.join();
        throw;
    }
    analysis_done.store(true, std::memory_order_release);
    sweep.join();
    {
        bool pruned = false;
        for (int64_t u = 0; u < n; ++u)
            if (present[u] && !reach[u]) { present[u] = 0; pruned = true; }
        tm.lap("analysis (reachability sweep alongside)");
        if (pruned) {
            E->S = Symbolic();
            ngt::analyse_sparse(g, present, E->A, E->B, E->opt.leaf_size, E->S, accel, sweep_min);
            tm.lap("analysis repeated on the pruned node set");
        }
    }
    const Symbolic &S = E->S;

    // the analysed structure (fronts, assembly lists, batches, pools) goes to the device beside the slot-table pass
    device_task.finish();
    SideTask structure_task;
    structure_task.start([&] {
        CK(cudaSetDevice(E->device));
        E->upload_structure();
    });
    count_task.finish();
    {
        int64_t kept_entries = 0;
        for (int64_t u = 0; u < n; ++u) kept_entries += row_cnt[u];
        bool was_pruned = false;
        for (int64_t u = 0; u < n && !was_pruned; ++u) was_pruned = !present[u];
        if (all_present && !was_pruned && kept_entries == nnz) {
            // ---- rate -> slot table on the device (no duplicates, nothing absent or pruned: row u = node u) ----------
            structure_task.finish();
            tm.lap("structure upload + pools (joined)");
            CK(cudaSetDevice(E->device));
            std::vector<int> ffirst(S.fronts.size());
            for (size_t i = 0; i < S.fronts.size(); ++i) ffirst[i] = S.fronts[i].first;
            DevBuf<int> d_posn, d_pos_front, d_ffirst;
            d_posn.upload(S.pos);
            d_pos_front.upload(S.pos_front);
            d_ffirst.upload(ffirst);
            E->d_dest.alloc((size_t)nnz);
            E->d_tau_dest.alloc((size_t)n);
            E->d_row_node.release();                     // null: the ingest kernel reads row u as node u
            E->n_rows = (int)n;
            ngt::launch_slot_table((int)n, E->d_row_ptr.p, E->d_korig.p, E->d_dstnodes.p, d_posn.p, d_pos_front.p,
                                   E->d_fronts.p, d_ffirst.p, E->d_fidx.p, E->d_idx_off.p, E->d_dest.p, E->d_tau_dest.p,
                                   E->d_err.p, E->stream);
            CK(cudaGetLastError());
            E->check_errflag("slot table");              // synchronises the stream: the temporaries may go
            E->d_dstnodes.release();
            tm.lap("slot table on the device");
            E->st.ms_symbolic = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            return;
        }
    }
    // CSR by source over present nodes, (src,dst) duplicates resolved "last wins"; rows = present nodes in id order,
    // each worker thread writes the source indices and slots of a contiguous slice of rows
    std::vector<int> &row_node = W.row_node;
    row_node.clear();
    for (int64_t u = 0; u < n; ++u)
        if (present[u]) row_node.push_back((int)u);
    const int64_t nrows = (int64_t)row_node.size();
    std::vector<long long> &row_ptr = W.row_ptr, &tau_dest = W.tau_dest, &dest = W.dest;
    std::vector<int> &korig = W.korig;
    row_ptr.assign(nrows + 1, 0);
    tau_dest.resize(nrows);
    {
        unsigned hw = ngt::host_threads();
        int nthr = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
        if (nrows < 4096) nthr = 1;
        std::vector<std::vector<int32_t> > &marks = W.marks;
        if ((int)marks.size() < nthr) marks.resize(nthr);
        auto run = [&](auto &&body) {
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; ++t) th.emplace_back([&, t] { body(t, nrows * t / nthr, nrows * (t + 1) / nthr); });
            for (auto &x : th) x.join();
        };
        for (int64_t r = 0; r < nrows; ++r) row_ptr[r + 1] = row_ptr[r] + row_cnt[row_node[r]];
        korig.resize(row_ptr[nrows]);
        dest.resize(row_ptr[nrows]);
        std::vector<char> bad(nthr, 0);
        run([&](int t, int64_t r0, int64_t r1) {
            marks[t].assign(n, -1);
            for (int64_t r = r0; r < r1; ++r) {
                const int64_t u = row_node[r];
                const int32_t pu = S.pos[u];
                long long w = row_ptr[r];
                dedup_row(u, marks[t], [&](int32_t e) {
                    const int32_t pv = S.pos[dst[e]];
                    const int32_t owner = S.pos_front[std::min(pu, pv)];
                    const ngt::Front &F = S.fronts[owner];
                    const int32_t lr = ngt::local_index(S, owner, pu), lc = ngt::local_index(S, owner, pv);
                    if (lr < 0 || lc < 0) bad[t] = 1;
                    korig[w] = e;
                    dest[w] = F.off + (long long)lr * F.ld + lc;
                    ++w;
                });
                const int32_t owner = S.pos_front[pu];
                const ngt::Front &F = S.fronts[owner];
                tau_dest[r] = F.off + (long long)ngt::local_index(S, owner, pu) * F.ld + F.f;
            }
        });
        for (char b : bad)
            if (b) throw std::logic_error("ingest: entry has no slot in its front");
    }
    E->n_rows = (int)row_node.size();
    tm.lap("CSR + slot table");
    CK(cudaSetDevice(E->device));
    E->d_row_ptr.upload(row_ptr);
    E->d_row_node.upload(row_node);
    E->d_korig.upload(korig);
    E->d_dest.upload(dest);
    E->d_tau_dest.upload(tau_dest);
    tm.lap("upload ingest tables");
    structure_task.finish();
    tm.lap("structure upload + pools (joined)");
    E->st.ms_symbolic = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

void build_dense(Engine *E, int64_t n, const double *K, bool on_device, int64_t nA, const int64_t *A, int64_t nB,
                 const int64_t *B) {
    auto t0 = std::chrono::steady_clock::now();
    if (n <= 1) throw InvalidError("empty network");
    if (n > 1000000) throw InvalidError("dense network too large");
    E->n = n;
    E->nnz = n * n;
    E->dense = true;
    E->A.assign(A, A + nA);
    E->B.assign(B, B + nB);
    E->validate_groups();
    ngt::analyse_dense(n, E->A, E->B, E->S);
    E->init_device();
    E->d_pos.upload(E->S.pos);
    E->d_node_at.upload(E->S.node_at);
    E->upload_structure();
    E->make_tensor_maps();
    if (K) {
        if (on_device) {
            // the caller's producer (a torch kernel on any stream) must have finished before the engine's own
            // non-blocking streams read K
            CK(cudaDeviceSynchronize());
            E->d_K = K;
        } else {
            E->d_K_owned.alloc((size_t)n * n * E->nbatch);
            CK(cudaMemcpy(E->d_K_owned.p, K, (size_t)n * n * E->nbatch * sizeof(double), cudaMemcpyHostToDevice));
            g_h2d_bytes += (long long)((size_t)n * n * E->nbatch * sizeof(double));
            E->d_K = E->d_K_owned.p;
        }
        E->values_loaded = true;
    }
    E->st.ms_symbolic = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

template <class Fn>
int guard(Fn &&fn) {
    try {
        fn();
        return NGT_OK;
    } catch (const InvalidError &e) {
        g_err = e.what();
        return NGT_E_INVALID;
    } catch (const CudaError &e) {
        g_err = e.what();
        return NGT_E_CUDA;
    } catch (const StateError &e) {
        g_err = e.what();
        return NGT_E_STATE;
    } catch (const NumericError &e) {
        g_err = e.what();
        return NGT_E_NUMERIC;
    } catch (const ngt::LaunchFailure &e) {
        g_err = e.what();
        return NGT_E_CUDA;
    } catch (const std::exception &e) {
        g_err = e.what();
        return NGT_E_INVALID;
    }
}

Engine *eng(ngt_handle h) {
    if (!h) throw InvalidError("null handle");
    return reinterpret_cast<Engine *>(h);
}
// getters that return one network's worth of data refuse batched handles instead of silently answering for network 0
Engine *eng_single(ngt_handle h, const char *what) {
    Engine *E = eng(h);
    if (E->nbatch != 1)
        throw InvalidError(std::string(what) + " returns one network; this handle holds a batch (use the ngt_batch_* getters)");
    return E;
}

}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char *ngt_last_error(void) { return g_err.c_str(); }
const char *ngt_version(void) { return "kmc_rates_b200 0.1 (sm_100a)"; }

int ngt_batch_create(int64_t nbatch, int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, const double *k,
                     int64_t nA, const int64_t *A, int64_t nB, const int64_t *B, const ngt_options *opt,
                     ngt_handle *out) {
    return guard([&] {
        if (!out) throw InvalidError("null output handle");
        *out = nullptr;
        if (!src || !dst || !A || !B) throw InvalidError("null input pointer");
        Engine *E = make_engine(opt, nbatch);
        try {
            build_sparse(E, n, nnz, src, dst, k, nA, A, nB, B);
        } catch (...) {
            delete E;
            throw;
        }
        *out = reinterpret_cast<ngt_handle>(E);
    });
}

int ngt_create(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, const double *k, int64_t nA,
               const int64_t *A, int64_t nB, const int64_t *B, const ngt_options *opt, ngt_handle *out) {
    return ngt_batch_create(1, n, nnz, src, dst, k, nA, A, nB, B, opt, out);
}

static int create_dense_any(int64_t nbatch, int64_t n, const double *K, bool dev, int64_t nA, const int64_t *A,
                            int64_t nB, const int64_t *B, const ngt_options *opt, ngt_handle *out) {
    return guard([&] {
        if (!out) throw InvalidError("null output handle");
        *out = nullptr;
        if (!A || !B) throw InvalidError("null input pointer");
        Engine *E = make_engine(opt, nbatch);
        try {
            build_dense(E, n, K, dev, nA, A, nB, B);
        } catch (...) {
            delete E;
            throw;
        }
        *out = reinterpret_cast<ngt_handle>(E);
    });
}

int ngt_create_dense(int64_t n, const double *K, int64_t nA, const int64_t *A, int64_t nB, const int64_t *B,
                     const ngt_options *opt, ngt_handle *out) {
    return create_dense_any(1, n, K, false, nA, A, nB, B, opt, out);
}
int ngt_create_dense_dev(int64_t n, const double *K, int64_t nA, const int64_t *A, int64_t nB, const int64_t *B,
                         const ngt_options *opt, ngt_handle *out) {
    return create_dense_any(1, n, K, true, nA, A, nB, B, opt, out);
}
int ngt_batch_create_dense(int64_t nbatch, int64_t n, const double *K, int64_t nA, const int64_t *A, int64_t nB,
                           const int64_t *B, const ngt_options *opt, ngt_handle *out) {
    return create_dense_any(nbatch, n, K, false, nA, A, nB, B, opt, out);
}
int ngt_batch_create_dense_dev(int64_t nbatch, int64_t n, const double *K, int64_t nA, const int64_t *A, int64_t nB,
                               const int64_t *B, const ngt_options *opt, ngt_handle *out) {
    return create_dense_any(nbatch, n, K, true, nA, A, nB, B, opt, out);
}

static int set_rates_any(ngt_handle h, const double *k, bool dev) {
    return guard([&] {
        Engine *E = eng(h);
        if (!k) throw InvalidError("null rate pointer");
        CK(cudaSetDevice(E->device));
        const size_t cnt = (size_t)E->nnz * E->nbatch;
        // device pointers: whatever produced k (on any stream of the caller) must be complete before the engine's
        // non-blocking streams touch it; the engine's own previous step must be done before its inputs change
        if (dev) CK(cudaDeviceSynchronize());
        else CK(cudaStreamSynchronize(E->stream));
        if (E->dense) {
            if (dev) {
                E->d_K = k;
            } else {
                if (!E->d_K_owned.p) E->d_K_owned.alloc(cnt);
                CK(cudaMemcpyAsync(E->d_K_owned.p, k, cnt * sizeof(double), cudaMemcpyHostToDevice, E->stream));
                CK(cudaStreamSynchronize(E->stream));
                g_h2d_bytes += (long long)(cnt * sizeof(double));
                E->d_K = E->d_K_owned.p;
            }
        } else {
            CK(cudaMemcpyAsync(E->d_k.p, k, cnt * sizeof(double), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               E->stream));
            CK(cudaStreamSynchronize(E->stream));
            if (!dev) g_h2d_bytes += (long long)(cnt * sizeof(double));
        }
        E->values_loaded = true;
        E->have_phase1 = E->have_phase2 = E->have_backsub = false;
    });
}
int ngt_set_rates(ngt_handle h, const double *k) { return set_rates_any(h, k, false); }
int ngt_set_rates_dev(ngt_handle h, const double *k) { return set_rates_any(h, k, true); }

int ngt_set_weights(ngt_handle h, int64_t nw, const int64_t *ids, const double *w) {
    return guard([&] {
        Engine *E = eng(h);
        if (nw > 0 && (!ids || !w)) throw InvalidError("null weight pointer");
        if (!E->have_weights) E->weights.assign(E->n, 1.0);
        E->have_weights = true;
        for (int64_t i = 0; i < nw; ++i)
            if (ids[i] >= 0 && ids[i] < E->n) E->weights[ids[i]] = w[i];   // unknown ids ignored, as ngt.pyx:99-103
        if (E->have_phase2) E->finish_rates();
    });
}

void ngt_destroy(ngt_handle h) {
    if (!h) return;
    Engine *E = reinterpret_cast<Engine *>(h);
    cudaSetDevice(E->device);
    delete E;
}

int ngt_phase_one(ngt_handle h) {
    return guard([&] {
        Engine *E = eng(h);
        E->phase_one();
        CK(cudaStreamSynchronize(E->stream));
        E->check_errflag("phase one");
        float ms = 0;
        cudaEventElapsedTime(&ms, E->ev[0], E->ev[1]); E->st.ms_ingest = ms;
        cudaEventElapsedTime(&ms, E->ev[1], E->ev[2]); E->st.ms_phase1 = ms;
    });
}

int ngt_phase_two(ngt_handle h) {
    return guard([&] { eng(h)->phase_two(); });
}

int ngt_phase_two_shard(ngt_handle h, int32_t rank, int32_t nranks) {
    return guard([&] {
        if (nranks < 1 || rank < 0 || rank >= nranks) throw InvalidError("bad rank / nranks");
        eng(h)->phase_two(rank, nranks);
    });
}

int ngt_compute_rates(ngt_handle h) {
    return guard([&] {
        Engine *E = eng(h);
        E->step_enqueue();
        E->phase_two_finish();
    });
}

int ngt_compute_rates_and_committors(ngt_handle h) {
    return guard([&] {
        Engine *E = eng(h);
        E->step_enqueue();
        E->phase_two_finish();
        E->backsub();
    });
}

int ngt_get_rates(ngt_handle h, double out4[4]) {
    return guard([&] {
        Engine *E = eng_single(h, "ngt_get_rates");
        if (!E->have_phase2 || E->rates.empty()) throw StateError("compute_rates has not been called");
        for (int i = 0; i < 4; ++i) out4[i] = E->rates[i];
    });
}

int ngt_batch_get_rates(ngt_handle h, double *out) {
    return guard([&] {
        Engine *E = eng(h);
        if (!E->have_phase2 || E->rates.empty()) throw StateError("compute_rates has not been called");
        std::memcpy(out, E->rates.data(), E->rates.size() * sizeof(double));
    });
}

int ngt_get_final(ngt_handle h, double *omPxx, double *tau) {
    return guard([&] {
        Engine *E = eng(h);
        if (!E->have_phase2) throw StateError("compute_rates has not been called");
        std::memcpy(omPxx, E->final_om.data(), (size_t)E->m * E->nbatch * sizeof(double));
        std::memcpy(tau, E->final_tau.data(), (size_t)E->m * E->nbatch * sizeof(double));
    });
}

int ngt_set_final(ngt_handle h, const double *omPxx, const double *tau) {
    return guard([&] {
        Engine *E = eng(h);
        if (!E->have_phase2) throw StateError("phase two has not been run");
        std::memcpy(E->final_om.data(), omPxx, (size_t)E->m * E->nbatch * sizeof(double));
        std::memcpy(E->final_tau.data(), tau, (size_t)E->m * E->nbatch * sizeof(double));
        E->finish_rates();
    });
}

int ngt_final_dev(ngt_handle h, double **om, double **tau, int64_t *ndoubles) {
    return guard([&] {
        Engine *E = eng(h);
        if (!E->have_phase2) throw StateError("phase two has not been run");
        *om = E->d_om.p;
        *tau = E->d_ftau.p;
        *ndoubles = (int64_t)E->m * E->nbatch;
    });
}

int ngt_adopt_final_dev(ngt_handle h) {
    return guard([&] {
        Engine *E = eng(h);
        if (!E->have_phase2) throw StateError("phase two has not been run");
        CK(cudaSetDevice(E->device));
        CK(cudaDeviceSynchronize());     // the caller's collective ran on a stream of its own
        CK(cudaMemcpy(E->final_om.data(), E->d_om.p, E->final_om.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(E->final_tau.data(), E->d_ftau.p, E->final_tau.size() * sizeof(double), cudaMemcpyDeviceToHost));
        E->finish_rates();
    });
}

int ngt_get_reduced(ngt_handle h, double *P, double *tau) {
    return guard([&] {
        Engine *E = eng_single(h, "ngt_get_reduced");
        if (!E->have_phase1) throw StateError("phase one has not been run");
        if (!E->have_phase2) {
            E->download_root();
            CK(cudaStreamSynchronize(E->stream));
        }
        for (int i = 0; i < E->m; ++i) {
            for (int j = 0; j < E->m; ++j) P[(size_t)i * E->m + j] = E->root[(size_t)i * E->root_ld + j];
            tau[i] = E->root[(size_t)i * E->root_ld + E->m];
        }
    });
}

int ngt_get_initial_tau(ngt_handle h, double *tau) {
    return guard([&] {
        Engine *E = eng_single(h, "ngt_get_initial_tau");
        if (!E->have_phase1) throw StateError("phase one has not been run");
        if (E->tau0.empty()) {
            E->download_root();
            CK(cudaStreamSynchronize(E->stream));
        }
        std::memcpy(tau, E->tau0.data(), (size_t)E->n * sizeof(double));
    });
}

int ngt_get_elimination_order(ngt_handle h, int64_t *nodes, int64_t cap, int64_t *count) {
    return guard([&] {
        Engine *E = eng(h);
        const std::vector<int32_t> &na = E->S.node_at;
        if (count) *count = (int64_t)na.size();
        if (nodes)
            for (int64_t i = 0; i < cap && i < (int64_t)na.size(); ++i) nodes[i] = na[i];
    });
}

static int get_x(ngt_handle h, double *out, int which) {
    return guard([&] {
        Engine *E = eng_single(h, which == 0 ? "ngt_get_committors" : "ngt_get_mfpt");
        if (!E->have_backsub) E->backsub();
        const double nan = std::nan("");
        for (int64_t u = 0; u < E->n; ++u) {
            const int32_t p = E->S.pos[u];
            out[u] = p < 0 ? nan : E->xq[(size_t)p * 2 + which];
        }
    });
}
int ngt_get_committors(ngt_handle h, double *q) { return get_x(h, q, 0); }
int ngt_get_mfpt(ngt_handle h, double *t) { return get_x(h, t, 1); }

int ngt_get_stats(ngt_handle h, ngt_stats *st) {
    return guard([&] {
        Engine *E = eng(h);
        *st = E->st;
        st->reserved[2] = (double)(g_h2d_bytes.load() + ngt::gpu_sweeps_h2d_bytes());   // process-wide byte counters
        st->reserved[3] = (double)(g_d2h_bytes.load() + ngt::gpu_sweeps_d2h_bytes());   // (differences are what matter)
    });
}

int ngt_timed_steps(ngt_handle h, int32_t steps, double *ms_total) {
    return guard([&] {
        Engine *E = eng(h);
        if (steps < 1) throw InvalidError("steps must be >= 1");
        CK(cudaSetDevice(E->device));
        CK(cudaStreamSynchronize(E->stream));
        CK(cudaEventRecord(E->ev[5], E->stream));
        for (int i = 0; i < steps; ++i) E->step_enqueue();
        cudaEvent_t stop = E->prof_event();
        CK(cudaEventRecord(stop, E->stream));
        CK(cudaStreamSynchronize(E->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, E->ev[5], stop));
        *ms_total = ms;
        E->phase_two_finish();
    });
}

int ngt_profile_schur(ngt_handle h, double *ms_schur, double *flops, int64_t *launches) {
    // reserved[1] of ngt_stats receives the duration of the whole profiled phase-1 pass (same pass, same event
    // overheads), so that the kernel's share of the step can be formed from consistent numbers
    return guard([&] {
        Engine *E = eng(h);
        CK(cudaSetDevice(E->device));
        // one stream for this pass: with the trailing updates on the look-ahead stream, event pairs on two streams
        // overlap in time and their sum could exceed the duration of the pass
        const bool la = E->use_lookahead;
        E->profile_schur = true;
        E->use_lookahead = false;
        try {
            E->phase_one();
        } catch (...) {
            E->profile_schur = false;
            E->use_lookahead = la;
            throw;
        }
        E->profile_schur = false;
        E->use_lookahead = la;
        CK(cudaStreamSynchronize(E->stream));
        E->check_errflag("phase one");
        double tot = 0;
        for (size_t i = 0; i + 1 < E->prof_used; i += 2) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, E->prof_ev[i], E->prof_ev[i + 1]));
            tot += ms;
        }
        E->st.ms_schur = tot;
        float pass_ms = 0;
        CK(cudaEventElapsedTime(&pass_ms, E->ev[0], E->ev[2]));
        E->st.reserved[1] = pass_ms;
        *ms_schur = tot;
        *flops = E->st.schur_flops;
        *launches = (int64_t)(E->prof_used / 2);
    });
}

void ngt_trim_device_cache(void) {
    g_cache.trim();
    ngt::release_gpu_sweeps();
    trim_host_workspace();
}

int ngt_dist_begin(ngt_handle h, int32_t rank, int32_t world) {
    return guard([&] { eng(h)->dist_begin(rank, world); });
}
int ngt_dist_rowsum(ngt_handle h, int64_t J, double **dptr, int64_t *ndoubles) {
    return guard([&] {
        *dptr = eng(h)->dist_rowsum((int)J);
        *ndoubles = ngt::NB_OUTER;
    });
}
int ngt_dist_factor(ngt_handle h, int64_t J) {
    return guard([&] { eng(h)->dist_factor((int)J); });
}
int ngt_dist_panel(ngt_handle h, int64_t J, double **dptr, int64_t *ndoubles) {
    return guard([&] { eng(h)->dist_panel((int)J, dptr, ndoubles); });
}
int ngt_dist_apply(ngt_handle h, int64_t J) {
    return guard([&] { eng(h)->dist_apply((int)J); });
}
int ngt_dist_root(ngt_handle h, double **dptr, int64_t *ndoubles) {
    return guard([&] {
        Engine *E = eng(h);
        *dptr = E->dist_root_contrib();
        *ndoubles = (int64_t)E->m * E->m;
    });
}
int ngt_dist_finish(ngt_handle h) {
    return guard([&] { eng(h)->dist_finish(); });
}
int ngt_dist_info(ngt_handle h, int64_t *n_pivots, int64_t *block) {
    return guard([&] {
        *n_pivots = eng(h)->S.nI;
        *block = ngt::NB_OUTER;
    });
}

int ngt_level_sweep(int32_t device, int32_t n, const int32_t *ptr, const int32_t *adj, int32_t start, int32_t *lvl,
                    int32_t *visit, int32_t *nlev, int32_t *handled) {
    return guard([&] {
        if (n < 1 || !ptr || !adj || !lvl || !visit || !nlev || !handled || start >= n) throw InvalidError("bad argument");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CudaError("no CUDA device — kmc_rates_b200 has no CPU fallback");
        if (device < 0 || device >= ndev) throw InvalidError("device ordinal out of range");
        ngt::SweepAccel *accel = ngt::gpu_sweeps(device);
        *handled = accel && accel->sweep(n, ptr, adj, (int64_t)ptr[n], start, lvl, visit, nlev) == 1 ? 1 : 0;
    });
}

int ngt_connected_components(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, int32_t device,
                             int32_t *labels, int64_t *ncomp) {
    return guard([&] {
        if (n < 0 || nnz < 0 || (nnz > 0 && (!src || !dst)) || (n > 0 && !labels)) throw InvalidError("bad argument");
        if (n > 2000000000LL) throw InvalidError("too many nodes (32-bit labels)");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CudaError("no CUDA device — kmc_rates_b200 has no CPU fallback");
        if (device < 0 || device >= ndev) throw InvalidError("device ordinal out of range");
        CK(cudaSetDevice(device));
        DevBuf<int> parent, err;
        DevBuf<long long> ds, dd;
        parent.alloc((size_t)std::max<int64_t>(n, 1));
        err.alloc(1);
        CK(cudaMemset(err.p, 0, sizeof(int)));
        ds.alloc((size_t)std::max<int64_t>(nnz, 1));
        dd.alloc((size_t)std::max<int64_t>(nnz, 1));
        static_assert(sizeof(long long) == sizeof(int64_t), "ids are uploaded as they are");
        if (nnz > 0) {
            CK(cudaMemcpy(ds.p, src, (size_t)nnz * sizeof(int64_t), cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dd.p, dst, (size_t)nnz * sizeof(int64_t), cudaMemcpyHostToDevice));
            g_h2d_bytes += (long long)nnz * 16;
        }
        ngt::launch_connected_components(parent.p, ds.p, dd.p, nnz, (int)n, err.p, 0);
        CK(cudaGetLastError());
        int flag = 0;
        CK(cudaMemcpy(&flag, err.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) throw InvalidError("rate endpoint out of range");
        if (n > 0) CK(cudaMemcpy(labels, parent.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        g_d2h_bytes += (long long)n * 4;
        if (ncomp) {
            int64_t c = 0;
            for (int64_t i = 0; i < n; ++i) c += labels[i] == (int32_t)i;
            *ncomp = c;
        }
    });
}

// ---- NCCL communicator + the sharded paths driven from C++ ----------------------------------------------------
int ngt_comm_unique_id(uint8_t id[128]) {
    return guard([&] {
        if (!id) throw InvalidError("null id buffer");
        NcclApi::unique_id u;
        NK(NcclApi::get().GetUniqueId(&u));
        std::memcpy(id, u.internal, 128);
    });
}

int ngt_comm_create(int32_t device, const uint8_t id[128], int32_t rank, int32_t world, ngt_comm *out) {
    return guard([&] {
        if (!out || !id) throw InvalidError("null pointer");
        *out = nullptr;
        if (world < 1 || rank < 0 || rank >= world) throw InvalidError("bad rank / world size");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError("no CUDA device — kmc_rates_b200 has no CPU fallback");
        if (device < 0 || device >= ndev) throw InvalidError("device ordinal out of range");
        CK(cudaSetDevice(device));
        std::unique_ptr<Comm> c(new Comm());
        c->rank = rank; c->world = world; c->device = device;
        NcclApi::unique_id u;
        std::memcpy(u.internal, id, 128);
        NK(NcclApi::get().CommInitRank(&c->comm, world, u, rank));
        if (world > 1 && NcclApi::get().CommSplit) NK(NcclApi::get().CommSplit(c->comm, 0, rank, &c->comm2, nullptr));
        *out = reinterpret_cast<ngt_comm>(c.release());
    });
}

void ngt_comm_destroy(ngt_comm c) {
    if (!c) return;
    Comm *cm = reinterpret_cast<Comm *>(c);
    cudaSetDevice(cm->device);
    delete cm;
}

int ngt_dist_phase_one(ngt_handle h, ngt_comm c, double *ms_device) {
    return guard([&] {
        if (!c) throw InvalidError("null communicator");
        eng(h)->dist_phase_one(*reinterpret_cast<Comm *>(c), ms_device);
    });
}

int ngt_phase_two_sharded(ngt_handle h, ngt_comm c, int32_t broadcast_root, double *ms_device) {
    return guard([&] {
        if (!c) throw InvalidError("null communicator");
        eng(h)->phase_two_sharded(*reinterpret_cast<Comm *>(c), broadcast_root != 0, ms_device);
    });
}

int ngt_reduced_block_dev(ngt_handle h, double **dptr, int64_t *ndoubles) {
    return guard([&] {
        Engine *E = eng(h);
        *dptr = E->d_pool.p + E->S.fronts[E->S.root_id].off;
        *ndoubles = (int64_t)E->m * E->root_ld;
    });
}

}  // extern "C"

// Host-side symbolic analysis for the multifrontal NGT engine.
//
// The reference keeps a pointer graph and discovers fill-in edge by edge while it eliminates
// (reference source/kmc_rates/graph.hpp:222-236, ngt.hpp:271-276) and re-sorts the node list by degree
// every step (ngt.hpp:192-200).  Here the whole elimination is planned up front:
//   * a nested-dissection ordering of the intermediate nodes (A ∪ B are ordered last and never
//     eliminated in phase 1),
//   * supernodes = the separators and leaves of the dissection tree,
//   * one dense "front" per supernode: its k pivots followed by the m later nodes it touches,
//   * an assembly tree (parent = front owning the earliest node of the update set), levels and
//     batches of fronts that are eliminated by the same kernel launches.
// No arithmetic happens here; the numeric phase never allocates.
#pragma once
#include <chrono>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

namespace ngt {

// section timer, printed to stderr when NGT_B200_TIMING is set
struct Timer {
    bool on;
    std::chrono::steady_clock::time_point t;
    Timer() : on(std::getenv("NGT_B200_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ngt timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

struct Front {
    int32_t k;        // pivots eliminated in this front
    int32_t m;        // update-set size (later nodes touched)
    int32_t f;        // k + m
    int32_t ld;       // leading dimension in doubles (>= f + 1; column f carries tau)
    int64_t off;      // offset of the front's matrix in the pool (doubles)
    int64_t idx_off;  // offset of its index list (elimination positions; pivots then sorted update set)
    int32_t level;    // 0 = leaf of the assembly tree
    int32_t parent;   // front id, or -1 (no later node touched: component not connected to A ∪ B)
    int32_t first;    // elimination position of its first pivot
    int32_t pad;
};

struct Batch {
    int32_t level;
    int32_t max_k, max_f;
    std::vector<int32_t> fronts;   // front ids processed by the same launches
};

struct Symbolic {
    int64_t n = 0;          // nodes (dense ids)
    int64_t nI = 0;         // intermediates = nodes eliminated in phase 1
    int64_t nA = 0, nB = 0;
    std::vector<int32_t> pos;       // node id -> elimination position (A then B occupy nI..n-1); -1 = node absent
    std::vector<int32_t> node_at;   // elimination position -> node id
    std::vector<Front> fronts;      // fronts[root_id] is the root (k = 0, index list = A ∪ B)
    int32_t root_id = -1;
    std::vector<int32_t> fidx;      // concatenated index lists
    std::vector<int64_t> child_ptr; // CSR of children per front
    std::vector<int32_t> child_ids;
    std::vector<int64_t> rel_off;   // per front: offset into rel (m entries): local index in the parent
    std::vector<int32_t> rel;
    std::vector<int32_t> pos_front; // elimination position -> owning front (pivot owner, root for A ∪ B)
    std::vector<Batch> batches;     // in execution order (by level)
    int64_t pool_doubles = 0;
    int32_t n_levels = 0;
    int32_t max_front = 0;
    double alg_flops = 0, alg_bytes = 0, schur_flops = 0;
};

// Worker threads the host analysis may use: hardware concurrency, capped by the environment variable
// NGT_B200_HOST_THREADS (several ranks sharing one host set it to cores / ranks).
unsigned host_threads();

// Undirected structure in CSR (no self loops, deduplicated, symmetrised).
struct Adjacency {
    int64_t n = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> adj;
};

// Build the symmetrised adjacency of the rate graph.
void build_adjacency(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, Adjacency &g);

// One parallel pass structure over the COO input (thread-private counting sort): range validation, presence flags,
// CSR-by-source entry lists (stable: increasing entry index inside a row) and the symmetrised, de-duplicated adjacency.
struct EdgeIndex {
    std::vector<int64_t> outptr;   // n + 1
    std::vector<int32_t> ent;      // nnz entry indices grouped by source node
    std::vector<char> present;     // node appears in some rate
    bool range_error = false;      // an endpoint was outside [0, n)
};
// Temporaries of index_edges.  A caller that analyses many networks passes the same object again: on a fresh
// allocation the page faults of these ~100 MB cost more than the indexing itself (measured 34 -> 18 ms for 2*10^6 rates).
struct IndexScratch {
    std::vector<std::vector<int32_t> > c_out, c_adj, seen;
    std::vector<std::vector<int64_t> > w_out, w_adj;
    std::vector<int32_t> tmp;
    std::vector<int64_t> aptr, ucnt;
};
void index_edges(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, EdgeIndex &ix, Adjacency &g,
                 IndexScratch *ws = nullptr);

// Optional accelerator for the level-structure sweeps of large subgraphs (the engine installs a CUDA kernel,
// ordering.cu).  A sweep is a breadth-first search that reports the serial queue order exactly: level L + 1 lists the
// nodes in the order (position of the first discovering parent in level L, position in that parent's row), so the
// ordering does not depend on who sweeps.
struct SweepAccel {
    virtual ~SweepAccel() = default;
    // CSR subgraph (local ids, symmetric, no self loops, rows without duplicates).
    // start >= 0: one sweep from `start`.  start < 0: pseudo-peripheral search -- sweep from node 0, pick the
    // minimum-degree node of the last level (the latest visited among equals) and, unless that is node 0, sweep again
    // from it.  Fills lvl[n], visit[n], *nlev and returns 1.  Otherwise the caller sweeps on the host: 0 = not now (no
    // free context), -1 = not this graph (size limits, not connected, levels too thin to pay off, device error).
    virtual int sweep(int32_t n, const int32_t *ptr, const int32_t *adj, int64_t nadj, int32_t start, int32_t *lvl,
                       int32_t *visit, int32_t *nlev) = 0;
    int32_t min_nodes = 65536;   // smaller subgraphs are swept on the host: a level costs the kernel ~10 us whatever its
                                 // size, a host core ~2.6 ns per adjacency entry (measured break-even ~7e4 nodes of the
                                 // 2-D landscapes); NGT_B200_GPU_SWEEP_MIN or ngt_options.sweep_min_nodes override it
    std::atomic<int> uploads{0}; // subgraphs copied to the accelerator so far (the engine orders its own copies after the first)
};

// Sparse network: nested dissection + supernodal symbolic factorisation.
// present[u] = 0 marks ids that appear in no rate (they are ignored, like the reference which only
// creates nodes named by a rate, ngt.hpp:114-117).
void analyse_sparse(const Adjacency &g, const std::vector<char> &present,
                    const std::vector<int64_t> &A, const std::vector<int64_t> &B,
                    int leaf_size, Symbolic &S, SweepAccel *accel = nullptr, int32_t accel_min_nodes = 0);

// Complete network: one front holding every intermediate, plus the root.
void analyse_dense(int64_t n, const std::vector<int64_t> &A, const std::vector<int64_t> &B, Symbolic &S);

// Local index of elimination position p inside front fr (pivot range or binary search of the update set); -1 if absent.
int32_t local_index(const Symbolic &S, int32_t fr, int32_t p);

}  // namespace ngt

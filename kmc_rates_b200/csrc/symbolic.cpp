// Host-side symbolic analysis: ordering, supernodes, fronts, assembly tree.  See symbolic.h.
#include "symbolic.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif
#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace ngt {

void build_adjacency(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, Adjacency &g) {
    g.n = n;
    g.ptr.assign(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) {
        if (src[e] == dst[e]) continue;
        g.ptr[src[e] + 1]++;
        g.ptr[dst[e] + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
    std::vector<int32_t> tmp(g.ptr[n]);
    std::vector<int64_t> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (int64_t e = 0; e < nnz; ++e) {
        if (src[e] == dst[e]) continue;
        tmp[fill[src[e]]++] = (int32_t)dst[e];
        tmp[fill[dst[e]]++] = (int32_t)src[e];
    }
    // dedupe each row with a marker array (order inside a row is irrelevant), compact in place
    std::vector<int64_t> nptr(n + 1, 0);
    std::vector<int32_t> seen(n, -1);
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t b = g.ptr[i], e = g.ptr[i + 1];
        nptr[i] = w;
        for (int64_t q = b; q < e; ++q) {
            const int32_t v = tmp[q];
            if (seen[v] == (int32_t)i) continue;
            seen[v] = (int32_t)i;
            tmp[w++] = v;   // w <= q, safe in place
        }
    }
    nptr[n] = w;
    tmp.resize(w);
    g.ptr.swap(nptr);
    g.adj.swap(tmp);
}

// host worker threads available to the analysis: hardware concurrency, optionally capped by NGT_B200_HOST_THREADS
// (1 = fully serial; the results do not depend on it, which tests/test_host_analysis.py checks)
unsigned host_threads() {
    static const unsigned n = [] {
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        if (const char *e = std::getenv("NGT_B200_HOST_THREADS")) hw = std::min(hw, (unsigned)std::max(1, std::atoi(e)));
        return hw;
    }();
    return n;
}
namespace {
template <class Fn>
void parallel_ranges(int nthr, int64_t total, Fn &&fn) {
    if (nthr <= 1) { fn(0, (int64_t)0, total); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; ++t) th.emplace_back([&, t] { fn(t, total * t / nthr, total * (t + 1) / nthr); });
    for (auto &x : th) x.join();
}
}  // namespace

void index_edges(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, EdgeIndex &ix, Adjacency &g,
                 IndexScratch *ws) {
    unsigned hw = host_threads();
    int T = (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
    while (T > 1 && (double)T * n * 8.0 > 512e6) T /= 2;      // private counters: 8 bytes per node and thread
    if (nnz < 200000) T = 1;
    Timer tm;
    IndexScratch local;
    if (!ws) ws = &local;
    std::vector<std::vector<int32_t> > &c_out = ws->c_out, &c_adj = ws->c_adj, &seen_pool = ws->seen;
    std::vector<std::vector<int64_t> > &w_out = ws->w_out, &w_adj = ws->w_adj;
    std::vector<int32_t> &tmp = ws->tmp;
    std::vector<int64_t> &aptr = ws->aptr, &ucnt = ws->ucnt;
    if ((int)c_out.size() < T) { c_out.resize(T); c_adj.resize(T); seen_pool.resize(T); w_out.resize(T); w_adj.resize(T); }
    std::vector<char> bad(T, 0);
    // pass A: private counts per thread (each thread owns a contiguous slice of the entries)
    parallel_ranges(T, nnz, [&](int t, int64_t e0, int64_t e1) {
        c_out[t].assign(n, 0);
        c_adj[t].assign(n, 0);
        for (int64_t e = e0; e < e1; ++e) {
            const int64_t u = src[e], v = dst[e];
            if (u < 0 || u >= n || v < 0 || v >= n) { bad[t] = 1; continue; }
            c_out[t][u]++;
            if (u != v) { c_adj[t][u]++; c_adj[t][v]++; }
        }
    });
    ix.range_error = false;
    for (char b : bad) ix.range_error = ix.range_error || b;
    if (ix.range_error) return;
    tm.lap("  index: count");
    // totals -> row pointers (and presence: a node is present iff some rate names it); private counts -> private
    // write cursors
    ix.outptr.assign(n + 1, 0);
    aptr.assign(n + 1, 0);
    ix.present.assign(n, 0);
    parallel_ranges(T, n, [&](int /*tid*/, int64_t u0, int64_t u1) {
        for (int64_t u = u0; u < u1; ++u) {
            int64_t so = 0, sa = 0;
            for (int t = 0; t < T; ++t) { so += c_out[t][u]; sa += c_adj[t][u]; }
            ix.outptr[u + 1] = so;
            aptr[u + 1] = sa;
            ix.present[u] = (so + sa) > 0;
        }
    });
    for (int64_t u = 0; u < n; ++u) {
        ix.outptr[u + 1] += ix.outptr[u];
        aptr[u + 1] += aptr[u];
    }
    for (int t = 0; t < T; ++t) { w_out[t].resize(n); w_adj[t].resize(n); }
    parallel_ranges(T, n, [&](int /*tid*/, int64_t u0, int64_t u1) {
        for (int64_t u = u0; u < u1; ++u) {
            int64_t po = ix.outptr[u], pa = aptr[u];
            for (int t = 0; t < T; ++t) {
                w_out[t][u] = po; po += c_out[t][u];
                w_adj[t][u] = pa; pa += c_adj[t][u];
            }
        }
    });
    tm.lap("  index: offsets");
    // pass B: scatter
    ix.ent.resize(nnz);
    tmp.resize(aptr[n]);
    parallel_ranges(T, nnz, [&](int t, int64_t e0, int64_t e1) {
        std::vector<int64_t> &wo = w_out[t], &wa = w_adj[t];
        for (int64_t e = e0; e < e1; ++e) {
            const int64_t u = src[e], v = dst[e];
            ix.ent[wo[u]++] = (int32_t)e;
            if (u != v) { tmp[wa[u]++] = (int32_t)v; tmp[wa[v]++] = (int32_t)u; }
        }
    });
    tm.lap("  index: scatter");
    // pass C: de-duplicate adjacency rows (marker per thread), then compact
    ucnt.assign(n + 1, 0);
    parallel_ranges(T, n, [&](int t, int64_t u0, int64_t u1) {
        std::vector<int32_t> &seen = seen_pool[t];
        seen.assign(n, -1);
        for (int64_t u = u0; u < u1; ++u) {
            int64_t w = aptr[u];
            for (int64_t q = aptr[u]; q < aptr[u + 1]; ++q) {
                const int32_t v = tmp[q];
                if (seen[v] == (int32_t)u) continue;
                seen[v] = (int32_t)u;
                tmp[w++] = v;          // in place inside the row
            }
            ucnt[u + 1] = w - aptr[u];
        }
    });
    for (int64_t u = 0; u < n; ++u) ucnt[u + 1] += ucnt[u];
    g.n = n;
    g.ptr.assign(ucnt.begin(), ucnt.end());
    g.adj.resize(ucnt[n]);
    parallel_ranges(T, n, [&](int /*t*/, int64_t u0, int64_t u1) {
        for (int64_t u = u0; u < u1; ++u)
            std::copy(tmp.begin() + aptr[u], tmp.begin() + aptr[u] + (ucnt[u + 1] - ucnt[u]), g.adj.begin() + ucnt[u]);
    });
    tm.lap("  index: dedupe + compact");
}

namespace {

// ---------------------------------------------------------------------------------------------
// Nested dissection by BFS level structures (George's automatic nested dissection) on the graph
// induced on the intermediate nodes.  Emits nodes in elimination order and the supernode sizes.
// ---------------------------------------------------------------------------------------------
// A subgraph in its own compact index space: local ids 0..n-1, `glob` maps back to node ids.  Every recursion
// level relabels its two halves in BFS order into fresh subgraphs, so sweeps run over contiguous memory and
// concurrent branches share no cache lines.
struct Sub {
    int32_t n = 0;
    std::vector<int32_t> ptr;    // n + 1
    std::vector<int32_t> adj;    // local ids
    std::vector<int32_t> glob;   // local -> node id
};

struct Dissector {
    int leaf;
    int max_par_depth;
    // branches currently being dissected by some thread.  Below max_par_depth a split still forks when a core is free:
    // the halves are uneven (min_bal), so with a fixed fork depth the largest depth-4 subtree (17k of 10^5 nodes) ran
    // serially for ~11 ms after the other fifteen had finished.  The result does not depend on who forks (branches
    // are appended in a fixed order).
    std::atomic<int> running{1};
    int hw_threads = 1;
    // sweeps of large subgraphs go to the accelerator when the engine supplied one (same queue order, so the same
    // ordering); a graph it keeps declining (not connected, thin levels) stops being offered
    SweepAccel *accel = nullptr;
    int32_t accel_min = 0;
    std::atomic<int> accel_declined{0};
    bool offer(const Sub &G) const {
        return accel && G.n >= accel_min && accel_declined.load(std::memory_order_relaxed) < 2;
    }
    bool accel_sweep(const Sub &G, int32_t start, std::vector<int32_t> &lvl, std::vector<int32_t> &visit, int32_t &nlev) {
        lvl.resize(G.n);
        visit.resize(G.n);
        int32_t nl = 0;
        const int r = accel->sweep(G.n, G.ptr.data(), G.adj.data(), (int64_t)G.adj.size(), start, lvl.data(), visit.data(), &nl);
        if (r == 1) {
            nlev = nl;
            return true;
        }
        if (r < 0) accel_declined.fetch_add(1, std::memory_order_relaxed);
        visit.clear();
        return false;
    }
    // smaller side of a split should hold at least this fraction.  Balanced trees cost more flops but fewer dependent
    // panel steps, and the GPU is latency-bound at the top of the tree: on the 10^5-minima landscape 0.25 / 0.30 /
    // 0.35 / 0.40 give 1.08 / 1.11 / 1.19 / 1.29e10 flops, 5.2 / 5.15 / 4.76 / 4.91 ms per step and 48 / 45 / 38 / 37 ms
    // of ordering time
    double min_bal = 0.35;
    int block_depth = 4;     // subtrees at this depth become the parallel blocks of the symbolic factorisation
    bool trace = std::getenv("NGT_B200_TIMING") != nullptr && std::getenv("NGT_B200_ND_TRACE") != nullptr;

    // result of dissecting one subset: nodes in elimination order + supernode sizes
    struct Out {
        std::vector<int32_t> order;
        std::vector<int32_t> sn_size;
        // ranges [first, last) of supernode indices that form complete dissection subtrees: no supernode outside a
        // range is a descendant of one inside it, so the symbolic factorisation can process the ranges side by side
        std::vector<std::pair<int32_t, int32_t> > blocks;
        void emit_glob(const Sub &G, const std::vector<int32_t> &locals) {
            if (locals.empty()) return;
            sn_size.push_back((int32_t)locals.size());
            for (int32_t v : locals) order.push_back(G.glob[v]);
        }
        void emit_all(const Sub &G) {
            if (G.n == 0) return;
            sn_size.push_back(G.n);
            order.insert(order.end(), G.glob.begin(), G.glob.end());
        }
        void append(Out &o) {
            const int32_t shift = (int32_t)sn_size.size();
            order.insert(order.end(), o.order.begin(), o.order.end());
            sn_size.insert(sn_size.end(), o.sn_size.begin(), o.sn_size.end());
            for (auto &b : o.blocks) blocks.emplace_back(b.first + shift, b.second + shift);
        }
    };

    explicit Dissector(int leaf_) : leaf(leaf_) {
        unsigned hw = host_threads();
        hw_threads = (int)hw;
        max_par_depth = hw >= 16 ? 4 : hw >= 8 ? 3 : hw >= 4 ? 2 : hw >= 2 ? 1 : 0;
        if (const char *e = std::getenv("NGT_B200_ND_THREADS_DEPTH")) max_par_depth = std::atoi(e);
    }

    // BFS from `start` over unvisited nodes; appends to `visit` and writes lvl[] of what it discovers.  The visited set
    // is a bitmap (`seen`, one bit per node: 12 KB for 10^5 nodes, resident in L1) because ~19 of 20 adjacency entries
    // lead to a node that is already visited, and that test was a miss in the 4-byte level array otherwise.  On hosts
    // with AVX-512 sixteen entries of a row are tested at once (gather from the bitmap); both versions visit in the
    // same order.
    static void bfs_scalar(const Sub &G, int32_t start, std::vector<uint64_t> &seen, std::vector<int32_t> &lvl,
                           std::vector<int32_t> &visit, int32_t &nlev) {
        size_t head = visit.size();
        visit.push_back(start);
        lvl[start] = 0;
        seen[start >> 6] |= 1ull << (start & 63);
        nlev = 1;
        const int32_t *adj = G.adj.data();
        const int32_t *ptr = G.ptr.data();
        uint64_t *bits = seen.data();
        int32_t *lv = lvl.data();
        while (head < visit.size()) {
            const int32_t u = visit[head++];
            const int32_t lu = lv[u] + 1;
            // the queue is known a few nodes ahead: pull the adjacency row of a later node towards the core
            if (head + 4 < visit.size()) {
                const int32_t *r = adj + ptr[visit[head + 4]];
                __builtin_prefetch(r);
                __builtin_prefetch(r + 16);
            }
            for (int32_t e = ptr[u], e1 = ptr[u + 1]; e < e1; ++e) {
                const int32_t v = adj[e];
                const uint64_t bit = 1ull << (v & 63);
                uint64_t &w = bits[v >> 6];
                if (w & bit) continue;
                w |= bit;
                lv[v] = lu;
                nlev = lu + 1;
                visit.push_back(v);
            }
        }
    }
#if defined(__x86_64__) && defined(__GNUC__)
    __attribute__((target("avx512f"))) static void bfs_avx512(const Sub &G, int32_t start, std::vector<uint64_t> &seen,
                                                               std::vector<int32_t> &lvl, std::vector<int32_t> &visit,
                                                               int32_t &nlev) {
        // the queue lives in a buffer sized for every node (no push_back in the loop); little-endian: bit v of the
        // 64-bit words is bit v & 31 of 32-bit word v >> 5
        const size_t first = visit.size();
        visit.resize(first + (size_t)G.n);
        int32_t *q = visit.data();
        size_t head = first, tail = first;
        q[tail++] = start;
        lvl[start] = 0;
        uint32_t *bits = reinterpret_cast<uint32_t *>(seen.data());
        bits[start >> 5] |= 1u << (start & 31);
        nlev = 1;
        const int32_t *adj = G.adj.data();
        const int32_t *ptr = G.ptr.data();
        int32_t *lv = lvl.data();
        const __m512i one = _mm512_set1_epi32(1), m31 = _mm512_set1_epi32(31);
        while (head < tail) {
            const int32_t u = q[head++];
            const int32_t lu = lv[u] + 1;
            if (head + 4 < tail) {
                const int32_t *r = adj + ptr[q[head + 4]];
                __builtin_prefetch(r);
                __builtin_prefetch(r + 16);
            }
            for (int32_t e = ptr[u], e1 = ptr[u + 1]; e < e1; e += 16) {
                const int rem = e1 - e;
                const __mmask16 km = rem >= 16 ? (__mmask16)0xffff : (__mmask16)((1u << rem) - 1u);
                const __m512i v = _mm512_maskz_loadu_epi32(km, adj + e);
                const __m512i w = _mm512_mask_i32gather_epi32(_mm512_setzero_si512(), km, _mm512_srli_epi32(v, 5), bits, 4);
                const __m512i b = _mm512_sllv_epi32(one, _mm512_and_epi32(v, m31));
                unsigned un = _mm512_mask_testn_epi32_mask(km, w, b);
                while (un) {                                   // rows hold no duplicates: the lanes are independent
                    const int l = __builtin_ctz(un);
                    un &= un - 1;
                    const int32_t x = adj[e + l];
                    bits[x >> 5] |= 1u << (x & 31);
                    lv[x] = lu;
                    nlev = lu + 1;
                    q[tail++] = x;
                }
            }
        }
        visit.resize(tail);
    }
#endif
    static void bfs(const Sub &G, int32_t start, std::vector<uint64_t> &seen, std::vector<int32_t> &lvl,
                    std::vector<int32_t> &visit, int32_t &nlev) {
#if defined(__x86_64__) && defined(__GNUC__)
        if (wide_host()) return bfs_avx512(G, start, seen, lvl, visit, nlev);
#endif
        bfs_scalar(G, start, seen, lvl, visit, nlev);
    }

    // induced subgraph on `members` (listed in the order that becomes the new local numbering).  Large subgraphs are
    // induced by a team of threads (count, prefix, fill: every member's slice of the adjacency has one writer, so the
    // result is the serial one bit for bit): at the top of the dissection only one or two branches exist and the other
    // cores idle, while these passes over the whole adjacency were ~40 % of the critical path of the ordering.
    // the entries of one adjacency row that belong to the induced subgraph, relabelled (newid < 0: not a member);
    // `out` has room for the whole row.  AVX-512: sixteen look-ups per gather, survivors packed by a compress store.
    static int32_t *filter_row_scalar(const int32_t *row, int32_t len, const int32_t *newid, int32_t *out) {
        for (int32_t e = 0; e < len; ++e) {
            if (e + 8 < len) __builtin_prefetch(&newid[row[e + 8]]);
            const int32_t x = newid[row[e]];
            if (x >= 0) *out++ = x;
        }
        return out;
    }
#if defined(__x86_64__) && defined(__GNUC__)
    __attribute__((target("avx512f"))) static int32_t *filter_row_avx512(const int32_t *row, int32_t len, const int32_t *newid,
                                                                         int32_t *out) {
        for (int32_t e = 0; e < len; e += 16) {
            const int rem = len - e;
            const __mmask16 km = rem >= 16 ? (__mmask16)0xffff : (__mmask16)((1u << rem) - 1u);
            const __m512i v = _mm512_maskz_loadu_epi32(km, row + e);
            const __m512i x = _mm512_mask_i32gather_epi32(_mm512_set1_epi32(-1), km, v, newid, 4);
            const __mmask16 keep = _mm512_mask_cmpge_epi32_mask(km, x, _mm512_setzero_si512());
            _mm512_mask_compressstoreu_epi32(out, keep, x);
            out += __builtin_popcount((unsigned)keep);
        }
        return out;
    }
#endif
    static bool wide_host() {
#if defined(__x86_64__) && defined(__GNUC__)
        static const bool wide = __builtin_cpu_supports("avx512f") && std::getenv("NGT_B200_NO_AVX512") == nullptr;
        return wide;
#else
        return false;
#endif
    }
    static int32_t *filter_row(const int32_t *row, int32_t len, const int32_t *newid, int32_t *out) {
#if defined(__x86_64__) && defined(__GNUC__)
        if (wide_host()) return filter_row_avx512(row, len, newid, out);
#endif
        return filter_row_scalar(row, len, newid, out);
    }
    static void induce(const Sub &G, const std::vector<int32_t> &members, std::vector<int32_t> &newid, Sub &H, int nthreads = 1,
                       bool reset = true) {
        H.n = (int32_t)members.size();
        H.glob.resize(H.n);
        H.ptr.assign(H.n + 1, 0);
        const int32_t *adj = G.adj.data();
        if (nthreads > 1 && H.n >= 8192) {
            // every thread filters the rows of its slice of the members into a buffer of its own (one look-up per
            // adjacency entry; counting first and filling afterwards looked every entry up twice), then the buffers are
            // laid end to end
            const int T = nthreads;
            std::vector<std::vector<int32_t> > buf(T);
            std::vector<int64_t> tot(T + 1, 0);
            parallel_ranges(T, H.n, [&](int, int64_t i0, int64_t i1) {
                for (int64_t i = i0; i < i1; ++i) newid[members[i]] = (int32_t)i;
            });
            parallel_ranges(T, H.n, [&](int t, int64_t i0, int64_t i1) {
                size_t cap = 0;
                for (int64_t i = i0; i < i1; ++i) cap += (size_t)(G.ptr[members[i] + 1] - G.ptr[members[i]]);
                std::vector<int32_t> &b = buf[t];
                b.resize(cap);
                int32_t *w = b.data();
                for (int64_t i = i0; i < i1; ++i) {
                    const int32_t v = members[i];
                    H.glob[i] = G.glob[v];
                    if (i + 4 < i1) __builtin_prefetch(adj + G.ptr[members[i + 4]]);
                    const int32_t *w0 = w;
                    w = filter_row(adj + G.ptr[v], G.ptr[v + 1] - G.ptr[v], newid.data(), w);
                    H.ptr[i + 1] = (int32_t)(w - w0);        // row length for now
                }
                tot[t + 1] = w - b.data();
            });
            for (int t = 0; t < T; ++t) tot[t + 1] += tot[t];
            H.adj.resize((size_t)tot[T]);
            parallel_ranges(T, H.n, [&](int t, int64_t i0, int64_t i1) {
                int32_t run = (int32_t)tot[t];
                for (int64_t i = i0; i < i1; ++i) {
                    run += H.ptr[i + 1];
                    H.ptr[i + 1] = run;
                }
                std::copy(buf[t].begin(), buf[t].begin() + (tot[t + 1] - tot[t]), H.adj.begin() + tot[t]);
                if (reset)
                    for (int64_t i = i0; i < i1; ++i) newid[members[i]] = -1;
            });
            return;
        }
        for (int32_t i = 0; i < H.n; ++i) newid[members[i]] = i;
        size_t cap = 0;
        for (int32_t v : members) cap += (size_t)(G.ptr[v + 1] - G.ptr[v]);
        H.adj.resize(cap);
        int32_t *w = H.adj.data();
        for (int32_t i = 0; i < H.n; ++i) {
            const int32_t v = members[i];
            H.glob[i] = G.glob[v];
            if (i + 4 < H.n) __builtin_prefetch(adj + G.ptr[members[i + 4]]);
            w = filter_row(adj + G.ptr[v], G.ptr[v + 1] - G.ptr[v], newid.data(), w);
            H.ptr[i + 1] = (int32_t)(w - H.adj.data());
        }
        H.adj.resize((size_t)(w - H.adj.data()));
        for (int32_t v : members) newid[v] = -1;
    }
    // threads one branch at recursion depth `depth` may use inside its own passes: the branches of a depth share the machine
    static int team_threads(int depth) {
        const unsigned hw = host_threads();
        return (int)std::max(1u, std::min(8u, hw >> std::min(depth + 1, 5)));
    }

    // `known_lvl`: levels of G's nodes in a sweep from node 0 whose visiting order is 0, 1, 2, ... -- true for the
    // near half of a split, which is numbered in the parent's sweep order from the parent's root: every node there is
    // discovered by a node of the same half one level down, in the same relative order, so the child's first sweep
    // would reproduce exactly that numbering and those levels.  It is skipped then (same ordering, bit for bit).
    void dissect(Sub &G, Out &out, int depth, const std::vector<int32_t> *known_lvl = nullptr) {
        if (G.n == 0) return;
        if (G.n <= leaf) { out.emit_all(G); return; }
        const int32_t sn_enter = (int32_t)out.sn_size.size();
        const auto t_enter = std::chrono::steady_clock::now();
        const int32_t n_enter = G.n;
        std::vector<int32_t> lvl, visit;
        std::vector<uint64_t> seen;
        int32_t nlev = 0;

        // ---- first sweep: connectivity + a far node
        bool swept = false;    // lvl / visit already hold the level structure rooted at the pseudo-peripheral node
        if (known_lvl && (int32_t)known_lvl->size() == G.n) {
            lvl = *known_lvl;
            visit.resize(G.n);
            for (int32_t i = 0; i < G.n; ++i) visit[i] = i;
            nlev = lvl.back() + 1;
        } else if (offer(G) && accel_sweep(G, -1, lvl, visit, nlev)) {
            swept = true;      // both sweeps done there (a graph that is not connected is declined)
        } else {
            lvl.assign(G.n, -1);
            visit.clear();
            visit.reserve(G.n);
            seen.assign(((size_t)G.n + 63) / 64, 0);
            bfs(G, 0, seen, lvl, visit, nlev);
        }
        if ((int32_t)visit.size() != G.n) {
            // several components: small ones are packed into leaf-sized bins, large ones dissected on their own
            std::vector<int32_t> comp_start{0};
            comp_start.push_back((int32_t)visit.size());
            for (int32_t v = 0; v < G.n; ++v) {
                if (lvl[v] >= 0) continue;
                int32_t nl;
                bfs(G, v, seen, lvl, visit, nl);
                comp_start.push_back((int32_t)visit.size());
            }
            std::vector<int32_t> bin, newid(G.n, -1);
            std::vector<Sub> big;
            for (size_t c = 0; c + 1 < comp_start.size(); ++c) {
                std::vector<int32_t> mem(visit.begin() + comp_start[c], visit.begin() + comp_start[c + 1]);
                if ((int)mem.size() > leaf) {
                    big.emplace_back();
                    induce(G, mem, newid, big.back());
                } else {
                    if ((int)(bin.size() + mem.size()) > leaf) { out.emit_glob(G, bin); bin.clear(); }
                    bin.insert(bin.end(), mem.begin(), mem.end());
                }
            }
            out.emit_glob(G, bin);
            G = Sub();
            for (auto &H : big) dissect(H, out, depth);
            return;
        }
        // ---- second sweep from a min-degree node of the last level (pseudo-peripheral, George & Liu)
        if (!swept) {
            int32_t best = visit.back();
            int32_t bestdeg = G.ptr[best + 1] - G.ptr[best];
            for (size_t i = visit.size(); i-- > 0;) {
                const int32_t v = visit[i];
                if (lvl[v] != nlev - 1) break;
                const int32_t d = G.ptr[v + 1] - G.ptr[v];
                if (d < bestdeg) { bestdeg = d; best = v; }
            }
            if (best != 0 && !(offer(G) && accel_sweep(G, best, lvl, visit, nlev))) {
                seen.assign(((size_t)G.n + 63) / 64, 0);   // connected: the sweep rewrites every level
                visit.clear();
                visit.reserve(G.n);
                bfs(G, best, seen, lvl, visit, nlev);
            }
        }
        if (nlev < 3) { out.emit_all(G); return; }   // no separator exists (e.g. a clique): one dense supernode

        // ---- choose the separator level: smallest level, imbalance penalised
        std::vector<int64_t> cnt(nlev, 0);
        for (int32_t v = 0; v < G.n; ++v) cnt[lvl[v]]++;
        std::vector<int64_t> before(nlev + 1, 0);
        for (int i = 0; i < nlev; ++i) before[i + 1] = before[i] + cnt[i];
        const int64_t tot = G.n;
        int bestl = -1;
        double bestcost = 1e300;
        for (int l = 1; l <= nlev - 2; ++l) {
            const int64_t a = before[l], b = tot - before[l + 1];
            const int64_t mn = std::min(a, b);
            if (mn <= 0) continue;
            const double bal = (double)mn / (double)tot;
            const double cost = (double)cnt[l] * (bal >= min_bal ? 1.0 : (min_bal / bal) * (min_bal / bal));
            if (cost < bestcost) { bestcost = cost; bestl = l; }
        }
        if (bestl < 0 || cnt[bestl] * 10 > tot * 6) { out.emit_all(G); return; }

        // ---- split in BFS order; separator nodes that do not touch the far side join the near side
        std::vector<int32_t> P1, P2, sep, lvl1;   // lvl1: levels of the near half, in its new numbering
        for (int32_t v : visit) {
            if (lvl[v] < bestl) { P1.push_back(v); lvl1.push_back(lvl[v]); }
            else if (lvl[v] > bestl) P2.push_back(v);
            else {
                bool far = false;
                for (int32_t e = G.ptr[v]; e < G.ptr[v + 1] && !far; ++e)
                    if (lvl[G.adj[e]] == bestl + 1) far = true;
                if (far) sep.push_back(v); else { P1.push_back(v); lvl1.push_back(lvl[v]); }
            }
        }
        std::vector<int32_t> sep_glob(sep.size());
        for (size_t i = 0; i < sep.size(); ++i) sep_glob[i] = G.glob[sep[i]];
        std::vector<int32_t>().swap(lvl);
        std::vector<int32_t>().swap(visit);
        // Measured and not done: rooting the children's level structures at the ends of this one to save their first
        // sweep (ordering ~25 % faster, 8-39 % more elimination flops on the 10^5-minima landscape); a level-synchronous
        // team sweep with the serial visiting order (3 barriers per level: 55-84 ms instead of 48 ms, the frontiers of
        // a ~600-level structure are too short to split).
        if (trace && depth <= 5)
            std::fprintf(stderr, "[ngt timing]     dissect depth %d n %d sep %zu: sweeps+split %.2f ms\n", depth, n_enter, sep.size(),
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count());
        const bool fork = (depth < max_par_depth && P1.size() > 4096 && P2.size() > 4096) ||
                          (max_par_depth > 0 && P1.size() > 1024 && P2.size() > 1024 &&
                           running.load(std::memory_order_relaxed) < hw_threads);
        if (fork) {
            // both halves are induced and dissected concurrently; G stays alive (read-only) until both are induced
            Out o2;
            running.fetch_add(1, std::memory_order_relaxed);
            std::thread th([&] {
                Sub H2;
                {
                    std::vector<int32_t> newid(G.n, -1);
                    induce(G, P2, newid, H2, team_threads(depth), false);
                }
                dissect(H2, o2, depth + 1);
                running.fetch_sub(1, std::memory_order_relaxed);
            });
            {
                Sub H1;
                {
                    std::vector<int32_t> newid(G.n, -1);
                    induce(G, P1, newid, H1, team_threads(depth), false);
                }
                dissect(H1, out, depth + 1, &lvl1);
            }
            running.fetch_sub(1, std::memory_order_relaxed);   // waiting is not running
            th.join();
            running.fetch_add(1, std::memory_order_relaxed);
            G = Sub();
            out.append(o2);
        } else {
            Sub H1, H2;
            {
                std::vector<int32_t> newid(G.n, -1);
                induce(G, P1, newid, H1);
                induce(G, P2, newid, H2);
            }
            G = Sub();                       // release this level before descending
            std::vector<int32_t>().swap(P1);
            std::vector<int32_t>().swap(P2);
            dissect(H1, out, depth + 1, &lvl1);
            dissect(H2, out, depth + 1);
        }
        if (!sep_glob.empty()) {
            out.sn_size.push_back((int32_t)sep_glob.size());
            out.order.insert(out.order.end(), sep_glob.begin(), sep_glob.end());
        }
        if (depth == block_depth) out.blocks.emplace_back(sn_enter, (int32_t)out.sn_size.size());
    }
};

void finish_layout(Symbolic &S) {
    // pool layout, levels, batches, statistics
    const int32_t nf = (int32_t)S.fronts.size();
    int64_t off = 0;
    S.max_front = 0;
    for (auto &F : S.fronts) {
        F.f = F.k + F.m;
        F.ld = ((F.f + 1 + 7) / 8) * 8;
        F.off = off;
        off += (int64_t)F.f * F.ld;
        off = ((off + 15) / 16) * 16;
        S.max_front = std::max(S.max_front, F.f);
    }
    S.pool_doubles = off;

    // children CSR
    S.child_ptr.assign(nf + 1, 0);
    for (int32_t i = 0; i < nf; ++i)
        if (S.fronts[i].parent >= 0) S.child_ptr[S.fronts[i].parent + 1]++;
    for (int32_t i = 0; i < nf; ++i) S.child_ptr[i + 1] += S.child_ptr[i];
    S.child_ids.assign(S.child_ptr[nf], 0);
    {
        std::vector<int64_t> fill(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int32_t i = 0; i < nf; ++i)
            if (S.fronts[i].parent >= 0) S.child_ids[fill[S.fronts[i].parent]++] = i;
    }
    // levels (fronts are in elimination order, children before parents; root last)
    for (auto &F : S.fronts) F.level = 0;
    for (int32_t i = 0; i < nf; ++i) {
        int32_t p = S.fronts[i].parent;
        if (p >= 0) S.fronts[p].level = std::max(S.fronts[p].level, S.fronts[i].level + 1);
    }
    S.n_levels = 0;
    for (auto &F : S.fronts) S.n_levels = std::max(S.n_levels, F.level + 1);

    // relative index maps child -> parent
    S.rel_off.assign(nf + 1, 0);
    for (int32_t i = 0; i < nf; ++i) S.rel_off[i + 1] = S.rel_off[i] + (S.fronts[i].parent >= 0 ? S.fronts[i].m : 0);
    S.rel.assign(S.rel_off[nf], -1);
    {
        unsigned hw = host_threads();
        int T = (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
        if (S.rel.size() < 200000) T = 1;
        std::vector<char> bad(T, 0);
        // ranges are balanced by map entries, not by front count (the top fronts hold most of them)
        parallel_ranges(T, (int64_t)S.rel.size(), [&](int t, int64_t e0, int64_t e1) {
            int32_t i = (int32_t)(std::upper_bound(S.rel_off.begin(), S.rel_off.end(), e0) - S.rel_off.begin()) - 1;
            for (int64_t e = e0; e < e1; ++e) {
                while (S.rel_off[i + 1] <= e) ++i;
                const Front &F = S.fronts[i];
                const int32_t jj = (int32_t)(e - S.rel_off[i]);
                const int32_t li = local_index(S, F.parent, S.fidx[F.idx_off + F.k + jj]);
                if (li < 0) bad[t] = 1;
                S.rel[e] = li;
            }
        });
        for (char b : bad)
            if (b) throw std::logic_error("symbolic: child update index missing from parent front");
    }

    // batches: per level, fronts with pivots, grouped by size class and capped for gridDim.y.  Fronts of one batch
    // share every launch, so the panel chains of same-level fronts run side by side; a wide class ratio keeps the
    // number of sequential chains per level small (grids are sized by the largest front, smaller ones exit early).
    const int size_class_ratio = 8;
    S.batches.clear();
    std::vector<std::vector<int32_t> > by_level(S.n_levels);
    for (int32_t i = 0; i < nf; ++i)
        if (S.fronts[i].k > 0) by_level[S.fronts[i].level].push_back(i);
    for (int32_t l = 0; l < S.n_levels; ++l) {
        auto &v = by_level[l];
        std::sort(v.begin(), v.end(), [&](int32_t a, int32_t b) {
            if (S.fronts[a].f != S.fronts[b].f) return S.fronts[a].f > S.fronts[b].f;
            return a < b;
        });
        size_t i = 0;
        while (i < v.size()) {
            Batch b;
            b.level = l;
            b.max_f = S.fronts[v[i]].f;
            b.max_k = 0;
            size_t j = i;
            while (j < v.size() && j - i < 32768 && S.fronts[v[j]].f * size_class_ratio >= b.max_f) {
                b.max_k = std::max(b.max_k, S.fronts[v[j]].k);
                b.fronts.push_back(v[j]);
                ++j;
            }
            S.batches.push_back(b);
            i = j;
        }
    }

    // algorithmic work: degrees at elimination time under this ordering
    S.alg_flops = S.alg_bytes = 0;
    for (auto &F : S.fronts) {
        for (int32_t i = 0; i < F.k; ++i) {
            double d = (double)(F.k - 1 - i) + F.m;
            S.alg_flops += 2.0 * d * d + 4.0 * d;
            S.alg_bytes += 16.0 * d * d + 24.0 * d + 16.0 * d + 8.0;
        }
    }
}

}  // namespace

int32_t local_index(const Symbolic &S, int32_t fr, int32_t p) {
    const Front &F = S.fronts[fr];
    if (F.k > 0 && p >= F.first && p < F.first + F.k) return p - F.first;
    const int32_t *b = S.fidx.data() + F.idx_off + F.k, *e = b + F.m;
    const int32_t *it = std::lower_bound(b, e, p);
    if (it == e || *it != p) return -1;
    return F.k + (int32_t)(it - b);
}

static void assign_kept_positions(const std::vector<int64_t> &A, const std::vector<int64_t> &B, Symbolic &S) {
    int32_t p = (int32_t)S.nI;
    for (int64_t a : A) { S.pos[a] = p; S.node_at[p] = (int32_t)a; ++p; }
    for (int64_t b : B) { S.pos[b] = p; S.node_at[p] = (int32_t)b; ++p; }
}

void analyse_dense(int64_t n, const std::vector<int64_t> &A, const std::vector<int64_t> &B, Symbolic &S) {
    S = Symbolic();
    S.n = n;
    S.nA = (int64_t)A.size();
    S.nB = (int64_t)B.size();
    S.nI = n - S.nA - S.nB;
    S.pos.assign(n, -1);
    S.node_at.assign(n, -1);
    std::vector<char> kept(n, 0);
    for (int64_t a : A) kept[a] = 1;
    for (int64_t b : B) kept[b] = 1;
    int32_t p = 0;
    for (int64_t u = 0; u < n; ++u)
        if (!kept[u]) { S.pos[u] = p; S.node_at[p] = (int32_t)u; ++p; }
    assign_kept_positions(A, B, S);
    const int32_t m = (int32_t)(S.nA + S.nB);
    S.pos_front.assign(n, 0);
    if (S.nI > 0) {
        Front F{};
        F.k = (int32_t)S.nI; F.m = m; F.first = 0; F.parent = 1; F.idx_off = 0;
        S.fronts.push_back(F);
        for (int32_t i = 0; i < n; ++i) S.fidx.push_back(i);
    }
    Front R{};
    R.k = 0; R.m = m; R.first = (int32_t)S.nI; R.parent = -1; R.idx_off = (int64_t)S.fidx.size();
    S.fronts.push_back(R);
    for (int32_t i = 0; i < m; ++i) S.fidx.push_back((int32_t)S.nI + i);
    S.root_id = (int32_t)S.fronts.size() - 1;
    for (int64_t q = 0; q < n; ++q) S.pos_front[q] = q < S.nI ? 0 : S.root_id;
    finish_layout(S);
}

void analyse_sparse(const Adjacency &g, const std::vector<char> &present,
                    const std::vector<int64_t> &A, const std::vector<int64_t> &B,
                    int leaf_size, Symbolic &S, SweepAccel *accel, int32_t accel_min_nodes) {
    S = Symbolic();
    const int64_t n = g.n;
    S.n = n;
    S.nA = (int64_t)A.size();
    S.nB = (int64_t)B.size();
    std::vector<char> kept(n, 0);
    for (int64_t a : A) kept[a] = 1;
    for (int64_t b : B) kept[b] = 1;

    // ---- ordering of the intermediates
    std::vector<int32_t> inter;
    for (int64_t u = 0; u < n; ++u)
        if (present[u] && !kept[u]) inter.push_back((int32_t)u);
    S.nI = (int64_t)inter.size();
    Timer tm;
    Dissector D(leaf_size > 0 ? leaf_size : 64);
    D.accel = accel;
    D.accel_min = accel_min_nodes > 0 ? accel_min_nodes : (accel ? accel->min_nodes : 0);
    // top-level subgraph: the rate graph induced on the intermediates (A ∪ B and absent nodes are left out, so
    // the dissector never walks through them)
    Sub top;
    {
        std::vector<int32_t> newid(n, -1);
        top.n = (int32_t)inter.size();
        top.glob = inter;
        for (int32_t i = 0; i < top.n; ++i) newid[inter[i]] = i;
        top.ptr.assign(top.n + 1, 0);
        // every thread filters the rows of its slice of the intermediates into a buffer of its own (one look-up per
        // adjacency entry), then the buffers are laid end to end
        const int T = top.n < 20000 ? 1 : (int)std::min(host_threads(), 8u);
        std::vector<std::vector<int32_t> > buf(T);
        std::vector<int64_t> tot(T + 1, 0);
        parallel_ranges(T, top.n, [&](int t, int64_t i0, int64_t i1) {
            size_t cap = 0;
            for (int64_t i = i0; i < i1; ++i) cap += (size_t)(g.ptr[inter[i] + 1] - g.ptr[inter[i]]);
            std::vector<int32_t> &b = buf[t];
            b.resize(cap);
            int32_t *w = b.data();
            for (int64_t i = i0; i < i1; ++i) {
                const int32_t u = inter[i];
                const int32_t *w0 = w;
                w = Dissector::filter_row(g.adj.data() + g.ptr[u], (int32_t)(g.ptr[u + 1] - g.ptr[u]), newid.data(), w);
                top.ptr[i + 1] = (int32_t)(w - w0);          // row length for now
            }
            tot[t + 1] = w - b.data();
        });
        for (int t = 0; t < T; ++t) tot[t + 1] += tot[t];
        if (tot[T] > 0x7fffffffLL) throw std::length_error("symbolic: more than 2^31 adjacency entries");
        top.adj.resize((size_t)tot[T]);
        parallel_ranges(T, top.n, [&](int t, int64_t i0, int64_t i1) {
            int32_t run = (int32_t)tot[t];
            for (int64_t i = i0; i < i1; ++i) {
                run += top.ptr[i + 1];
                top.ptr[i + 1] = run;
            }
            std::copy(buf[t].begin(), buf[t].begin() + (tot[t + 1] - tot[t]), top.adj.begin() + tot[t]);
        });
    }
    tm.lap("top-level subgraph");
    Dissector::Out ord;
    ord.order.reserve(inter.size());
    D.dissect(top, ord, 0);
    if ((int64_t)ord.order.size() != S.nI) throw std::logic_error("symbolic: ordering lost nodes");
    std::vector<int32_t> sn_start;
    {
        int32_t acc = 0;
        for (int32_t z : ord.sn_size) { sn_start.push_back(acc); acc += z; }
    }
    tm.lap("nested dissection");

    const int64_t npos = S.nI + S.nA + S.nB;
    S.pos.assign(n, -1);
    S.node_at.assign(npos, -1);
    for (int64_t p = 0; p < S.nI; ++p) { S.pos[ord.order[p]] = (int32_t)p; S.node_at[p] = ord.order[p]; }
    assign_kept_positions(A, B, S);

    // ---- supernodal symbolic factorisation
    const int32_t ns = (int32_t)sn_start.size();
    sn_start.push_back((int32_t)S.nI);
    S.pos_front.assign(npos, -1);
    S.fronts.resize(ns + 1);
    std::vector<int32_t> mark(npos, -1);
    std::vector<std::vector<int32_t> > kids(ns + 1);
    std::vector<std::vector<int32_t> > upd(ns);   // update sets (freed into fidx at the end)
    for (int32_t s = 0; s < ns; ++s)
        for (int32_t p = sn_start[s]; p < sn_start[s + 1]; ++p) S.pos_front[p] = s;
    // phase A (parallel over supernodes): the later neighbours of each supernode's own pivots, sorted and unique
    {
        unsigned hw = host_threads();
        int T = (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
        if (ns < 256) T = 1;
        parallel_ranges(T, ns, [&](int /*t*/, int64_t s0, int64_t s1) {
            std::vector<int32_t> mk(npos, -1);
            for (int64_t s = s0; s < s1; ++s) {
                const int32_t first = sn_start[s], last = sn_start[s + 1];
                std::vector<int32_t> &U = upd[s];
                for (int32_t p = first; p < last; ++p) {
                    const int32_t u = S.node_at[p];
                    for (int64_t e = g.ptr[u]; e < g.ptr[u + 1]; ++e) {
                        const int32_t q = S.pos[g.adj[e]];
                        if (q >= last && mk[q] != (int32_t)s) { mk[q] = (int32_t)s; U.push_back(q); }
                    }
                }
            }
        });
    }
    tm.lap("  symbolic: own neighbours (parallel)");
    // phase B (children before parents): add what the children pass up.  Complete dissection subtrees (ord.blocks)
    // are independent of each other and run side by side; what is left (the separators above them and supernodes
    // that never reached the block depth) follows in elimination order.  A child whose parent lies outside its
    // block is attached to it afterwards, in increasing child order, so the children lists equal the serial ones.
    auto process = [&](int32_t s, std::vector<int32_t> &mk, int32_t blk_first, int32_t blk_last,
                       std::vector<std::pair<int32_t, int32_t> > *deferred) {
        const int32_t first = sn_start[s], last = sn_start[s + 1];
        std::vector<int32_t> &U = upd[s];
        for (int32_t q : U) mk[q] = s;
        for (int32_t c : kids[s])
            for (int32_t q : upd[c])
                if (q >= last && mk[q] != s) { mk[q] = s; U.push_back(q); }
        std::sort(U.begin(), U.end());
        Front &F = S.fronts[s];
        F.k = last - first;
        F.m = (int32_t)U.size();
        F.first = first;
        if (U.empty()) F.parent = -1;
        else if (U[0] >= S.nI) F.parent = ns;
        else F.parent = S.pos_front[U[0]];
        if (F.parent >= 0) {
            if (deferred && (F.parent < blk_first || F.parent >= blk_last)) deferred->emplace_back(F.parent, s);
            else kids[F.parent].push_back(s);
        }
    };
    {
        std::vector<char> in_block(ns, 0);
        const auto &blocks = ord.blocks;
        const int nb = (int)blocks.size();
        unsigned hw = host_threads();
        const int T = (ns < 256 || nb < 2) ? 1 : (int)std::max(1u, std::min(hw ? hw : 1u, 16u));
        if (T > 1) {
            for (auto &b : blocks)
                for (int32_t s = b.first; s < b.second; ++s) in_block[s] = 1;
            std::vector<std::vector<std::pair<int32_t, int32_t> > > deferred(nb);
            std::atomic<int> next{0};
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t)
                th.emplace_back([&] {
                    std::vector<int32_t> mk(npos, -1);
                    for (int b = next.fetch_add(1); b < nb; b = next.fetch_add(1))
                        for (int32_t s = blocks[b].first; s < blocks[b].second; ++s)
                            process(s, mk, blocks[b].first, blocks[b].second, &deferred[b]);
                });
            for (auto &x : th) x.join();
            std::vector<std::pair<int32_t, int32_t> > all;   // (parent, child)
            for (auto &d : deferred) all.insert(all.end(), d.begin(), d.end());
            std::sort(all.begin(), all.end(), [](const std::pair<int32_t, int32_t> &a, const std::pair<int32_t, int32_t> &b) {
                return a.second < b.second;
            });
            for (auto &pc : all) kids[pc.first].push_back(pc.second);
        }
        for (int32_t s = 0; s < ns; ++s)
            if (!in_block[s]) process(s, mark, 0, 0, nullptr);
    }
    tm.lap("  symbolic: merge children (subtrees in parallel)");
    // root
    {
        Front &R = S.fronts[ns];
        R.k = 0;
        R.m = (int32_t)(S.nA + S.nB);
        R.first = (int32_t)S.nI;
        R.parent = -1;
        for (int64_t p = S.nI; p < npos; ++p) S.pos_front[p] = ns;
    }
    S.root_id = ns;
    // index lists
    int64_t tot = 0;
    for (int32_t s = 0; s < ns; ++s) tot += S.fronts[s].k + S.fronts[s].m;
    tot += S.fronts[ns].m;
    S.fidx.reserve(tot);
    for (int32_t s = 0; s < ns; ++s) {
        Front &F = S.fronts[s];
        F.idx_off = (int64_t)S.fidx.size();
        for (int32_t p = F.first; p < F.first + F.k; ++p) S.fidx.push_back(p);
        S.fidx.insert(S.fidx.end(), upd[s].begin(), upd[s].end());
        std::vector<int32_t>().swap(upd[s]);
    }
    S.fronts[ns].idx_off = (int64_t)S.fidx.size();
    for (int64_t p = S.nI; p < npos; ++p) S.fidx.push_back((int32_t)p);
    tm.lap("  symbolic: index lists");
    finish_layout(S);
    tm.lap("layout, maps, batches");
}

}  // namespace ngt

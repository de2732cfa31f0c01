// Host-only timing harness for the symbolic analysis (no CUDA needed):
//   g++ -O3 -std=c++17 -pthread -I.. analysis_bench.cpp ../symbolic.cpp -o /tmp/analysis_bench
//   NGT_B200_TIMING=1 /tmp/analysis_bench graph.bin [reps [leaf_size]]
// graph.bin: int64 n, int64 nnz, int64 nA, int64 nB, src[nnz], dst[nnz], A[nA], B[nB]   (see scripts/dump_graph.py)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "symbolic.h"

// Stand-in for the device sweeps (csrc/ordering.cu), selected by NGT_ANALYSIS_MOCK_ACCEL: a serial breadth-first sweep
// behind the SweepAccel interface, so that the dissector's use of an accelerator (both sweeps in one call, the single
// sweep of a near half, declined calls) is covered without a GPU.  With the value 2 every third call is declined,
// alternately as "not now" and "not this graph".
struct MockSweeps : ngt::SweepAccel {
    int mode = 1;
    std::atomic<int> calls{0};
    static int bfs(int32_t n, const int32_t *ptr, const int32_t *adj, int32_t start, int32_t *lvl, int32_t *visit) {
        std::fill(lvl, lvl + n, -1);
        int32_t head = 0, tail = 0;
        visit[tail++] = start;
        lvl[start] = 0;
        while (head < tail) {
            const int32_t u = visit[head++];
            for (int32_t e = ptr[u]; e < ptr[u + 1]; ++e)
                if (lvl[adj[e]] < 0) {
                    lvl[adj[e]] = lvl[u] + 1;
                    visit[tail++] = adj[e];
                }
        }
        return tail;
    }
    int sweep(int32_t n, const int32_t *ptr, const int32_t *adj, int64_t, int32_t start, int32_t *lvl, int32_t *visit,
              int32_t *nlev) override {
        const int c = calls.fetch_add(1);
        if (mode == 2 && c % 3 == 2) return (c / 3) % 2 ? -1 : 0;
        if (bfs(n, ptr, adj, start < 0 ? 0 : start, lvl, visit) != n) return -1;
        if (start < 0) {
            const int32_t last = lvl[visit[n - 1]];
            int32_t best = visit[n - 1], bestdeg = ptr[best + 1] - ptr[best];
            for (int32_t i = n - 1; i >= 0 && lvl[visit[i]] == last; --i) {
                const int32_t d = ptr[visit[i] + 1] - ptr[visit[i]];
                if (d < bestdeg) { bestdeg = d; best = visit[i]; }
            }
            if (best != 0) bfs(n, ptr, adj, best, lvl, visit);
        }
        *nlev = lvl[visit[n - 1]] + 1;
        return 1;
    }
};

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s graph.bin [reps]\n", argv[0]); return 2; }
    FILE *fp = std::fopen(argv[1], "rb");
    if (!fp) { std::perror("open"); return 1; }
    int64_t hdr[4];
    if (std::fread(hdr, 8, 4, fp) != 4) return 1;
    const int64_t n = hdr[0], nnz = hdr[1], nA = hdr[2], nB = hdr[3];
    std::vector<int64_t> src(nnz), dst(nnz), A(nA), B(nB);
    if (std::fread(src.data(), 8, nnz, fp) != (size_t)nnz || std::fread(dst.data(), 8, nnz, fp) != (size_t)nnz ||
        std::fread(A.data(), 8, nA, fp) != (size_t)nA || std::fread(B.data(), 8, nB, fp) != (size_t)nB)
        return 1;
    std::fclose(fp);
    const int reps = argc > 2 ? std::atoi(argv[2]) : 3;
    const int leaf = argc > 3 ? std::atoi(argv[3]) : 0;
    for (int r = 0; r < reps; ++r) {
        auto t0 = std::chrono::steady_clock::now();
        ngt::EdgeIndex ix;
        ngt::Adjacency g;
        ngt::index_edges(n, nnz, src.data(), dst.data(), ix, g);
        auto t1 = std::chrono::steady_clock::now();
        ngt::Symbolic S;
        MockSweeps mock;
        const char *me = std::getenv("NGT_ANALYSIS_MOCK_ACCEL");
        if (me) mock.mode = std::atoi(me);
        ngt::analyse_sparse(g, ix.present, A, B, leaf, S, me ? &mock : nullptr, me ? 300 : 0);
        if (me) std::printf("rep %d: %d sweeps offered to the stand-in accelerator\n", r, mock.calls.load());
        auto t2 = std::chrono::steady_clock::now();
        unsigned long long hsh = 1469598103934665603ULL;   // FNV-1a over the elimination order
        for (int32_t v : S.node_at) { hsh ^= (unsigned long long)(unsigned)v; hsh *= 1099511628211ULL; }
        auto mix = [&](long long v) { hsh ^= (unsigned long long)v; hsh *= 1099511628211ULL; };
        for (const auto &F : S.fronts) { mix(F.k); mix(F.m); mix(F.parent); mix(F.level); mix(F.off); }
        for (int32_t v : S.fidx) mix(v);
        for (int32_t v : S.rel) mix(v);
        for (int32_t v : S.child_ids) mix(v);
        for (int64_t v : ix.outptr) mix(v);          // the edge index and the adjacency are threaded too
        for (int32_t v : ix.ent) mix(v);
        for (char v : ix.present) mix(v);
        for (int64_t v : g.ptr) mix(v);
        for (int32_t v : g.adj) mix(v);
        std::printf("rep %d: order + structure hash %016llx\n", r, hsh);
        std::printf("rep %d: index %.2f ms  analyse %.2f ms  fronts %zu levels %d max_front %d pool %.3f GB flops %.3e launches~%zu batches\n",
                    r, std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(t2 - t1).count(), S.fronts.size(), S.n_levels, S.max_front,
                    S.pool_doubles * 8e-9, S.alg_flops, S.batches.size());
    }
    return 0;
}

// Host-only timing harness for the symbolic analysis (no CUDA needed):
//   g++ -O3 -std=c++17 -pthread -I.. analysis_bench.cpp ../symbolic.cpp -o /tmp/analysis_bench
//   NGT_B200_TIMING=1 /tmp/analysis_bench graph.bin [reps [leaf_size]]
// graph.bin: int64 n, int64 nnz, int64 nA, int64 nB, src[nnz], dst[nnz], A[nA], B[nB]   (see scripts/dump_graph.py)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "symbolic.h"

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
        ngt::analyse_sparse(g, ix.present, A, B, leaf, S);
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

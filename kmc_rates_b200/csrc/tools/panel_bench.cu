// Micro-benchmark of the panel / row-panel / thin Schur-update chain on one synthetic front (needs a GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DNGT_PANEL_CLOCKS -I.. panel_bench.cu -o panel_bench
//   ./panel_bench [f] [k]
// Prints CUDA-event times per launch and, with -DNGT_PANEL_CLOCKS, the clock64 phase split inside the panel kernel.
#include "../kernels.cu"

#include <vector>

#define CKB(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char **argv) {
    const int f = argc > 1 ? std::atoi(argv[1]) : 1000;
    const int k = argc > 2 ? std::atoi(argv[2]) : 640;
    const int ld = ((f + 1 + 7) / 8) * 8;
    std::vector<double> h((size_t)f * ld, 0.0);
    unsigned long long st = 12345;
    auto rnd = [&] { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (double)(st >> 11) / 9007199254740992.0; };
    for (int r = 0; r < f; ++r) {
        double sum = 0;
        for (int c = 0; c < f; ++c) if (c != r) { h[(size_t)r * ld + c] = rnd(); sum += h[(size_t)r * ld + c]; }
        for (int c = 0; c < f; ++c) h[(size_t)r * ld + c] /= sum;
        h[(size_t)r * ld + f] = 1.0;
    }
    double *pool, *wbuf, *rsbuf;
    int *ids, *err;
    ngt::DFront F{};
    F.off = 0; F.f = f; F.k = k; F.ld = ld; F.parent = -1;
    ngt::DFront *dF;
    CKB(cudaMalloc(&pool, h.size() * 8));
    CKB(cudaMalloc(&wbuf, 64 * 1024 * 8));
    CKB(cudaMalloc(&rsbuf, 1024 * 1024));
    CKB(cudaMalloc(&ids, 4)); CKB(cudaMalloc(&err, 4)); CKB(cudaMalloc(&dF, sizeof(F)));
    CKB(cudaMemset(ids, 0, 4)); CKB(cudaMemset(err, 0, 4));
    CKB(cudaMemcpy(dF, &F, sizeof(F), cudaMemcpyHostToDevice));
    ngt::init_kernel_attributes();
    ngt::LaunchCtx c{};
    c.pool = pool; c.pool_stride = (long long)h.size(); c.nbatch = 1; c.fronts = dF; c.wbuf = wbuf; c.maxfb = 1;
    c.rsbuf = rsbuf; c.maxchunks = 64; c.errflag = err; c.stream = 0; c.gemm_path = 0;
    c.ko = ngt::KOpt{0, 0, 0, ngt::FUSE_COLS};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    auto timeit = [&](const char *name, auto &&fn) {
        CKB(cudaMemcpy(pool, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        for (int i = 0; i < 3; ++i) fn();
        CKB(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) fn();
        cudaEventRecord(e1);
        CKB(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::printf("%-42s %8.2f us per launch\n", name, 1e3 * ms / reps);
        return 0;
    };
    std::printf("front f=%d k=%d ld=%d\n", f, k, ld);
    timeit("panel (j=0)", [&] { ngt::launch_panel(c, ids, 1, 0); });
#ifdef NGT_PANEL_CLOCKS
    {
        long long clk[8];
        CKB(cudaMemcpyFromSymbol(clk, ngt::g_panel_clk, sizeof(clk)));
        std::printf("  clocks: rowsum %lld  gauss-jordan %lld  sync %lld  W store %lld  row panel %lld\n", clk[1] - clk[0],
                    clk[2] - clk[1], clk[3] - clk[2], clk[4] - clk[3], clk[5] > clk[4] ? clk[5] - clk[4] : 0);
        CKB(cudaMemcpyFromSymbol(clk, ngt::g_gj_clk, sizeof(clk)));
        std::printf("  blocked gauss-jordan, warp 0, sum over the 4 block steps: 8x8 inversion %lld  barrier %lld  DMMA updates %lld  "
                    "closing barrier %lld\n", clk[0], clk[1], clk[2], clk[3]);
    }
#endif
    timeit("rowpanel (j=0)", [&] { ngt::launch_rowpanel(c, ids, 1, 0, f + 1 - 32); });
    {
        ngt::Region R;
        ngt::get_region(f, 0, k, 0, ngt::MODE_A, R);
        const int ta = ((R.r1 - R.r0 + 127) / 128) * ((R.c1 - R.c0 + 63) / 64);
        ngt::get_region(f, 0, k, 0, ngt::MODE_B, R);
        const int tb = ((R.r1 - R.r0 + 127) / 128) * ((R.c1 - R.c0 + 63) / 64);
        std::printf("AB tiles %d + %d\n", ta, tb);
        timeit("update AB (j=0)", [&] { ngt::launch_update(c, ids, 1, 0, ngt::MODE_AB, ta + tb, ta); });
#ifdef NGT_PANEL_CLOCKS
        {
            long long clk[8];
            CKB(cudaMemcpyFromSymbol(clk, ngt::g_upd_clk, sizeof(clk)));
            std::printf("  update clocks (CTA 0): setup+C loads issued %lld  prologue issue %lld  first stage landed %lld  main loop %lld  epilogue %lld\n",
                        clk[1] - clk[0], clk[2] - clk[1], clk[3] - clk[2], clk[4] - clk[3], clk[5] - clk[4]);
        }
#endif
        {
            ngt::get_region(f, 0, k, 0, ngt::MODE_A, R);
            const int ta2 = ((R.r1 - R.r0 + 63) / 64) * ((R.c1 - R.c0 + 63) / 64);
            ngt::get_region(f, 0, k, 0, ngt::MODE_B, R);
            const int tb2 = ((R.r1 - R.r0 + 63) / 64) * ((R.c1 - R.c0 + 63) / 64);
            ngt::LaunchCtx c64 = c;
            c64.tile_m = 64;
            std::printf("AB half-height tiles %d + %d\n", ta2, tb2);
            timeit("update AB (j=0), 64-row tiles", [&] { ngt::launch_update(c64, ids, 1, 0, ngt::MODE_AB, ta2 + tb2, ta2); });
        }
        timeit("empty-ish kernel (rowsum early exit)", [&] { ngt::launch_rowsum(c, ids, 1, 0, 512); });
        {
            ngt::LaunchCtx cc = c;
            cc.panel_cluster = 1;
            timeit("panel, cluster variant (row sums + GJ + row panel)", [&] { ngt::launch_panel(cc, ids, 1, 0); });
            ngt::get_region(f, 0, k, 0, ngt::MODE_A, R);
            const int ta2 = ((R.r1 - R.r0 + 63) / 64) * ((R.c1 - R.c0 + 63) / 64);
            ngt::get_region(f, 0, k, 0, ngt::MODE_B, R);
            const int tb2 = ((R.r1 - R.r0 + 63) / 64) * ((R.c1 - R.c0 + 63) / 64);
            ngt::LaunchCtx c64 = c;
            c64.tile_m = 64;
            timeit("cluster panel + AB(64-row) chain", [&] {
                ngt::launch_panel(cc, ids, 1, 0);
                ngt::launch_update(c64, ids, 1, 0, ngt::MODE_AB, ta2 + tb2, ta2);
            });
        }
        timeit("panel+rowpanel+AB chain", [&] {
            ngt::launch_panel(c, ids, 1, 0);
            ngt::launch_rowpanel(c, ids, 1, 0, f + 1 - 32);
            ngt::launch_update(c, ids, 1, 0, ngt::MODE_AB, ta + tb, ta);
        });
    }
    return 0;
}

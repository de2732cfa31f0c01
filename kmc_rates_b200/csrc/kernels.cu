// CUDA kernels of the multifrontal NGT engine, written for sm_100a (B200).
// Layout and the blocked elimination scheme are described in kernels.cuh.
#include "kernels.cuh"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <string>

namespace ngt {

// =====================================================================================================
// helpers
// =====================================================================================================
// Programmatic dependent launch: every kernel on the elimination chain waits for its predecessor (first thing, in
// every CTA, before any early exit) and then lets its successor start launching, so the launch latency of the next
// kernel overlaps this kernel's execution instead of adding to the chain.  A successor's CTAs do nothing before
// their own wait returns (= this grid complete and flushed), so ordering and memory visibility are those of ordinary
// stream order.  (Triggering before the wait pipelines launches deeper but measured slower on the config 2 step:
// 5.73 vs 5.58 ms.)  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// thread-block cluster helpers (raw PTX; every thread of the CTA must call cluster_sync_all convergently)
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// a launch that the runtime refuses (bad configuration, missing attribute) must not pass silently
static void check_launch(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw LaunchFailure(std::string(what) + ": " + cudaGetErrorString(e));
}

static bool use_pdl() {
    static const bool on = std::getenv("NGT_B200_NO_PDL") == nullptr;
    return on;
}

template <typename... KArgs, typename... Args>
static void launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    if (use_pdl()) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    check_launch(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...), "cudaLaunchKernelEx");
}

// same, as thread-block clusters of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
static void launch_chain_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                 Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cluster_x;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
    if (use_pdl()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    check_launch(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...), "cudaLaunchKernelEx (cluster)");
}

// single-front launches carry their front descriptor in the kernel parameters (saves two dependent global loads)
__device__ __forceinline__ DFront fetch_front(const DFront *fronts, const int *ids, int slot, const DFront &F0, int has_f0) {
    return has_f0 ? F0 : fronts[ids[slot]];
}

// FP64 tensor-core MMA, C(8x8) += A(8x4) B(4x8): lane (g = lane/4, t = lane%4) holds A[g][t], B[t][g], C[g][2t..2t+1].
// On B200 the vector FP64 pipe issues one warp DFMA per 8 cycles per scheduler (16 FMA/clk/SM, measured with
// tools/lat_bench.cu); DMMA runs at ~62 FMA/clk/SM, so every product with a GEMM shape goes through it.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// U' = W X for one slab of a row panel, on the tensor cores.  Wd: W, 32 x 32 with row stride WD_LD; xs: the slab,
// 32 x CW with row stride CW + 4 (both strides = 4 mod 16 doubles: conflict-free fragment loads).  Warps take column
// groups of 8 round-robin; the result goes straight to global memory (rows < nbp, columns < ncols_live).
constexpr int WD_LD = NB_INNER + 4;
template <int CW, int NWARPS>
__device__ __forceinline__ void rowpanel_slab_mma(const double *Wd, const double *xs, double *Mout, long long ld,
                                                  int nbp, int ncols_live, int warp, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    for (int nt = warp; nt < CW / 8; nt += NWARPS) {
        if (nt * 8 >= ncols_live) break;
        double acc[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) acc[mi][0] = acc[mi][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
            const double b = xs[(k4 * 4 + t4) * (CW + 4) + nt * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][0], acc[mi][1], Wd[(mi * 8 + g) * WD_LD + k4 * 4 + t4], b);
        }
        const int c = nt * 8 + 2 * t4;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int r = mi * 8 + g;
            if (r < nbp) {
                if (c < ncols_live) Mout[(long long)r * ld + c] = acc[mi][0];
                if (c + 1 < ncols_live) Mout[(long long)r * ld + c + 1] = acc[mi][1];
            }
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// =====================================================================================================
// partial row sums for wide fronts: grid (column chunks, fronts, networks).  The panel kernel adds the
// chunks in a fixed order, so results are deterministic.
// =====================================================================================================
__global__ void __launch_bounds__(1024)
rowsum_kernel(const double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j,
              double *rsbuf, int maxfb, int maxchunks, DFront F0, int has_f0) {
    pdl_enter();
    const DFront F = fetch_front(fronts, ids, blockIdx.y, F0, has_f0);
    if (j >= F.k) return;
    const int nbp = min(NB_INNER, F.k - j);
    const int je = j + nbp;
    if (F.f - je <= RS_SPLIT) return;            // narrow front: the panel kernel sums its own rows
    const int c0 = je + blockIdx.x * RS_CHUNK;
    if (c0 >= F.f) return;
    const int c1 = min(c0 + RS_CHUNK, F.f);
    const double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    double *out = rsbuf + (((long long)blockIdx.z * maxfb + blockIdx.y) * maxchunks + blockIdx.x) * NB_INNER;
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (r >= nbp) return;
    const double *row = M + (long long)(j + r) * F.ld;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = c0 + lane;
    for (; c + 96 < c1; c += 128) { s0 += row[c]; s1 += row[c + 32]; s2 += row[c + 64]; s3 += row[c + 96]; }
    for (; c < c1; c += 32) s0 += row[c];
    const double sum = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) out[r] = sum;
}

void launch_rowsum(const LaunchCtx &c, const int *ids, int nfb, int j, int max_cols) {
    dim3 grid((max_cols + RS_CHUNK - 1) / RS_CHUNK, nfb, c.nbatch);
    launch_chain(rowsum_kernel, grid, dim3(1024), 0, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, c.rsbuf, c.maxfb,
                 c.maxchunks, c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
}

// =====================================================================================================
// panel kernel: one CTA per front and network.
//   s_u  = sum_{v >= je} P[j+u, v]                (mass leaving the panel; GTH row sums)     -- all warps
//   W    = (I - P_XX)^-1 by subtraction-free Gauss-Jordan; pivot d_p = sum_{q > p} T[p,q] + s_p  -- 4 warps
//   U'   = W [P_XR | tau_X] for fronts with few columns                                        -- all warps
// The Gauss-Jordan is blocked (gj_blocked below): 8-pivot diagonal blocks inverted by one warp with the scalar GTH
// recurrence, everything else as 8 x 8 x 8 DMMA products.  T[u][u] is never used: the self-loop mass is discarded and
// every pivot is re-summed from off-diagonal entries only, which keeps the recurrence free of subtractions.
// NT = 512: single-front launches at the top of the assembly tree (16 warps share the row sums and the row panel);
// NT = 128: launches with many fronts (leaf levels, batched networks, phase-2 tasks), 4 CTAs per SM.
// =====================================================================================================
constexpr int GJ_WARPS = 4;
__device__ __forceinline__ void gj_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(GJ_WARPS * 32) : "memory"); }

#ifdef NGT_PANEL_CLOCKS
__device__ long long g_panel_clk[8];
__device__ long long g_gj_clk[8];
#define NGT_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_panel_clk[i] = clock64(); } while (0)
#define NGT_GJ_T(i) do { const long long t_ = clock64(); gj_acc[i] += t_ - gj_last; gj_last = t_; } while (0)
#else
#define NGT_CLK(i) do { } while (0)
#define NGT_GJ_T(i) do { } while (0)
#endif

// -----------------------------------------------------------------------------------------------------
// Blocked subtraction-free Gauss-Jordan: W = (I - T)^-1 of a 32-pivot panel in 8-pivot blocks (4 warps).
//   working matrix G: NB_INNER x GJ_LD in shared memory; columns 0..31 hold T (diagonal ignored), column 32 the mass s
//   that leaves the panel.  Block step b (pivots B = 8b..8b+7):
//     (1) warp 0 inverts the 8 x 8 diagonal block with the scalar GTH recurrence -- pivot d_p = sum_{q>p, q in B} T_pq +
//         sigma_p, sigma_p = the row's mass into later blocks and out of the panel -> W_b = (I - T_BB)^-1;
//     (2) all warps, on the FP64 tensor cores (DMMA m8n8k4): the block row becomes Y = W_b G[B][other columns], every
//         other row tile gets G[r][c] += G[r][B] Y[B][c]  (8 x 8 x 8 products, one 8-column tile of G per warp; the s
//         column rides as a fifth tile), and the block's own columns become G[r][B] W_b: they join the inverse.
//   After the four block steps columns 0..31 of G hold W.  Every quantity is a sum of products of non-negative
//   numbers, as in the scalar recurrence (Wales' 1-P_xx as a sum of off-diagonal probabilities, blocked).
//   The scalar 32-step recurrence of round 1 (4 warps, lane = row, one named barrier per pivot) issued ~37 warp-wide
//   FP64 instructions per pivot at 8 cycles each on all four schedulers: 23.6k cycles per panel measured.  Here the
//   scalar chain runs over 8 columns in one warp and the O(32^2) work per pivot moves to ~40 DMMAs per block: 12.2k
//   cycles for the four inversions + 2.8k for the DMMA steps (csrc/tools/panel_bench.cu, -DNGT_PANEL_CLOCKS).
// -----------------------------------------------------------------------------------------------------
constexpr int GJ_LD = NB_INNER + 4;      // = 4 mod 16 doubles: conflict-free DMMA fragment loads; column NB_INNER = s
constexpr int WB_LD = 12;                // row stride of the 8 x 8 inverse W_b
static_assert(NB_INNER == 32, "the blocked Gauss-Jordan is written for 32-pivot panels (4 warps x one 8-column tile)");

// sum of the row's entries in block-local columns >= t; every lane of the row (lane = r + 8 cg) receives it
__device__ __forceinline__ double gj8_tail(double x0, double x1, int cg, int t) {
    double v = ((2 * cg >= t) ? x0 : 0.0) + ((2 * cg + 1 >= t) ? x1 : 0.0);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// (1): one warp.  lane = r + 8 cg holds row r of the 8 x 8 block, columns 2cg and 2cg+1, and a copy of sigma_r.
// As in the 32-step version the rows stay unnormalised while they are eliminated (row u carries the factor d_u once it
// has been the pivot) and are scaled by 1/d_u at the end; the NEXT pivot is formed one step ahead from sums of the
// state before the current step: d_{p+1} = A + f B, A = sum_{q>p+1} T[p+1][q] + sigma_{p+1}, B = sum_{q>p+1} T[p][q] +
// sigma_p, f = T[p+1][p] / d_p, so the chain reciprocal -> multiply -> fma -> reciprocal is all that is serial.
__device__ __forceinline__ void gj8_invert(const double *G, int b, double *Wb, int lane, int *errflag) {
    constexpr unsigned FULL = 0xffffffffu;
    const int r = lane & 7, cg = lane >> 3;
    const double *g = G + (8 * b + r) * GJ_LD;
    double x0 = (2 * cg == r) ? 0.0 : g[8 * b + 2 * cg];
    double x1 = (2 * cg + 1 == r) ? 0.0 : g[8 * b + 2 * cg + 1];
    double sg = 0.0;
    for (int c = 8 * (b + 1) + cg; c < NB_INNER; c += 4) sg += g[c];
    sg += __shfl_xor_sync(FULL, sg, 8);
    sg += __shfl_xor_sync(FULL, sg, 16);
    sg += g[NB_INNER];
    double RD = 0.0;
    double tl = gj8_tail(x0, x1, cg, 2) + sg;
    double dcur = __shfl_sync(FULL, gj8_tail(x0, x1, cg, 1) + sg, 0);
    double rd = 1.0 / dcur;
    bool bad = false;
    // unrolled on purpose: rolled (runtime p) the inversion measured 18.1k instead of 12.2k cycles per panel
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        bad = bad || !(dcur > 0.0) || isinf(dcur);
        const double xp0 = __shfl_sync(FULL, x0, p + 8 * cg), xp1 = __shfl_sync(FULL, x1, p + 8 * cg);
        const double sp = __shfl_sync(FULL, sg, p);
        const double colv = __shfl_sync(FULL, (p & 1) ? x1 : x0, r + 8 * (p >> 1));
        double dn = 1.0, rdn = 1.0;
        if (p < 7) {
            const double A = __shfl_sync(FULL, tl, p + 1), B = __shfl_sync(FULL, tl, p);
            const double cn = __shfl_sync(FULL, colv, p + 1);
            const double fn = (cn == 0.0) ? 0.0 : cn * rd;
            dn = fma(fn, B, A);
            rdn = 1.0 / dn;
        }
        if (r == p) RD = rd;
        // rows that do not reach p stay clean even if d_p == 0
        const double f = (r == p || colv == 0.0) ? 0.0 : colv * rd;
        x0 = fma(f, xp0, x0);
        x1 = fma(f, xp1, x1);
        sg = fma(f, sp, sg);
        if (cg == (p >> 1)) {                      // column p joins the inverse
            const double v = (r == p) ? 1.0 : f;
            if (p & 1) x1 = v; else x0 = v;
        }
        if (p < 7) tl = gj8_tail(x0, x1, cg, p + 3) + sg;
        rd = rdn;
        dcur = dn;
    }
    if (bad && lane == 0) atomicOr(errflag, 1);
    Wb[r * WB_LD + 2 * cg] = x0 * RD;
    Wb[r * WB_LD + 2 * cg + 1] = x1 * RD;
}

__device__ __forceinline__ void gj_bar2_arrive() { asm volatile("bar.arrive 2, %0;" ::"n"(GJ_WARPS * 32) : "memory"); }
__device__ __forceinline__ void gj_bar2_sync() { asm volatile("bar.sync 2, %0;" ::"n"(GJ_WARPS * 32) : "memory"); }

// all GJ_WARPS warps call this convergently; G is overwritten by W (rows >= nbp zeroed), Wb is scratch (8 x WB_LD)
__device__ __forceinline__ void gj_blocked(double *G, double *Wb, int nbp, int *errflag, int warp, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
#ifdef NGT_PANEL_CLOCKS
    long long gj_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, gj_last = clock64();
#endif
#pragma unroll 1
    for (int b = 0; b < NB_INNER / 8; ++b) {
        if (warp == 0) gj8_invert(G, b, Wb, lane, errflag);
        NGT_GJ_T(0);
        gj_barrier();                                  // W_b is visible
        NGT_GJ_T(1);
        // this warp's column tile: one of the three other 8-column tiles of the working matrix; warp 3: the s column
        const int ct = warp < 3 ? (warp < b ? warp : warp + 1) : 4;
        // operand fragments, read before anything is overwritten
        double wa[2], bb[2], ab[3][2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            wa[ks] = Wb[g * WB_LD + t4 + 4 * ks];
            const double *row = G + (8 * b + t4 + 4 * ks) * GJ_LD;
            bb[ks] = ct < 4 ? row[8 * ct + g] : (g == 0 ? row[NB_INNER] : 0.0);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int rt = i < b ? i : i + 1;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) ab[i][ks] = G[(8 * rt + g) * GJ_LD + 8 * b + t4 + 4 * ks];
        }
        if (warp != 3) gj_bar2_arrive();               // "my copy of the old column block is in registers"
        // the block row: Y = W_b G[B][ct]
        double y0 = 0.0, y1 = 0.0;
        dmma884(y0, y1, wa[0], bb[0]);
        dmma884(y0, y1, wa[1], bb[1]);
        {
            double *row = G + (8 * b + g) * GJ_LD;
            if (ct < 4) { row[8 * ct + 2 * t4] = y0; row[8 * ct + 2 * t4 + 1] = y1; }
            else if (t4 == 0) row[NB_INNER] = y0;
        }
        __syncwarp();
        double yb[2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const double *row = G + (8 * b + t4 + 4 * ks) * GJ_LD;
            yb[ks] = ct < 4 ? row[8 * ct + g] : (g == 0 ? row[NB_INNER] : 0.0);
        }
        // every other row tile: G[rt][ct] += G[rt][B] Y
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int rt = i < b ? i : i + 1;
            double *row = G + (8 * rt + g) * GJ_LD;
            double c0, c1;
            if (ct < 4) { c0 = row[8 * ct + 2 * t4]; c1 = row[8 * ct + 2 * t4 + 1]; }
            else { c0 = (t4 == 0) ? row[NB_INNER] : 0.0; c1 = 0.0; }
            dmma884(c0, c1, ab[i][0], yb[0]);
            dmma884(c0, c1, ab[i][1], yb[1]);
            if (ct < 4) { row[8 * ct + 2 * t4] = c0; row[8 * ct + 2 * t4 + 1] = c1; }
            else if (t4 == 0) row[NB_INNER] = c0;
        }
        if (warp == 3) {
            gj_bar2_sync();                            // the other warps hold their copies of the old column block
            // column block b joins the inverse: G[rt][B] = G_old[rt][B] W_b, G[B][B] = W_b
            double wf[2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) wf[ks] = Wb[(t4 + 4 * ks) * WB_LD + g];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int rt = i < b ? i : i + 1;
                double c0 = 0.0, c1 = 0.0;
                dmma884(c0, c1, ab[i][0], wf[0]);
                dmma884(c0, c1, ab[i][1], wf[1]);
                double *row = G + (8 * rt + g) * GJ_LD + 8 * b;
                row[2 * t4] = c0;
                row[2 * t4 + 1] = c1;
            }
            double *row = G + (8 * b + g) * GJ_LD + 8 * b;
            row[2 * t4] = Wb[g * WB_LD + 2 * t4];
            row[2 * t4 + 1] = Wb[g * WB_LD + 2 * t4 + 1];
        }
        NGT_GJ_T(2);
        gj_barrier();                                  // step b complete (also protects W_b against the next inversion)
        NGT_GJ_T(3);
    }
#ifdef NGT_PANEL_CLOCKS
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int i = 0; i < 8; ++i) g_gj_clk[i] = gj_acc[i];
#endif
    // padding rows (panels with fewer than 32 pivots) were treated as isolated absorbing nodes: clear them
    for (int i = warp * 32 + lane; i < (NB_INNER - nbp) * NB_INNER; i += GJ_WARPS * 32)
        G[(nbp + i / NB_INNER) * GJ_LD + (i % NB_INNER)] = 0.0;
}


// CL > 1: the launch is made of thread-block clusters of CL CTAs per front (single fronts at the top of the assembly
// tree, where the whole GPU waits on this chain): all CTAs sum one column chunk of the 32 panel rows each, CTA 0 runs
// the Gauss-Jordan, then all CTAs share the slabs of the row panel -- the stand-alone row-sum and row-panel launches
// disappear from the chain.  Hand-over between the CTAs goes through global memory (rsbuf, wbuf) ordered by
// release/acquire cluster barriers.
template <int NT, int CL>
__global__ void __launch_bounds__(NT, (NT >= 512 ? 1 : 4))
panel_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j,
             double *wbuf, double *rsbuf, int maxfb, int maxchunks, int *errflag, KOpt ko, DFront F0, int has_f0) {
    pdl_enter();
    constexpr int NW = NT / 32;              // warps
    constexpr int CW = 64;                   // row-panel columns per slab
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int slot = CL > 1 ? (int)blockIdx.x / CL : (int)blockIdx.x;
    const DFront F = fetch_front(fronts, ids, slot, F0, has_f0);
    if (j >= F.k) return;                    // uniform over a cluster
    double *M = pool + (long long)blockIdx.y * pool_stride + F.off;
    const int nbp = min(NB_INNER, F.k - j);
    const int je = j + nbp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    __shared__ double Xs[NB_INNER][GJ_LD];          // T (and s in column NB_INNER) on the way in, W on the way out
    __shared__ double Wb8[8 * WB_LD];
    __shared__ double xs[NB_INNER * (CW + 4)];
    __shared__ double Wd[NB_INNER * WD_LD];
    NGT_CLK(0);

    // row sums over everything to the right of the panel (tau column excluded) and the T block.  A warp owns rows
    // warp, warp + NW, ...; their loads are issued together so one round of memory latency covers all of them.
    constexpr int RPW = NB_INNER / NW;
    double *rs = rsbuf + (((long long)blockIdx.y * maxfb + slot) * maxchunks) * NB_INNER;
    // sum of columns [ca, ce) of the warp's rows (fixed order: four strided partial sums, then the lanes)
    auto chunk_sums = [&](int ca, int ce, double (&sum)[RPW]) {
        double s0[RPW], s1[RPW], s2[RPW], s3[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i) s0[i] = s1[i] = s2[i] = s3[i] = 0.0;
        const double *base = M + (long long)j * F.ld;
        int c = ca + lane;
        for (; c + 96 < ce; c += 128) {
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int u = warp + i * NW;
                if (u < nbp) {
                    const double *row = base + (long long)u * F.ld;
                    s0[i] += row[c]; s1[i] += row[c + 32]; s2[i] += row[c + 64]; s3[i] += row[c + 96];
                }
            }
        }
        for (; c < ce; c += 32) {
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int u = warp + i * NW;
                if (u < nbp) s0[i] += base[(long long)u * F.ld + c];
            }
        }
#pragma unroll
        for (int i = 0; i < RPW; ++i) sum[i] = warp_sum((s0[i] + s1[i]) + (s2[i] + s3[i]));
    };
    // sum of `nch` partial sums per row left in rs by other CTAs (this cluster's, or rowsum_kernel's for wide fronts)
    auto partial_sums = [&](int nch, double (&sum)[RPW]) {
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const int u = warp + i * NW;
            double a = 0.0;
            if (u < nbp)
                for (int ch = lane; ch < nch; ch += 32) a += __ldcg(rs + ch * NB_INNER + u);
            sum[i] = warp_sum(a);
        }
    };
    double rowsum[RPW];
    if (CL > 1) {
        const int cw = (((F.f - je + CL - 1) / CL) + 31) & ~31;
        const int ca = min(je + crank * cw, F.f);
        chunk_sums(ca, min(ca + cw, F.f), rowsum);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < RPW; ++i) rs[crank * NB_INNER + warp + i * NW] = rowsum[i];
        }
        cluster_sync_all();
    }
    if (crank == 0) {
        if (CL > 1) partial_sums(CL, rowsum);
        else if (F.f - je > RS_SPLIT) partial_sums((F.f - je + RS_CHUNK - 1) / RS_CHUNK, rowsum);
        else chunk_sums(je, F.f, rowsum);
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
            const int u = warp + i * NW;
            const double sum = (u < nbp) ? rowsum[i] : 1.0;   // padding rows behave like isolated absorbing nodes
            const double t = (u < nbp && lane < nbp && lane != u) ? M[(long long)(j + u) * F.ld + j + lane] : 0.0;
            Xs[u][lane] = t;
            if (lane < GJ_LD - NB_INNER) Xs[u][NB_INNER + lane] = (lane == 0) ? sum : 0.0;
        }
    }
    __syncthreads();
    NGT_CLK(1);

    if (crank == 0 && warp < GJ_WARPS) gj_blocked(&Xs[0][0], Wb8, nbp, errflag, warp, lane);
    NGT_CLK(2);
    __syncthreads();
    NGT_CLK(3);
    double *Wout = wbuf + ((long long)blockIdx.y * maxfb + slot + ko.wslot) * (NB_INNER * NB_INNER);
    if (crank == 0)
        for (int i = threadIdx.x; i < NB_INNER * NB_INNER; i += NT) Wout[i] = Xs[i >> 5][i & 31];
    if (CL > 1) cluster_sync_all();

    // fused row panel (narrow fronts; always in the cluster variant): columns je..f (f = tau column) in slabs of CW
    // staged through smem; the CTAs of a cluster take the slabs round-robin
    const int ncols = F.f + 1 - je;
    NGT_CLK(4);
    if (CL == 1 && ncols > ko.fuse_cols) return;
    if (crank == 0) {
        for (int i = threadIdx.x; i < NB_INNER * NB_INNER; i += NT) Wd[(i >> 5) * WD_LD + (i & 31)] = Xs[i >> 5][i & 31];
    } else {
        for (int i = threadIdx.x; i < NB_INNER * NB_INNER; i += NT) Wd[(i >> 5) * WD_LD + (i & 31)] = __ldcg(Wout + i);
    }
    for (int cs = crank * CW; cs < ncols; cs += CL * CW) {
        __syncthreads();
        for (int i = threadIdx.x; i < NB_INNER * CW; i += NT) {
            const int r = i / CW, cc = i % CW;
            xs[r * (CW + 4) + cc] = (cs + cc < ncols && r < nbp) ? M[(long long)(j + r) * F.ld + je + cs + cc] : 0.0;
        }
        __syncthreads();
        rowpanel_slab_mma<CW, NW>(Wd, xs, M + (long long)j * F.ld + je + cs, F.ld, nbp, min(CW, ncols - cs), warp, lane);
    }
    NGT_CLK(5);
}

void launch_panel(const LaunchCtx &c, const int *ids, int nfb, int j) {
    dim3 grid(nfb, c.nbatch);
    if (c.panel_cluster) {       // few fronts: a cluster of CTAs per front also does the row sums and the row panel
        grid.x = nfb * PANEL_CLUSTER;
        launch_chain_cluster(panel_kernel<512, PANEL_CLUSTER>, grid, dim3(512), 0, c.stream, PANEL_CLUSTER, c.pool, c.pool_stride,
                             c.fronts, ids, j, c.wbuf, c.rsbuf, c.maxfb, c.maxchunks, c.errflag, c.ko,
                             c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
    } else if ((long long)nfb * c.nbatch >= 2 * 148)   // many panels: favour occupancy over single-panel latency
        launch_chain(panel_kernel<128, 1>, grid, dim3(128), 0, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, c.wbuf,
                     c.rsbuf, c.maxfb, c.maxchunks, c.errflag, c.ko, c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
    else
        launch_chain(panel_kernel<512, 1>, grid, dim3(512), 0, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, c.wbuf,
                     c.rsbuf, c.maxfb, c.maxchunks, c.errflag, c.ko, c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
}

// =====================================================================================================
// row panel for wide fronts: U' = W [P_XR | tau_X] in place, columns je..f (f = tau column), on the tensor cores.
// One CTA (8 warps) per slab of 64 columns: stage W and the slab in shared memory, one 32 x 8 x 32 DMMA product per warp.
// =====================================================================================================
constexpr int RP_CW = 64;
__global__ void __launch_bounds__(256)
rowpanel_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j,
                const double *wbuf, int maxfb, KOpt ko, DFront F0, int has_f0) {
    pdl_enter();
    const DFront F = fetch_front(fronts, ids, blockIdx.y, F0, has_f0);
    if (j >= F.k) return;
    const int nbp = min(NB_INNER, F.k - j);
    const int je = j + nbp;
    const int ncols = F.f + 1 - je;
    if (ncols <= ko.fuse_cols) return;    // done inside the panel kernel
    const int cs = blockIdx.x * RP_CW;
    if (cs >= ncols) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;

    __shared__ double Wd[NB_INNER * WD_LD];
    __shared__ double xs[NB_INNER * (RP_CW + 4)];
    const double *W = wbuf + ((long long)blockIdx.z * maxfb + blockIdx.y + ko.wslot) * (NB_INNER * NB_INNER);
    for (int i = threadIdx.x; i < NB_INNER * NB_INNER; i += 256) Wd[(i >> 5) * WD_LD + (i & 31)] = W[i];
    for (int i = threadIdx.x; i < NB_INNER * RP_CW; i += 256) {
        const int r = i / RP_CW, cc = i % RP_CW;
        const int col = je + cs + cc;
        const bool live = cs + cc < ncols && r < nbp && owned_col(ko, col, F.f);
        xs[r * (RP_CW + 4) + cc] = live ? M[(long long)(j + r) * F.ld + col] : 0.0;
    }
    __syncthreads();
    if (ko.own_G > 1) {
        // multi-GPU dense path: only owned columns are rewritten (ownership is uniform over a slab's 8-column groups
        // except at block edges, so fall back to a guarded scalar write-back)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
        double acc[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) acc[mi][0] = acc[mi][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
            const double b = xs[(k4 * 4 + t4) * (RP_CW + 4) + warp * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) dmma884(acc[mi][0], acc[mi][1], Wd[(mi * 8 + g) * WD_LD + k4 * 4 + t4], b);
        }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int r = mi * 8 + g;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = je + cs + warp * 8 + 2 * t4 + e;
                if (r < nbp && col <= F.f && owned_col(ko, col, F.f)) M[(long long)(j + r) * F.ld + col] = acc[mi][e];
            }
        }
        return;
    }
    rowpanel_slab_mma<RP_CW, 8>(Wd, xs, M + (long long)j * F.ld + je + cs, F.ld, nbp, min(RP_CW, ncols - cs),
                                threadIdx.x >> 5, threadIdx.x & 31);
}

void launch_rowpanel(const LaunchCtx &c, const int *ids, int nfb, int j, int max_cols) {
    dim3 grid((max_cols + RP_CW - 1) / RP_CW, nfb, c.nbatch);
    if (grid.x == 0) grid.x = 1;
    launch_chain(rowpanel_kernel, grid, dim3(256), 0, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, c.wbuf, c.maxfb, c.ko,
                 c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
}

// =====================================================================================================
// Schur update, scalar DFMA path (debug / cross-check): C[r,c] += sum_p F[r,p] F[p,c]
// 64 x 64 tile, 256 threads, 4 x 4 per thread
// =====================================================================================================
constexpr int FT = 64, FK = 16;

__global__ void __launch_bounds__(256)
update_dfma_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j, int mode_in,
                   int split, KOpt ko) {
    const DFront F = fronts[ids[blockIdx.y]];
    Region R;
    int mode = mode_in, tile = blockIdx.x;
    if (mode_in == MODE_AB) {
        if (tile < split) mode = MODE_A; else { mode = MODE_B; tile -= split; }
    } else if (mode_in == MODE_D1) {
        if (tile < split) mode = MODE_D1A; else { mode = MODE_D1B; tile -= split; }
    }
    if (!get_region(F.f, F.fr, F.k, j, mode, R)) return;
    const int tiles_c = (R.c1 - R.c0 + FT - 1) / FT;
    const int tiles_r = (R.r1 - R.r0 + FT - 1) / FT;
    const int tr = tile / tiles_c, tc = tile % tiles_c;
    if (tr >= tiles_r) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    const double *MA = F.a_ld ? pool + (long long)blockIdx.z * pool_stride + F.a_off : M;
    const long long lda = F.a_ld ? F.a_ld : F.ld;
    const int rb = R.r0 + tr * FT, cb = R.c0 + tc * FT;
    if (ko.own_G > 1 && !owned_col(ko, cb, F.f) && !owned_col(ko, min(cb + FT, R.c1) - 1, F.f)) return;

    __shared__ double As[FT][FK + 1];
    __shared__ double Bs[FK][FT + 1];
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    for (int kb = R.k0; kb < R.k1; kb += FK) {
        for (int i = threadIdx.x; i < FT * FK; i += 256) {
            const int r = i / FK, kk = i % FK;
            As[r][kk] = (rb + r < R.r1 && kb + kk < R.k1) ? MA[(long long)(rb + r) * lda + kb + kk] : 0.0;
        }
        for (int i = threadIdx.x; i < FK * FT; i += 256) {
            const int kk = i / FT, cc = i % FT;
            Bs[kk][cc] = (cb + cc < R.c1 && kb + kk < R.k1) ? M[(long long)(kb + kk) * F.ld + cb + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < FK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[ty * 4 + i][kk]; b[i] = Bs[kk][tx + 16 * i]; }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] += a[x] * b[y];
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        const int r = rb + ty * 4 + x;
        if (r >= R.r1) continue;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            const int cidx = cb + tx + 16 * y;
            if (cidx < R.c1 && owned_col(ko, cidx, F.f)) M[(long long)r * F.ld + cidx] += acc[x][y];
        }
    }
}

// =====================================================================================================
// Schur update on the FP64 tensor cores: DMMA m8n8k4 (the sm_100a FP64 tensor instruction; tcgen05 has
// no f64 kind).  CTA tile 128 x 64, 8 warps as 4 (M) x 2 (N), warp tile 32 x 32 = 4 x 4 MMA tiles.
// K is staged 16 at a time through a 3-stage cp.async ring (global -> shared, no register staging); two CTAs
// are resident per SM so one CTA's prologue / epilogue hides behind the other's DMMAs.  The accumulators start
// from the C tile itself (loaded while the pipeline fills), so the epilogue is a plain store.
// Shared strides 20 / 68 doubles make both fragment loads bank-conflict free (stride = 4 mod 16).
// =====================================================================================================
constexpr int TM = 128, TN = 64, TK = 16, TSTG = 3;
constexpr int AS_LD = TK + 4, BS_LD = TN + 4;
constexpr int TSTAGE_DOUBLES = TM * AS_LD + TK * BS_LD;

#ifdef NGT_PANEL_CLOCKS
__device__ long long g_upd_clk[8];
#define NGT_UCLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_upd_clk[i] = clock64(); } while (0)
#else
#define NGT_UCLK(i) do { } while (0)
#endif
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 8 : 0;   // 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cp_async16(double *smem_dst, const double *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0;   // 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}

// TMt = 128: the throughput tile.  TMt = 64: same kernel with half-height tiles (8 warps as 2 x 4, warp tile 32 x 16)
// for launches that would otherwise leave most SMs idle -- the thin (K = 32) updates at the top of the assembly tree
// are bound by each SM's L2 port (C tile in + out), so more, smaller CTAs finish sooner.
template <int TMt>
__global__ void __launch_bounds__(256, 2)
update_dmma_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j, int mode_in,
                   int split, KOpt ko, DFront F0, int has_f0) {
    constexpr int WM = TMt / 32, WN = 8 / WM, NI = TN / (8 * WN);
    pdl_enter();
    NGT_UCLK(0);
    const DFront F = fetch_front(fronts, ids, blockIdx.y, F0, has_f0);
    Region R;
    int mode = mode_in, tile = blockIdx.x;
    if (mode_in == MODE_AB) {
        if (tile < split) mode = MODE_A; else { mode = MODE_B; tile -= split; }
    } else if (mode_in == MODE_D1) {
        if (tile < split) mode = MODE_D1A; else { mode = MODE_D1B; tile -= split; }
    }
    if (!get_region(F.f, F.fr, F.k, j, mode, R)) return;
    const int tiles_c = (R.c1 - R.c0 + TN - 1) / TN;
    const int tiles_r = (R.r1 - R.r0 + TMt - 1) / TMt;
    const int tr = tile / tiles_c, tc = tile % tiles_c;
    if (tr >= tiles_r) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    const double *MA = F.a_ld ? pool + (long long)blockIdx.z * pool_stride + F.a_off : M;
    const long long lda = F.a_ld ? F.a_ld : F.ld;
    const int rb = R.r0 + tr * TMt, cb = R.c0 + tc * TN;
    const long long ld = F.ld;
    if (ko.own_G > 1 && !owned_col(ko, cb, F.f) && !owned_col(ko, min(cb + TN, R.c1) - 1, F.f)) return;

    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp % WM, wn = warp / WM;
    const int g = lane >> 2, t4 = lane & 3;

    // accumulators = C tile (issued first: their latency hides behind the pipeline prologue); 16-byte accesses when
    // the tile starts on an even column (rows are 64-byte aligned) and every column is owned
    const bool vec_c = ((cb & 1) == 0) && ko.own_G <= 1;
    double acc[4][NI][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int r = rb + wm * 32 + mi * 8 + g;
        const double *row = M + (long long)r * ld;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int c0 = cb + wn * (8 * NI) + ni * 8 + 2 * t4;
            if (vec_c && r < R.r1 && c0 + 1 < R.c1) {
                const double2 v = *reinterpret_cast<const double2 *>(row + c0);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            } else {
                acc[mi][ni][0] = (r < R.r1 && c0 < R.c1) ? row[c0] : 0.0;
                acc[mi][ni][1] = (r < R.r1 && c0 + 1 < R.c1) ? row[c0 + 1] : 0.0;
            }
        }
    }

    NGT_UCLK(1);
    const int nk = (R.k1 - R.k0 + TK - 1) / TK;
    // Operand staging.  Interior tiles on even columns with an even pivot count move 16 bytes per cp.async:
    //   A chunk (row ar2 + 32 i, pivots kb + ak2 .. +1), B chunk (pivot kb + bk2 + 8 i, columns bc2 .. +1);
    // everything else moves single doubles: A element (row ar + 16 i, pivot kb + ak), B element (pivot kb + bk + 4 i,
    // column bc).  Out-of-range slots are zero-filled either way.
    const bool vec_ab = ((cb & 1) == 0) && (cb + TN <= R.c1) && (((R.k1 - R.k0) & 1) == 0) && F.a_ld == 0;
    const int ar = threadIdx.x >> 4, ak = threadIdx.x & 15;
    const int bk = threadIdx.x >> 6, bc = threadIdx.x & 63;
    const int ar2 = threadIdx.x >> 3, ak2 = (threadIdx.x & 7) * 2;
    const int bk2 = threadIdx.x >> 5, bc2 = (threadIdx.x & 31) * 2;
    const double *pa = vec_ab ? MA + (long long)(rb + ar2) * lda + ak2 : MA + (long long)(rb + ar) * lda + ak;
    const double *pb = vec_ab ? M + (long long)bk2 * ld + cb + bc2 : M + (long long)bk * ld + cb + bc;
    unsigned amask = 0;
#pragma unroll
    for (int i = 0; i < TMt / 16; ++i) amask |= (rb + (vec_ab ? ar2 + 32 * i : ar + 16 * i) < R.r1) ? (1u << i) : 0u;
    const bool bcol_ok = cb + bc < R.c1;
    double *sa = vec_ab ? smem + ar2 * AS_LD + ak2 : smem + ar * AS_LD + ak;
    double *sb = vec_ab ? smem + TM * AS_LD + bk2 * BS_LD + bc2 : smem + TM * AS_LD + bk * BS_LD + bc;
    auto load_stage = [&](int stg, int kb) {
        if (vec_ab) {
            const bool ka = kb + ak2 < R.k1;
#pragma unroll
            for (int i = 0; i < TMt / 32; ++i) {
                const bool ok = ka && ((amask >> i) & 1u);
                cp_async16(sa + stg * TSTAGE_DOUBLES + i * 32 * AS_LD, ok ? pa + (long long)i * 32 * lda + kb : M, ok);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool ok = kb + bk2 + 8 * i < R.k1;
                cp_async16(sb + stg * TSTAGE_DOUBLES + i * 8 * BS_LD, ok ? pb + (long long)(kb + 8 * i) * ld : M, ok);
            }
        } else {
            const bool ka = kb + ak < R.k1;
#pragma unroll
            for (int i = 0; i < TMt / 16; ++i) {
                const bool ok = ka && ((amask >> i) & 1u);
                cp_async8(sa + stg * TSTAGE_DOUBLES + i * 16 * AS_LD, ok ? pa + (long long)i * 16 * lda + kb : M, ok);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = bcol_ok && (kb + bk + 4 * i < R.k1);
                cp_async8(sb + stg * TSTAGE_DOUBLES + i * 4 * BS_LD, ok ? pb + (long long)(kb + 4 * i) * ld : M, ok);
            }
        }
    };
#pragma unroll
    for (int st = 0; st < TSTG - 1; ++st) {
        if (st < nk) load_stage(st, R.k0 + st * TK);
        cp_async_commit();
    }
    NGT_UCLK(2);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<TSTG - 2>();
        __syncthreads();
        if (kt == 0) NGT_UCLK(3);
        const int nxt = kt + TSTG - 1;
        if (nxt < nk) load_stage(nxt % TSTG, R.k0 + nxt * TK);
        cp_async_commit();
        const double *As = smem + (kt % TSTG) * TSTAGE_DOUBLES;
        const double *Bs = As + TM * AS_LD;
#pragma unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            double a[4], b[NI];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[(wm * 32 + mi * 8 + g) * AS_LD + k4 * 4 + t4];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) b[ni] = Bs[(k4 * 4 + t4) * BS_LD + wn * (8 * NI) + ni * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    NGT_UCLK(4);
    // epilogue: thread holds C[g][2*t4 + {0,1}] of every 8 x 8 tile
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int r = rb + wm * 32 + mi * 8 + g;
        if (r >= R.r1) continue;
        double *row = M + (long long)r * ld;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int c0 = cb + wn * (8 * NI) + ni * 8 + 2 * t4;
            if (vec_c && c0 + 1 < R.c1) {
                *reinterpret_cast<double2 *>(row + c0) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            } else {
                if (c0 < R.c1 && owned_col(ko, c0, F.f)) row[c0] = acc[mi][ni][0];
                if (c0 + 1 < R.c1 && owned_col(ko, c0 + 1, F.f)) row[c0 + 1] = acc[mi][ni][1];
            }
        }
    }
    NGT_UCLK(5);
}

// mbarrier / bulk-copy helpers of the tensor-map fed Schur kernel below
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk copies of the copy engine (TMA): global -> shared landing on an mbarrier, shared -> global in a bulk group
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// =====================================================================================================
// Schur update fed by 2-D tensor-map TMA (`cp.async.bulk.tensor.2d` -> SASS UTMALDG) for large single-front
// trailing updates.  Per 16-pivot stage the producer warp issues FIVE copies: one 16 x 128 box of A (16 KB) and four
// 16 x 16 boxes of B (2 KB each; a 128-byte-swizzled box may not be wider than 128 bytes), all landing on the stage's
// "full" mbarrier; 4 stages (96 KB) x 2 CTAs per SM.  Tiles sit densely in shared memory with the hardware
// 128-byte swizzle (16-byte chunk c of row r is stored at chunk c ^ (r & 7)), which makes the A fragment loads
// conflict free and the B fragment loads 2-way (4 wavefronts instead of 2).  The math warps only wait, LDS and DMMA.
// Tile geometry and accumulation order equal update_dmma_kernel, so results are bitwise identical.
// =====================================================================================================
constexpr int TMS = 4;                                  // stages
constexpr int TM_A_DOUBLES = TM * TK;                   // 128 x 16
constexpr int TM_B_DOUBLES = TK * TN;                   // 16 x 64 as four 16 x 16 sub-tiles
constexpr int TM_STAGE_DOUBLES = TM_A_DOUBLES + TM_B_DOUBLES;
constexpr unsigned TM_STAGE_BYTES = TM_STAGE_DOUBLES * 8;

__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(256, 2)
update_dmma_tmap_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j, int mode_in,
                        int split, KOpt ko, const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB) {
    pdl_enter();
    const DFront F = fronts[ids[blockIdx.y]];
    Region R;
    int mode = mode_in, tile = blockIdx.x;
    if (mode_in == MODE_AB) {
        if (tile < split) mode = MODE_A; else { mode = MODE_B; tile -= split; }
    } else if (mode_in == MODE_D1) {
        if (tile < split) mode = MODE_D1A; else { mode = MODE_D1B; tile -= split; }
    }
    if (!get_region(F.f, F.fr, F.k, j, mode, R)) return;
    const int tiles_c = (R.c1 - R.c0 + TN - 1) / TN;
    const int tiles_r = (R.r1 - R.r0 + TM - 1) / TM;
    const int tr = tile / tiles_c, tc = tile % tiles_c;
    if (tr >= tiles_r) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    const int rb = R.r0 + tr * TM, cb = R.c0 + tc * TN;
    const long long ld = F.ld;

    extern __shared__ __align__(1024) double smem[];
    __shared__ unsigned long long full_bar[TMS], empty_bar[TMS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < TMS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nk = (R.k1 - R.k0) / TK;                      // host guarantees K % 16 == 0

    // the copy engine is driven by one elected thread of warp 0, which is also a math warp: issuing a stage is five
    // instructions, so a dedicated producer warp would only cost registers and occupancy (2 CTAs/SM need <= 8 warps)
    auto issue_stage = [&](int kt) {
        const int stg = kt % TMS;
        if (kt >= TMS) mbar_wait(&empty_bar[stg], ((kt / TMS) - 1) & 1);   // all 8 warps are done with its last use
        const int kb = R.k0 + kt * TK;
        double *As = smem + stg * TM_STAGE_DOUBLES;
        double *Bs = As + TM_A_DOUBLES;
        mbar_expect_tx(&full_bar[stg], TM_STAGE_BYTES);
        tma_load_2d(As, &tmA, kb, rb, &full_bar[stg]);                      // A[rb.., kb..kb+16)
#pragma unroll
        for (int sb = 0; sb < TN / 16; ++sb)                                 // B[kb..kb+16, cb+16 sb ..)
            tma_load_2d(Bs + sb * 256, &tmB, cb + 16 * sb, kb, &full_bar[stg]);
    };
    if (threadIdx.x == 0)
        for (int kt = 0; kt < TMS - 1 && kt < nk; ++kt) issue_stage(kt);

    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, t4 = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int r = rb + wm * 32 + mi * 8 + g;
        const double *row = M + (long long)r * ld;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int c0 = cb + wn * 32 + ni * 8 + 2 * t4;
            acc[mi][ni][0] = (r < R.r1 && c0 < R.c1) ? row[c0] : 0.0;
            acc[mi][ni][1] = (r < R.r1 && c0 + 1 < R.c1) ? row[c0 + 1] : 0.0;
        }
    }
    // swizzled fragment addresses: a lane-constant base plus compile-time offsets.
    //   A[m][k]: m = wm*32 + mi*8 + g, k = 4 k4 + t4 -> doubles m*16 + 2*((k>>1) ^ g) + (k&1); (k>>1) ^ g = (g ^ (t4>>1)) ^ 2 k4
    //   B[k][n]: n = wn*32 + ni*8 + g -> sub-tile n>>4, column c = (ni&1)*8 + g -> doubles
    //            (n>>4)*256 + k*16 + 2*((c>>1) ^ (k&7)) + (c&1);  (c>>1) ^ (k&7) = 4*((ni ^ k4) & 1) + ((g>>1) ^ t4)
    const int a_base = (wm * 32 + g) * 16 + (t4 & 1);
    const int ga = g ^ (t4 >> 1);
    const int b_base = wn * 512 + t4 * 16 + ((((g >> 1) ^ t4) << 1) | (g & 1));
    for (int kt = 0; kt < nk; ++kt) {
        const int stg = kt % TMS;
        if (threadIdx.x == 0 && kt + TMS - 1 < nk) issue_stage(kt + TMS - 1);
        mbar_wait(&full_bar[stg], (kt / TMS) & 1);
        const double *As = smem + stg * TM_STAGE_DOUBLES;
        const double *Bs = As + TM_A_DOUBLES;
#pragma unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[a_base + mi * 128 + ((ga ^ (2 * k4)) << 1)];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[b_base + (ni >> 1) * 256 + k4 * 64 + (((ni ^ k4) & 1) << 3)];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stg]);
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const int r = rb + wm * 32 + mi * 8 + g;
        if (r >= R.r1) continue;
        double *row = M + (long long)r * ld;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int c0 = cb + wn * 32 + ni * 8 + 2 * t4;
            if (c0 < R.c1 && owned_col(ko, c0, F.f)) row[c0] = acc[mi][ni][0];
            if (c0 + 1 < R.c1 && owned_col(ko, c0 + 1, F.f)) row[c0 + 1] = acc[mi][ni][1];
        }
    }
}

void launch_update_tmap(const LaunchCtx &c, const int *ids, int j, int mode, int max_tiles, const void *tmA,
                        const void *tmB) {
    if (max_tiles <= 0) return;
    dim3 grid(max_tiles, 1, c.nbatch);
    const size_t smem = (size_t)TMS * TM_STAGE_BYTES;
    launch_chain(update_dmma_tmap_kernel, grid, dim3(256), smem, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, mode, 0,
                 c.ko, *static_cast<const CUtensorMap *>(tmA), *static_cast<const CUtensorMap *>(tmB));
}

int update_tile_m(int gemm_path) { return gemm_path == 1 ? FT : TM; }
int update_tile_n(int gemm_path) { return gemm_path == 1 ? FT : TN; }
int update_tile_k() { return TK; }

__global__ void dense_network_kernel(const double *K, int n, const int *node_at, int k, double *pool, long long pool_stride,
                                     DFront F0, DFront RF, double *tau0, int *errflag);

void init_kernel_attributes() {
    cudaFuncSetAttribute(dense_network_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_network_smem(DENSE_NET_MAX_N));
    cudaFuncSetAttribute(update_dmma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)TSTG * TSTAGE_DOUBLES * sizeof(double)));
    cudaFuncSetAttribute(update_dmma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)TSTG * TSTAGE_DOUBLES * sizeof(double)));
    cudaFuncSetAttribute(update_dmma_tmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)TMS * TM_STAGE_BYTES));
    cudaFuncSetAttribute(update_dmma_tmap_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);   // two 97 KB CTAs per SM
}

void launch_update(const LaunchCtx &c, const int *ids, int nfb, int j, int mode, int max_tiles, int split) {
    if (max_tiles <= 0) return;
    dim3 grid(max_tiles, nfb, c.nbatch);
    if (c.gemm_path == 1) {
        update_dfma_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, mode, split, c.ko);
    } else {
        const size_t smem = (size_t)TSTG * TSTAGE_DOUBLES * sizeof(double);
        if (c.tile_m == 64)
            launch_chain(update_dmma_kernel<64>, grid, dim3(256), smem, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, mode,
                         split, c.ko, c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
        else
            launch_chain(update_dmma_kernel<128>, grid, dim3(256), smem, c.stream, c.pool, c.pool_stride, c.fronts, ids, j, mode,
                         split, c.ko, c.f0 ? *c.f0 : DFront{}, c.f0 ? 1 : 0);
    }
}

// =====================================================================================================
// ingest
// =====================================================================================================
// sparse: one warp per source node.  tau_u = 1 / sum_v k_uv; P_uv = k_uv tau_u scattered to its front slot.
__global__ void __launch_bounds__(256)
ingest_sparse_kernel(double *pool, long long pool_stride, const double *k, long long k_stride, int n_rows,
                     const long long *row_ptr, const int *row_node, const int *korig, const long long *dest,
                     const long long *tau_dest, double *tau0, long long tau0_stride, int *errflag) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const long long b = blockIdx.y;
    const double *kb = k + b * k_stride;
    double *P = pool + b * pool_stride;
    const long long e0 = row_ptr[row], e1 = row_ptr[row + 1];
    double s = 0.0;
    for (long long e = e0 + lane; e < e1; e += 32) s += kb[korig[e]];
    s = warp_sum(s);
    const double tau = 1.0 / s;
    if (lane == 0) {
        if (!(s > 0.0) || isinf(tau)) atomicOr(errflag, 2);
        const int u = row_node ? row_node[row] : row;     // null: every node is a row (device-built slot table)
        tau0[b * tau0_stride + u] = tau;
        if (tau_dest[row] >= 0) P[tau_dest[row]] = tau;
    }
    for (long long e = e0 + lane; e < e1; e += 32)
        if (dest[e] >= 0) P[dest[e]] = kb[korig[e]] * tau;
}

void launch_ingest_sparse(double *pool, long long pool_stride, int nbatch, const double *k, long long k_stride,
                          int n_rows, const long long *row_ptr, const int *row_node, const int *korig,
                          const long long *dest, const long long *tau_dest, double *tau0, long long tau0_stride,
                          int *errflag, cudaStream_t s) {
    if (n_rows <= 0) return;
    dim3 grid((n_rows + 7) / 8, nbatch);
    ingest_sparse_kernel<<<grid, 256, 0, s>>>(pool, pool_stride, k, k_stride, n_rows, row_ptr, row_node, korig, dest,
                                             tau_dest, tau0, tau0_stride, errflag);
}

// Rate -> front-slot table built on the device (networks without duplicate rates, absent or pruned nodes; the host
// builds the table otherwise).  CSR by source node as produced by the host's edge index: row u = node u, slot q holds
// entry korig[q].  dest[q] = pool offset of P[u][v] inside the front that owns the earlier of the two nodes;
// tau_dest[u] = the tau slot of u's own row.  One warp per source node.
__device__ __forceinline__ int slot_local_index(const DFront &F, int first, const int *idx, int p) {
    if (F.k > 0 && p >= first && p < first + F.k) return p - first;
    int lo = 0, hi = F.f - F.k;                     // update set: ascending elimination positions
    const int *b = idx + F.k;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (b[mid] < p) lo = mid + 1; else hi = mid; }
    return (lo < F.f - F.k && b[lo] == p) ? F.k + lo : -1;
}

__global__ void __launch_bounds__(256)
slot_table_kernel(int n, const long long *row_ptr, const int *korig, const long long *dst, const int *pos,
                  const int *pos_front, const DFront *fronts, const int *ffirst, const int *fidx, const long long *idx_off,
                  long long *dest, long long *tau_dest, int *errflag) {
    const int u = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const int pu = pos[u];
    if (lane == 0) {
        const int owner = pos_front[pu];
        const DFront F = fronts[owner];
        const int lr = slot_local_index(F, ffirst[owner], fidx + idx_off[owner], pu);
        if (lr < 0) atomicOr(errflag, 4);
        tau_dest[u] = F.off + (long long)lr * F.ld + F.f;
    }
    for (long long q = row_ptr[u] + lane; q < row_ptr[u + 1]; q += 32) {
        const int pv = pos[dst[korig[q]]];
        const int owner = pos_front[min(pu, pv)];
        const DFront F = fronts[owner];
        const int *idx = fidx + idx_off[owner];
        const int lr = slot_local_index(F, ffirst[owner], idx, pu), lc = slot_local_index(F, ffirst[owner], idx, pv);
        if (lr < 0 || lc < 0) atomicOr(errflag, 4);
        dest[q] = F.off + (long long)lr * F.ld + lc;
    }
}

void launch_slot_table(int n, const long long *row_ptr, const int *korig, const long long *dst, const int *pos,
                       const int *pos_front, const DFront *fronts, const int *ffirst, const int *fidx,
                       const long long *idx_off, long long *dest, long long *tau_dest, int *errflag, cudaStream_t s) {
    if (n <= 0) return;
    slot_table_kernel<<<(n + 7) / 8, 256, 0, s>>>(n, row_ptr, korig, dst, pos, pos_front, fronts, ffirst, fidx, idx_off, dest,
                                                 tau_dest, errflag);
}

// dense: one CTA per row of K.  Front 0 holds every intermediate (+ A ∪ B as its update set); entries between two
// kept nodes go to the root front.
__global__ void __launch_bounds__(256)
ingest_dense_kernel(double *pool, long long pool_stride, const double *K, int n, const int *pos, int nI,
                    DFront f0, DFront root, double *tau0, int *errflag) {
    const int u = blockIdx.x;
    const long long b = blockIdx.y;
    const double *row = K + (b * n + u) * (long long)n;
    double *P = pool + b * pool_stride;
    __shared__ double red[8];
    __shared__ double tau_s;
    double s = 0.0;
    for (int v = threadIdx.x; v < n; v += 256) s += row[v];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        const double tau = 1.0 / t;
        if (!(t > 0.0) || isinf(tau)) atomicOr(errflag, 2);
        tau_s = tau;
        tau0[b * n + u] = tau;
        const int pu = pos[u];
        if (pu < nI) P[f0.off + (long long)pu * f0.ld + f0.f] = tau;
        else P[root.off + (long long)(pu - nI) * root.ld + root.f] = tau;
    }
    __syncthreads();
    const double tau = tau_s;
    const int pu = pos[u];
    for (int v = threadIdx.x; v < n; v += 256) {
        const int pv = pos[v];
        const double val = row[v] * tau;
        if (pu < nI || pv < nI) P[f0.off + (long long)pu * f0.ld + pv] = val;
        else P[root.off + (long long)(pu - nI) * root.ld + (pv - nI)] = val;
    }
}

void launch_ingest_dense(double *pool, long long pool_stride, int nbatch, const double *K, int n, const int *pos,
                         int nI, DFront f0, DFront root, double *tau0, int *errflag, cudaStream_t s) {
    dim3 grid(n, nbatch);
    ingest_dense_kernel<<<grid, 256, 0, s>>>(pool, pool_stride, K, n, pos, nI, f0, root, tau0, errflag);
}

// =====================================================================================================
// Whole-network kernel for batches of small complete networks (BASELINE config 5: thousands of 256-node networks of a
// temperature sweep).  ONE CTA eliminates ONE network, row block by row block ("up-looking"): a block of 32 rows is
// read from the rate matrix K straight into shared memory (tau = 1 / row sum, P = k tau: the ingest step is fused),
// receives the updates of every earlier pivot block J,
//     R[:, c] += R[:, X_J] U'_J[X_J, c]        c right of X_J            (FP64 DMMA; U'_J comes back from L2)
// in order -- after link J the block's own columns X_{J+1} are final and become the A operand of link J+1 -- and is
// then either a pivot block itself (row sums -> blocked Gauss-Jordan -> U' = W R, written to the front for the later
// blocks and for the back-substitution) or a block of kept rows (A ∪ B), which is the reduced graph and goes straight
// to the root front.  HBM traffic: K once in, the factor rows once out -- the right-looking multi-kernel path streams
// every front through HBM once per 32 pivots (8 x for a 256-node network).  Same arithmetic (sums of non-negative
// products, GTH pivots) in a different order.
// =====================================================================================================
constexpr int DN_THREADS = 256, DN_WARPS = DN_THREADS / 32;
#ifdef NGT_DN_CLOCKS   // phase split of one CTA (printed by block 0): scripts/build_variant.sh dnclk -DNGT_DN_CLOCKS
#define DN_T(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); dn_acc[i] += t_ - dn_last; dn_last = t_; } } while (0)
#else
#define DN_T(i) do { } while (0)
#endif

__host__ __device__ inline int dense_net_ldr(int n) {      // row stride of the row block: >= n + 8 and = 4 mod 16
    int ld = n + 8;
    return ld + ((4 - (ld & 15)) & 15);
}
size_t dense_network_smem(int n) {
    return ((size_t)NB_INNER * dense_net_ldr(n) + NB_INNER * GJ_LD + 8 * WB_LD + NB_INNER) * sizeof(double) + (size_t)n * sizeof(int);
}

__global__ void __launch_bounds__(DN_THREADS, 2)
dense_network_kernel(const double *K, int n, const int *node_at, int k, double *pool, long long pool_stride, DFront F0,
                     DFront RF, double *tau0, int *errflag) {
    extern __shared__ double smem[];
    const int LDR = dense_net_ldr(n);
    double *R = smem;                                 // NB_INNER x LDR: the current row block (column n = tau)
    double *Xs = R + NB_INNER * LDR;                  // NB_INNER x GJ_LD: T | s  ->  W
    double *Wb8 = Xs + NB_INNER * GJ_LD;
    double *ssum = Wb8 + 8 * WB_LD;                   // NB_INNER row sums
    int *nat = reinterpret_cast<int *>(ssum + NB_INNER);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const long long b = blockIdx.x;
    const double *Kb = K + b * (long long)n * n;
    double *M = pool + b * pool_stride + F0.off;      // front 0: pivots 0..k-1, then the kept nodes
    double *root = pool + b * pool_stride + RF.off;
    const long long ld = F0.ld;
    const int f = n;                                  // order of front 0; column f carries tau
    for (int i = threadIdx.x; i < n; i += DN_THREADS) nat[i] = node_at[i];
    __syncthreads();
#ifdef NGT_DN_CLOCKS
    long long dn_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dn_last = clock64();
#endif

    const int npiv_blocks = (k + NB_INNER - 1) / NB_INNER;
    const int nupd_blocks = (f - k + NB_INNER - 1) / NB_INNER;
    for (int blk = 0; blk < npiv_blocks + nupd_blocks; ++blk) {
        const bool is_pivot = blk < npiv_blocks;
        const int p0 = is_pivot ? blk * NB_INNER : k + (blk - npiv_blocks) * NB_INNER;     // first elimination position
        const int cnt = min(NB_INNER, (is_pivot ? k : f) - p0);
        // ---- ingest: rows of K in elimination order, columns permuted the same way; tau = 1 / row sum --------------
        // a warp owns rows warp, warp + 8, ...; the loads of all its rows are issued together (one round of memory
        // latency per 256 columns instead of one per element)
        {
            constexpr int RPW = NB_INNER / DN_WARPS;
            static_assert(RPW == 4, "row sums are selected by hand below");
            double s[RPW];
            const double *kr[RPW];
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int r = warp + i * DN_WARPS;
                s[i] = 0.0;
                kr[i] = r < cnt ? Kb + (long long)nat[p0 + r] * n : nullptr;
            }
            for (int c0 = 0; c0 < n; c0 += 256) {
                double v[RPW][8];
                int src[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) { const int c = c0 + lane + 32 * jj; src[jj] = c < n ? nat[c] : -1; }
#pragma unroll
                for (int i = 0; i < RPW; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) v[i][jj] = (kr[i] && src[jj] >= 0) ? kr[i][src[jj]] : 0.0;
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    double *row = R + (warp + i * DN_WARPS) * LDR;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int c = c0 + lane + 32 * jj;
                        if (c < n) { row[c] = v[i][jj]; s[i] += v[i][jj]; }
                    }
                }
            }
#pragma unroll 1
            for (int i = 0; i < RPW; ++i) {
                const int r = warp + i * DN_WARPS;
                double *row = R + r * LDR;
                const double sum = warp_sum(i == 0 ? s[0] : i == 1 ? s[1] : i == 2 ? s[2] : s[3]);
                if (r < cnt) {
                    const double tau = 1.0 / sum;
                    for (int c = lane; c < n; c += 32) row[c] *= tau;
                    for (int c = n + 1 + lane; c < LDR; c += 32) row[c] = 0.0;
                    if (lane == 0) {
                        row[n] = tau;
                        tau0[b * n + nat[p0 + r]] = tau;
                        if (!(sum > 0.0) || isinf(tau)) atomicOr(errflag, 2);
                    }
                } else {
                    for (int c = n + lane; c < LDR; c += 32) row[c] = 0.0;     // columns < n were written as zeros above
                }
            }
        }
        __syncthreads();
        DN_T(0);
        // ---- links: the updates of the earlier pivot blocks, in order -------------------------------------------------
        const int nlinks = is_pivot ? blk : npiv_blocks;
        for (int J = 0; J < nlinks; ++J) {
            const int jJ = J * NB_INNER, nbJ = min(NB_INNER, k - jJ), jeJ = jJ + nbJ;
            const double *UJ = M + (long long)jJ * ld;            // U'_J rows (written earlier by this CTA; read through L2)
            // B fragments (U'_J) come from L2: the next tile's are requested before the current tile is computed
            auto load_b = [&](int tile, double (&bf)[NB_INNER / 4]) {
                const int colb = tile * 8 + g;
#pragma unroll
                for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
                    const int kk = k4 * 4 + t4;
                    bf[k4] = (kk < nbJ && colb >= jeJ && colb <= f) ? __ldcg(UJ + (long long)kk * ld + colb) : 0.0;
                }
            };
            double bf[NB_INNER / 4], bn[NB_INNER / 4];
            int tile = jeJ / 8 + warp;
            if (tile <= f / 8) load_b(tile, bf);
            for (; tile <= f / 8; tile += DN_WARPS) {
                const bool more = tile + DN_WARPS <= f / 8;
                if (more) load_b(tile + DN_WARPS, bn);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    double *row = R + (mi * 8 + g) * LDR;
                    double c0 = row[tile * 8 + 2 * t4], c1 = row[tile * 8 + 2 * t4 + 1];
#pragma unroll
                    for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
                        const int kk = k4 * 4 + t4;
                        const double a = kk < nbJ ? row[jJ + kk] : 0.0;
                        dmma884(c0, c1, a, bf[k4]);
                    }
                    row[tile * 8 + 2 * t4] = c0;
                    row[tile * 8 + 2 * t4 + 1] = c1;
                }
                if (more) {
#pragma unroll
                    for (int k4 = 0; k4 < NB_INNER / 4; ++k4) bf[k4] = bn[k4];
                }
            }
            __syncthreads();                                      // the next link's A operand is final now
        }
        DN_T(1);
        if (is_pivot) {
            const int j = p0, nbp = cnt, je = j + nbp;
            // ---- mass leaving the panel (GTH row sums) and the T block -------------------------------------------------
            for (int r = warp; r < NB_INNER; r += DN_WARPS) {
                double sacc = 0.0;
                if (r < nbp)
                    for (int c = je + lane; c < f; c += 32) sacc += R[r * LDR + c];
                sacc = warp_sum(sacc);
                if (lane == 0) ssum[r] = (r < nbp) ? sacc : 1.0;   // padding rows behave like isolated absorbing nodes
            }
            __syncthreads();
            for (int i = threadIdx.x; i < NB_INNER * GJ_LD; i += DN_THREADS) {
                const int r = i / GJ_LD, q = i % GJ_LD;
                double v = 0.0;
                if (q < NB_INNER) { if (r < nbp && q < nbp && q != r) v = R[r * LDR + j + q]; }
                else if (q == NB_INNER) v = ssum[r];
                Xs[i] = v;
            }
            __syncthreads();
            DN_T(2);
            if (warp < GJ_WARPS) gj_blocked(Xs, Wb8, nbp, errflag, warp, lane);
            __syncthreads();
            DN_T(3);
            // ---- U' = W [P_XR | tau_X]: kept in the front for the later blocks and the back-substitution ----------------
            for (int tile = je / 8 + warp; tile <= f / 8; tile += DN_WARPS) {
                const int colb = tile * 8 + g;
                double bf[NB_INNER / 4];
#pragma unroll
                for (int k4 = 0; k4 < NB_INNER / 4; ++k4) bf[k4] = R[(k4 * 4 + t4) * LDR + colb];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int k4 = 0; k4 < NB_INNER / 4; ++k4) dmma884(c0, c1, Xs[(mi * 8 + g) * GJ_LD + k4 * 4 + t4], bf[k4]);
                    const int r = mi * 8 + g, cc = tile * 8 + 2 * t4;
                    if (r < nbp) {
                        double *out = M + (long long)(j + r) * ld;
                        if (cc >= je && cc <= f) out[cc] = c0;
                        if (cc + 1 >= je && cc + 1 <= f) out[cc + 1] = c1;
                    }
                }
            }
        } else {
            // ---- kept rows: the reduced graph on A ∪ B, with the full tau' in the last column ---------------------------
            const int mcols = f - k + 1;
            for (int i = threadIdx.x; i < cnt * mcols; i += DN_THREADS) {
                const int r = i / mcols, c = i % mcols;
                root[(long long)(p0 - k + r) * RF.ld + c] = R[r * LDR + k + c];
            }
        }
        __syncthreads();                                          // R is reloaded by the next block
        DN_T(4);
    }
#ifdef NGT_DN_CLOCKS
    if (threadIdx.x == 0 && blockIdx.x == 0)
        printf("dense_network_kernel clocks (CTA 0): ingest %lld  links %lld  rowsum+T %lld  gauss-jordan %lld  U'/root %lld\n",
               dn_acc[0], dn_acc[1], dn_acc[2], dn_acc[3], dn_acc[4]);
#endif
}

void launch_dense_network(const double *K, int n, int nbatch, const int *node_at, int k, double *pool, long long pool_stride,
                          DFront F0, DFront RF, double *tau0, int *errflag, cudaStream_t s) {
    dense_network_kernel<<<nbatch, DN_THREADS, dense_network_smem(n), s>>>(K, n, node_at, k, pool, pool_stride, F0, RF, tau0,
                                                                          errflag);
}

// =====================================================================================================
// assembly (extend-add), one warp per row of a parent front.  The host lists, per parent row, the (child, child row)
// pairs that contribute to it, in the fixed order of the children; the warp walks each contributing child row with
// unit stride and scatters through the child's relative map (child update index -> parent local index):
//   parent[a, rel_c[j]] += child[k_c + i, k_c + j],   parent[a, tau] += child[k_c + i, tau]     for every (c, i) of row a
// Every parent row has exactly one writer (its warp), so there are no atomics and the order of the sums is fixed.
// Work is proportional to the children's update blocks, not to (rows touched) x (parent order).
// grid (row tiles of 8, parents of the level, networks).
// =====================================================================================================
// WPR warps share a parent row, each owning a contiguous range of PARENT columns (the relative maps are ascending, so
// a warp finds its part of a child row by two binary searches): still one writer per entry, and levels with only a
// few large parents (the top of the tree) get WPR times the warps in flight.
template <int WPR>
__global__ void __launch_bounds__(256)
assemble_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, const long long *row_base,
                const long long *row_ptr, const int *row_child, const int *row_crow, const int *__restrict__ rel,
                const long long *rel_off, const int *__restrict__ child_split) {
    pdl_enter();
    const int pid = ids[blockIdx.y];
    const DFront P = fronts[pid];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = blockIdx.x * (8 / WPR) + warp / WPR;
    if (a >= P.f) return;
    const int part = warp % WPR;                          // this warp's share of the parent's columns (see child_split)
    double *base = pool + (long long)blockIdx.z * pool_stride;
    double *prow = base + P.off + (long long)a * P.ld;
    const long long r = row_base[pid] + a;
    for (long long e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
        const int cid = row_child[e];
        const DFront C = fronts[cid];
        const int m = C.f - C.k;
        const double *crow = base + C.off + (long long)(C.k + row_crow[e]) * C.ld + C.k;
        const int *rl = rel + rel_off[cid];
        int j0 = 0, j1 = m;
        if (WPR > 1) {
            // first child column mapped at or beyond each quarter of the parent's columns: the same for every row of
            // this child, so the host tabulates it (a pair of binary searches here cost ~20 dependent loads per
            // contribution, most of this kernel's time at the top of the tree)
            const int *sp = child_split + (long long)cid * (WPR - 1);
            j0 = part == 0 ? 0 : sp[part - 1];
            j1 = part == WPR - 1 ? m : sp[part];
        }
        int j = j0 + lane;
        for (; j + 96 < j1; j += 128) {                   // four gathers in flight per lane
            const double v0 = crow[j], v1 = crow[j + 32], v2 = crow[j + 64], v3 = crow[j + 96];
            const int i0 = rl[j], i1 = rl[j + 32], i2 = rl[j + 64], i3 = rl[j + 96];
            const double p0 = prow[i0], p1 = prow[i1], p2 = prow[i2], p3 = prow[i3];
            if (v0 != 0.0) prow[i0] = p0 + v0;
            if (v1 != 0.0) prow[i1] = p1 + v1;
            if (v2 != 0.0) prow[i2] = p2 + v2;
            if (v3 != 0.0) prow[i3] = p3 + v3;
        }
        for (; j < j1; j += 32) {
            const double v = crow[j];
            if (v != 0.0) prow[rl[j]] += v;
        }
        if (part == 0 && lane == 0) {
            const double v = crow[m];              // the child's delta-tau column
            if (v != 0.0) prow[P.f] += v;
        }
        __syncwarp();                              // the next child may touch the same entries from other lanes
    }
}

bool assemble_splits_rows(int nparents, int nbatch, int max_f) {
    // only levels with a handful of parents: with tens of parents the extra warps per row cost more than they hide
    // (measured on the 10^5-minima landscape: 0.60 -> 0.69 ms when rows of >= 384 columns were split as well)
    return (long long)nparents * nbatch * ((max_f + 7) / 8) < 4 * 148;
}

void launch_assemble(const LaunchCtx &c, const int *ids, int nparents, int max_f, const long long *row_base,
                     const long long *row_ptr, const int *row_child, const int *row_crow, const int *rel,
                     const long long *rel_off, const int *child_split) {
    if (nparents <= 0 || max_f <= 0) return;
    // few, large parents: four warps per row
    if (child_split && assemble_splits_rows(nparents, c.nbatch, max_f)) {
        dim3 grid((max_f + 1) / 2, nparents, c.nbatch);
        launch_chain(assemble_kernel<ASM_WPR>, grid, dim3(256), 0, c.stream, c.pool, c.pool_stride, c.fronts, ids, row_base,
                     row_ptr, row_child, row_crow, rel, rel_off, child_split);
    } else {
        dim3 grid((max_f + 7) / 8, nparents, c.nbatch);
        launch_chain(assemble_kernel<1>, grid, dim3(256), 0, c.stream, c.pool, c.pool_stride, c.fronts, ids, row_base, row_ptr,
                     row_child, row_crow, rel, rel_off, child_split);
    }
}

// =====================================================================================================
// phase 2 (reference ngt.hpp:362-431): for every a of a group, 1-P'_aa and tau'_a after the rest of the group is
// eliminated.  The reference repeats a full elimination per a (O(|A|^4) in all).  Eliminations commute, so the
// per-a results are the leaves of a binary tree of eliminations instead: a tree node holds the chain reduced to
// (kept set S) + (the other group collapsed into one absorbing column); its two children eliminate one half of S each
// and keep the other.  Every a is still obtained by eliminating group \ {a} -- only the order differs from the
// reference's and the shared prefixes are computed once: 4/3 |A|^3 flops per group, log2|A| levels of fronts.
//
// A phase-2 front (order npk + 1): pivots = parent's kept members that this task drops, then the kept members T, then
// the collapsed column; column f carries tau.  After its elimination rows/cols k.. are the chain on T: the source of
// its two children (or, when |T| = 1, the result: 1-P'_aa = M[k][k+1], tau'_a = M[k][f]).
// =====================================================================================================
__device__ __forceinline__ int p2_parent_index(const P2Task &t, int l) {
    const int k = t.npk - t.t_len;                 // dropped members come first
    if (l < k) return t.t_off == 0 ? t.t_len + l : l;
    return t.t_off + (l - k);
}

__global__ void __launch_bounds__(256)
phase2_gather_kernel(const double *pool, long long pool_stride, DFront root, double *p2pool, long long p2_stride,
                     const DFront *p2fronts, const P2Task *tasks, int task0) {
    const int tid = task0 + blockIdx.y;
    const P2Task t = tasks[tid];
    const DFront D = p2fronts[tid];
    const int lr = blockIdx.x;                     // row of the new front
    const int fD = t.npk + 1;                      // members + the collapsed column
    if (lr >= fD) return;
    double *M = p2pool + (long long)blockIdx.z * p2_stride + D.off;
    double *dst = M + (long long)lr * D.ld;
    if (lr == fD - 1) {                            // the collapsed node is never a pivot: a zero row
        for (int lc = threadIdx.x; lc <= fD; lc += 256) dst[lc] = 0.0;
        return;
    }
    const int pi = p2_parent_index(t, lr);
    const double *src;
    int col0, ccol, tcol;                          // first member column, collapsed column, tau column of the source
    if (t.src < 0) {
        src = pool + (long long)blockIdx.z * pool_stride + root.off + (long long)(t.src_k + pi) * root.ld;
        col0 = t.src_k; ccol = -1; tcol = root.f;
    } else {
        const DFront S = p2fronts[t.src];
        src = p2pool + (long long)blockIdx.z * p2_stride + S.off + (long long)(t.src_k + pi) * S.ld;
        col0 = t.src_k; ccol = t.src_k + t.npk; tcol = S.f;
    }
    for (int lc = threadIdx.x; lc < fD - 1; lc += 256)
        dst[lc] = (lc == lr) ? 0.0 : src[col0 + p2_parent_index(t, lc)];
    if (threadIdx.x < 32) {
        double s = 0.0;
        if (ccol < 0) {                            // root source: mass into the other group, fixed summation order
            for (int c = threadIdx.x; c < t.no; c += 32) s += src[t.o0 + c];
            s = warp_sum(s);
        } else {
            s = src[ccol];
        }
        if (threadIdx.x == 0) {
            dst[fD - 1] = s;
            dst[fD] = src[tcol];
        }
    }
}

void launch_phase2_gather(const LaunchCtx &c, DFront root, double *p2pool, long long p2_stride, const DFront *p2fronts,
                          const P2Task *tasks, int task0, int ntask, int max_f) {
    if (ntask <= 0) return;
    dim3 grid(max_f, ntask, c.nbatch);
    phase2_gather_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, root, p2pool, p2_stride, p2fronts, tasks, task0);
}

// leaves of the tree: front `leaf_front[i]` kept exactly one member, whose result goes to slot leaf_out[i]
__global__ void phase2_collect_kernel(const double *p2pool, long long p2_stride, const DFront *p2fronts,
                                      const int *leaf_front, const int *leaf_out, int nleaf, double *om, double *tau,
                                      int out_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nleaf) return;
    const long long b = blockIdx.y;
    const DFront T = p2fronts[leaf_front[i]];
    const double *M = p2pool + b * p2_stride + T.off;
    om[b * out_stride + leaf_out[i]] = M[(long long)T.k * T.ld + T.k + 1];
    tau[b * out_stride + leaf_out[i]] = M[(long long)T.k * T.ld + T.f];
}

void launch_phase2_collect(const LaunchCtx &c, const double *p2pool, long long p2_stride, const DFront *p2fronts,
                           const int *leaf_front, const int *leaf_out, int nleaf, double *om, double *tau,
                           int out_stride) {
    if (nleaf <= 0) return;
    dim3 grid((nleaf + 127) / 128, c.nbatch);
    phase2_collect_kernel<<<grid, 128, 0, c.stream>>>(p2pool, p2_stride, p2fronts, leaf_front, leaf_out, nleaf, om, tau,
                                                     out_stride);
}

// =====================================================================================================
// multi-GPU dense path (1-D block-cyclic columns): helper kernels.  Every rank holds the whole
// matrix but only keeps its own column blocks (width NB_OUTER) up to date; the outer block being eliminated travels
// as a rectangular "panel front" (rows J.., its NB_OUTER columns, one column S with the mass outside the block, tau).
// =====================================================================================================
// partial S_r = sum over OWNED columns c in [JE, f) of P[J + r, c]; one CTA per row
__global__ void __launch_bounds__(256)
dist_rowsum_kernel(const double *M, int ld, int f, int J, int JE, KOpt ko, double *out) {
    const int r = blockIdx.x;
    const double *row = M + (long long)(J + r) * ld;
    double s = 0.0;
    for (int c = JE + threadIdx.x; c < f; c += 256)
        if (owned_col(ko, c, f)) s += row[c];
    s = warp_sum(s);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) out[r] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// owner: panel front <- column panel of the big matrix + S + tau
__global__ void __launch_bounds__(128)
dist_pack_kernel(const double *M, int ld, int J, int kb, int fr, double *pan, int ld_pan, const double *S,
                 const double *tau_glob) {
    const int r = blockIdx.x;
    if (r >= fr) return;
    const double *row = M + (long long)(J + r) * ld + J;
    double *dst = pan + (long long)r * ld_pan;
    for (int c = threadIdx.x; c < ld_pan; c += 128) {
        double v = 0.0;
        if (c < kb) v = row[c];
        else if (c == kb) v = (r < kb) ? S[r] : 0.0;
        else if (c == kb + 1) v = tau_glob[J + r];
        dst[c] = v;
    }
}

// everybody, after the broadcast: the panel's tau column is the new tau of rows J..
__global__ void dist_tau_kernel(const double *pan, int ld_pan, int kb, int fr, int J, double *tau_glob) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < fr) tau_glob[J + r] = pan[(long long)r * ld_pan + kb + 1];
}

__global__ void dist_tau_init_kernel(const double *tau0, const int *pos, int n, double *tau_glob) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) tau_glob[pos[u]] = tau0[u];
}

// contributions of this rank's columns to the reduced A ∪ B block (others zero; summed with one all-reduce)
__global__ void dist_root_contrib_kernel(const double *M, int ld, int f, int nI, int m, KOpt ko, double *out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * m) return;
    const int i = idx / m, jx = idx % m;
    out[idx] = owned_col(ko, nI + jx, f) ? M[(long long)(nI + i) * ld + nI + jx] : 0.0;
}

__global__ void dist_root_finish_kernel(double *root, int ld_root, int m, const double *contrib, const double *tau_glob,
                                        int nI) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * (m + 1)) return;
    const int i = idx / (m + 1), jx = idx % (m + 1);
    if (jx < m) root[(long long)i * ld_root + jx] += contrib[i * m + jx];
    else root[(long long)i * ld_root + m] = tau_glob[nI + i];
}

// Every rank, after the panel broadcast: the block's row operations replayed on the own columns right of the block,
// all eight inner panels in ONE launch (the per-panel rowpanel + row-strip update launches cost 16 dependent launches
// per outer block on the critical chain of the multi-GPU loop).  A CTA owns a slab of 64 columns of the block's kb
// rows in shared memory; inside it a warp owns 8 columns for the whole replay, so the only synchronisation is the
// staging of the panel's sub-diagonal blocks (the A operands, shared by all warps):
//   for every inner panel X_j:   S[X_j, :]  = W_j S[X_j, :]                       (U'_j for these columns)
//                                S[i, :]   += P[i, X_j] S[X_j, :]   rows i below X_j inside the block
// with W_j from wbuf and P[i, X_j] from the broadcast panel (both FP64 DMMA).
constexpr int SLAB_W = 64, SLAB_LD = SLAB_W + 4, SLAB_A_LD = NB_INNER + 4;
size_t dist_apply_slab_smem() { return ((size_t)NB_OUTER * SLAB_LD + (size_t)NB_OUTER * SLAB_A_LD) * sizeof(double); }

__global__ void __launch_bounds__(256, 1)
dist_apply_slab_kernel(double *M, long long ld, int f, int J, int kb, const double *pan, int ld_pan, const double *W, KOpt ko) {
    extern __shared__ double smem[];
    double *S = smem;                              // kb x SLAB_LD: rows J.. of the slab's columns
    double *As = smem + NB_OUTER * SLAB_LD;        // rows below the current inner panel x its 32 columns
    const int JE = J + kb;
    const int c0 = JE + blockIdx.x * SLAB_W;
    if (c0 >= f) return;
    // a slab lies inside one 256-column ownership block unless the block is the last, short one: test both ends
    if (ko.own_G > 1 && !owned_col(ko, c0, f) && !owned_col(ko, min(c0 + SLAB_W, f) - 1, f)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    // staging goes through cp.async (fire and forget, zero fill for masked slots): a plain load -> store loop stalls on
    // every element with the 8 warps of this CTA (measured 76 us per launch instead of ~20)
    // (16-byte copies when the slab starts on an even column and f is even -- always, except for a short last block:
    // 8-byte LDGSTS instructions were the bottleneck of the first version)
    if (((c0 | f) & 1) == 0) {
        const int cc = 2 * (threadIdx.x % (SLAB_W / 2)), c = c0 + cc;
        const bool col_ok = c < f && owned_col(ko, c, f);          // c even, f even: c + 1 < f and the same owner
        for (int r = threadIdx.x / (SLAB_W / 2); r < kb; r += 256 / (SLAB_W / 2))
            cp_async16(S + r * SLAB_LD + cc, col_ok ? M + (long long)(J + r) * ld + c : M, col_ok);
    } else {
        const int cc = threadIdx.x % SLAB_W, c = c0 + cc;
        const bool col_ok = c < f && owned_col(ko, c, f);
        for (int r = threadIdx.x / SLAB_W; r < kb; r += 256 / SLAB_W)
            cp_async8(S + r * SLAB_LD + cc, col_ok ? M + (long long)(J + r) * ld + c : M, col_ok);
    }
    cp_async_commit();
    const int wc = warp * 8;                       // this warp's 8 columns of the slab
    for (int j = 0; j < kb; j += NB_INNER) {
        const int nbp = min(NB_INNER, kb - j), je = j + nbp;
        const int nbelow = kb - je;
        __syncthreads();                           // previous panel's A block no longer needed
        if ((nbp & 1) == 0) {
            const int q = 2 * (threadIdx.x % (NB_INNER / 2));
            for (int r = threadIdx.x / (NB_INNER / 2); r < nbelow; r += 256 / (NB_INNER / 2))
                cp_async16(As + r * SLAB_A_LD + q, q < nbp ? pan + (long long)(je + r) * ld_pan + j + q : pan, q < nbp);
        } else {
            const int q = threadIdx.x % NB_INNER;
            for (int r = threadIdx.x / NB_INNER; r < nbelow; r += 256 / NB_INNER)
                cp_async8(As + r * SLAB_A_LD + q, q < nbp ? pan + (long long)(je + r) * ld_pan + j + q : pan, q < nbp);
        }
        cp_async_commit();
        cp_async_wait<0>();                        // (also the slab itself, on the first panel)
        __syncthreads();
        // U'_j = W_j S[X_j, own columns]
        const double *Wj = W + (long long)(j / NB_INNER) * NB_INNER * NB_INNER;
        double bf[NB_INNER / 4];
#pragma unroll
        for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
            const int kk = k4 * 4 + t4;
            bf[k4] = kk < nbp ? S[(j + kk) * SLAB_LD + wc + g] : 0.0;
        }
        double u[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            u[mi][0] = u[mi][1] = 0.0;
#pragma unroll
            for (int k4 = 0; k4 < NB_INNER / 4; ++k4) dmma884(u[mi][0], u[mi][1], __ldg(Wj + (mi * 8 + g) * NB_INNER + k4 * 4 + t4), bf[k4]);
        }
        __syncwarp();                              // every lane has read the old rows of this warp's columns
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int r = mi * 8 + g;
            if (r < nbp) { S[(j + r) * SLAB_LD + wc + 2 * t4] = u[mi][0]; S[(j + r) * SLAB_LD + wc + 2 * t4 + 1] = u[mi][1]; }
        }
        __syncwarp();
        // rows below X_j inside the block: S[i, :] += P[i, X_j] U'_j
#pragma unroll
        for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
            const int kk = k4 * 4 + t4;
            bf[k4] = kk < nbp ? S[(j + kk) * SLAB_LD + wc + g] : 0.0;
        }
        // four row tiles at a time: four independent accumulator chains (a single chain of 8 dependent DMMAs per tile
        // left the tensor pipe idle most of the time)
        for (int r0 = 0; r0 < nbelow; r0 += 32) {
            double cv[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 8 + g;      // row below the panel (As index); rows >= nbelow: masked
                const double *srow = S + (je + r) * SLAB_LD + wc + 2 * t4;
                cv[u][0] = r < nbelow ? srow[0] : 0.0;
                cv[u][1] = r < nbelow ? srow[1] : 0.0;
            }
#pragma unroll
            for (int k4 = 0; k4 < NB_INNER / 4; ++k4)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u * 8 + g;
                    dmma884(cv[u][0], cv[u][1], r < nbelow ? As[r * SLAB_A_LD + k4 * 4 + t4] : 0.0, bf[k4]);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 8 + g;
                if (r < nbelow) {
                    double *srow = S + (je + r) * SLAB_LD + wc + 2 * t4;
                    srow[0] = cv[u][0];
                    srow[1] = cv[u][1];
                }
            }
        }
    }
    __syncthreads();
    {
        const int cc = threadIdx.x % SLAB_W, c = c0 + cc;
        if (c < f && owned_col(ko, c, f))
            for (int r = threadIdx.x / SLAB_W; r < kb; r += 256 / SLAB_W) M[(long long)(J + r) * ld + c] = S[r * SLAB_LD + cc];
    }
}

void launch_dist_apply_slab(double *M, long long ld, int f, int J, int kb, const double *pan, int ld_pan, const double *W,
                            KOpt ko, cudaStream_t s) {
    const int ncols = f - (J + kb);
    if (ncols <= 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(dist_apply_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dist_apply_slab_smem());
        attr_set = true;
    }
    dist_apply_slab_kernel<<<(ncols + SLAB_W - 1) / SLAB_W, 256, dist_apply_slab_smem(), s>>>(M, ld, f, J, kb, pan, ld_pan, W, ko);
}

// Owner, after the SQUARE factorisation of the block (its kb rows only): the rows below the block, restricted to the
// block's columns and the tau column, take the block's column operations in ONE launch -- the right-looking
// factorisation of the full rectangular panel spent 8 thin (K = 32) update launches over up to 16 384 rows on this.
// A CTA owns 64 rows of the panel in shared memory, a warp 8 of them for the whole replay:
//   for every inner panel X_j:  P[i, X_j] is final (the multiplier that travels with the panel);
//                               P[i, c] += P[i, X_j] U'_j[X_j, c]   for the block columns c right of X_j and for tau
// U'_j comes from the factored block rows of the same panel buffer (staged per inner panel, shared by the warps).
constexpr int CS_ROWS = 64;
size_t dist_colstrip_smem(int ld_pan) { return (size_t)(CS_ROWS + NB_INNER) * ld_pan * sizeof(double); }

// Shared-memory tiles keep the panel's own row stride (ld_pan), so that a tile is ONE contiguous piece of the panel
// buffer and is staged by ONE bulk copy of the copy engine (cp.async.bulk onto an mbarrier; written back the same
// way).  Measured on the way here: element-wise cp.async staging cost 40 % of the kernel in LDGSTS issue alone, one bulk
// copy per row ~60 cycles each in the copy engine (32 rows x 8 panels), one copy per tile is free.  The stride
// (264 = 8 mod 16 doubles) makes the fragment loads 2-way bank conflicted; the kernel is bound by the DMMA pipe.
__global__ void __launch_bounds__(256, 1)
dist_colstrip_slab_kernel(double *pan, int ld_pan, int kb, int fr) {
    extern __shared__ __align__(16) double smem[];
    const int LD = ld_pan, ncol = kb + 2;                // block columns, the S column (carried, unused), tau
    double *S = smem;                                    // CS_ROWS x LD
    double *Bs = smem + CS_ROWS * LD;                    // NB_INNER x LD: rows X_j of the factored block (U'_j right of X_j)
    const int r0 = kb + blockIdx.x * CS_ROWS;
    if (r0 >= fr) return;
    const int nrows = min(CS_ROWS, fr - r0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    __shared__ unsigned long long bar;
    unsigned phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = nrows * LD + threadIdx.x; i < CS_ROWS * LD; i += 256) S[i] = 0.0;      // rows past the panel's end
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)((size_t)nrows * LD * sizeof(double));
        mbar_expect_tx(&bar, bytes);
        bulk_g2s(S, pan + (long long)r0 * ld_pan, bytes, &bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    double *srow = S + (warp * 8 + g) * LD;              // this lane's row of the warp's 8-row tile
#ifdef NGT_DN_CLOCKS
    long long dn_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dn_last = clock64();
#endif
    for (int j = 0; j < kb; j += NB_INNER) {
        const int nbp = min(NB_INNER, kb - j), je = j + nbp;
        __syncthreads();                                 // every warp is done reading the previous U'_j
        DN_T(0);
        // (no proxy fence: the copy overwrites what the warps only READ in the previous step, and the block barrier above
        // orders those reads before this issue)
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)((size_t)nbp * LD * sizeof(double));
            mbar_expect_tx(&bar, bytes);
            bulk_g2s(Bs, pan + (long long)j * ld_pan, bytes, &bar);
        }
        DN_T(1);
        mbar_wait(&bar, phase);
        phase ^= 1;
        DN_T(2);
        double a[NB_INNER / 4];
#pragma unroll
        for (int k4 = 0; k4 < NB_INNER / 4; ++k4) {
            const int kk = k4 * 4 + t4;
            a[k4] = kk < nbp ? srow[j + kk] : 0.0;       // (rows of Bs past nbp are stale: their multipliers are zero)
        }
        for (int ct0 = je / 8; ct0 * 8 < ncol; ct0 += 4) {      // four column tiles = four independent accumulator chains
            double cv[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = (ct0 + u) * 8 + 2 * t4;
                const bool ok = (ct0 + u) * 8 < ncol;
                cv[u][0] = ok ? srow[c] : 0.0;
                cv[u][1] = ok ? srow[c + 1] : 0.0;
            }
#pragma unroll
            for (int k4 = 0; k4 < NB_INNER / 4; ++k4)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int cb = (ct0 + u) * 8 + g;    // columns left of je (only in a short last panel) must not contribute
                    const bool ok = (ct0 + u) * 8 < ncol && cb >= je && k4 * 4 + t4 < nbp;
                    dmma884(cv[u][0], cv[u][1], a[k4], ok ? Bs[(k4 * 4 + t4) * LD + cb] : 0.0);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((ct0 + u) * 8 < ncol) {
                    const int c = (ct0 + u) * 8 + 2 * t4;
                    srow[c] = cv[u][0];
                    srow[c + 1] = cv[u][1];
                }
        }
        __syncwarp();                                    // the tile's columns X_{j+1} are read by other lanes next
        DN_T(3);
    }
    fence_proxy_async_smem();                            // the rows go back through the copy engine (generic writes -> async reads)
    __syncthreads();
    DN_T(4);
    if (threadIdx.x == 0) {
        bulk_s2g(pan + (long long)r0 * ld_pan, S, (unsigned)((size_t)nrows * LD * sizeof(double)));
        bulk_commit_wait_all();
    }
    DN_T(5);
#ifdef NGT_DN_CLOCKS
    if (threadIdx.x == 0 && blockIdx.x == 0 && kb == NB_OUTER)
        printf("colstrip slab clocks (CTA 0, %d CTAs): barrier %lld  issue staging %lld  wait staging %lld  dmma %lld  final barrier %lld  write back %lld\n",
               (int)gridDim.x, dn_acc[0], dn_acc[1], dn_acc[2], dn_acc[3], dn_acc[4], dn_acc[5]);
#endif
}

void launch_dist_colstrip_slab(double *pan, int ld_pan, int kb, int fr, cudaStream_t s) {
    if (fr <= kb) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(dist_colstrip_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dist_colstrip_smem(ld_pan));
        attr_set = true;
    }
    dist_colstrip_slab_kernel<<<(fr - kb + CS_ROWS - 1) / CS_ROWS, 256, dist_colstrip_smem(ld_pan), s>>>(pan, ld_pan, kb, fr);
}

void launch_dist_rowsum(const double *M, int ld, int f, int J, int JE, KOpt ko, double *out, cudaStream_t s) {
    dist_rowsum_kernel<<<JE - J, 256, 0, s>>>(M, ld, f, J, JE, ko, out);
}
void launch_dist_pack(const double *M, int ld, int J, int kb, int fr, double *pan, int ld_pan, const double *S,
                      const double *tau_glob, cudaStream_t s) {
    dist_pack_kernel<<<fr, 128, 0, s>>>(M, ld, J, kb, fr, pan, ld_pan, S, tau_glob);
}
void launch_dist_tau(const double *pan, int ld_pan, int kb, int fr, int J, double *tau_glob, cudaStream_t s) {
    dist_tau_kernel<<<(fr + 255) / 256, 256, 0, s>>>(pan, ld_pan, kb, fr, J, tau_glob);
}
void launch_dist_tau_init(const double *tau0, const int *pos, int n, double *tau_glob, cudaStream_t s) {
    dist_tau_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(tau0, pos, n, tau_glob);
}
void launch_dist_root_contrib(const double *M, int ld, int f, int nI, int m, KOpt ko, double *out, cudaStream_t s) {
    dist_root_contrib_kernel<<<(m * m + 255) / 256, 256, 0, s>>>(M, ld, f, nI, m, ko, out);
}
void launch_dist_root_finish(double *root, int ld_root, int m, const double *contrib, const double *tau_glob, int nI,
                             cudaStream_t s) {
    dist_root_finish_kernel<<<(m * (m + 1) + 255) / 256, 256, 0, s>>>(root, ld_root, m, contrib, tau_glob, nI);
}

// =====================================================================================================
// Graph validation on the device: connected components of the undirected structure of the rate graph
// by a lock-free union-find -- what the reference does with networkx sweeps before it accepts A and B (reference
// kmc_rates/_kmc_rates.py:384-454) and when it drops rates that cannot reach B (rates_linalg.py:13-41).
// Roots are always hooked towards the smaller id, so a component's label is its smallest node id (deterministic).
// =====================================================================================================
__device__ __forceinline__ int cc_find(int *parent, int x) {
    int p = parent[x];
    while (p != x) {                       // path halving; concurrent writers only ever move entries closer to the root
        const int gp = parent[p];
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}

__global__ void cc_init_kernel(int *parent, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = i;
}

__global__ void cc_hook_kernel(int *parent, const long long *src, const long long *dst, long long nnz, int n, int *errflag) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const long long u0 = src[e], v0 = dst[e];
    if (u0 < 0 || u0 >= n || v0 < 0 || v0 >= n) { atomicOr(errflag, 8); return; }
    int u = (int)u0, v = (int)v0;
    while (true) {
        u = cc_find(parent, u);
        v = cc_find(parent, v);
        if (u == v) break;
        const int hi = max(u, v), lo = min(u, v);
        const int old = atomicCAS(parent + hi, hi, lo);      // hook the larger root under the smaller one
        if (old == hi) break;
        u = old;                                             // somebody else moved hi first: continue from there
        v = lo;
    }
}

// every entry is rewritten by its own thread only (a halving write of another thread could otherwise put an old
// ancestor back over a freshly written root); readers see either the old ancestor or the root, both lead to the root
__global__ void cc_flatten_kernel(int *parent, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = i;
    for (int p = parent[r]; p != r; p = parent[r]) r = p;
    parent[i] = r;
}

void launch_connected_components(int *parent, const long long *src, const long long *dst, long long nnz, int n, int *errflag,
                                 cudaStream_t s) {
    if (n <= 0) return;
    cc_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(parent, n);
    if (nnz > 0) cc_hook_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, s>>>(parent, src, dst, nnz, n, errflag);
    cc_flatten_kernel<<<(n + 255) / 256, 256, 0, s>>>(parent, n);
}

// =====================================================================================================
// back-substitution (committors q and mean first-passage times t to A ∪ B for every eliminated node):
//   q_X = U' q_R,  t_X = tau'_X + U' t_R     with U' = rows X of the eliminated front, R = everything later.
// One CTA per (row of the current panel, front, network); panels are swept in reverse by the host.
// x holds (q, t) interleaved per elimination position.
// =====================================================================================================
__global__ void __launch_bounds__(128)
backsub_kernel(const double *pool, long long pool_stride, const DFront *fronts, const int *ids, const int *fidx,
               const long long *idx_off, double *x, long long x_stride, int j) {
    const int fid = ids[blockIdx.y];
    const DFront F = fronts[fid];
    if (j >= F.k) return;
    const int nbp = min(NB_INNER, F.k - j);
    const int r = blockIdx.x;
    if (r >= nbp) return;
    const int je = j + nbp;
    const double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    double *xb = x + (long long)blockIdx.z * x_stride;
    const int *idx = fidx + idx_off[fid];
    const double *row = M + (long long)(j + r) * F.ld;
    double q = 0.0, t = 0.0;
    for (int c = je + threadIdx.x; c < F.f; c += 128) {
        const double u = row[c];
        const long long p = idx[c];
        q += u * xb[2 * p];
        t += u * xb[2 * p + 1];
    }
    __shared__ double rq[4], rt[4];
    q = warp_sum(q);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) { rq[threadIdx.x >> 5] = q; rt[threadIdx.x >> 5] = t; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long p = idx[j + r];
        xb[2 * p] = rq[0] + rq[1] + rq[2] + rq[3];
        xb[2 * p + 1] = rt[0] + rt[1] + rt[2] + rt[3] + row[F.f];
    }
}

void launch_backsub_step(const LaunchCtx &c, const int *ids, int nfb, const int *fidx, const long long *idx_off,
                         double *x, long long x_stride, int j) {
    dim3 grid(NB_INNER, nfb, c.nbatch);
    backsub_kernel<<<grid, 128, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, fidx, idx_off, x, x_stride, j);
}

}  // namespace ngt

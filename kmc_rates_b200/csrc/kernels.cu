// CUDA kernels of the multifrontal NGT engine, written for sm_100a (B200).
// Layout and the blocked elimination scheme are described in kernels.cuh.
#include "kernels.cuh"

#include <cuda.h>

#include <cstdio>

namespace ngt {

// =====================================================================================================
// helpers
// =====================================================================================================
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
              double *rsbuf, int maxfb, int maxchunks) {
    const DFront F = fronts[ids[blockIdx.y]];
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
    rowsum_kernel<<<grid, 1024, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, c.rsbuf, c.maxfb, c.maxchunks);
}

// =====================================================================================================
// panel kernel: one CTA (1024 threads, thread (u,q) owns T[u][q] and G[u][q] in registers) per front and network.
//   s_u  = sum_{v >= je} P[j+u, v]                (mass leaving the panel; GTH row sums)
//   W    = (I - P_XX)^-1 by subtraction-free Gauss-Jordan; pivot d_p = sum_{q != p} T[p,q] + s_p.
// Row p is published to shared memory (double buffered) by its own warp one step ahead, together with 1/d_p,
// so a step costs one barrier, one shuffle and three FMAs for everybody else.
// Fronts with few columns also apply W to their row panel here (U' = W [P_XR | tau_X]).
// =====================================================================================================
__global__ void __launch_bounds__(1024, 1)
panel_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j,
             double *wbuf, const double *rsbuf, int maxfb, int maxchunks, int *errflag, KOpt ko) {
    const DFront F = fronts[ids[blockIdx.x]];
    if (j >= F.k) return;
    double *M = pool + (long long)blockIdx.y * pool_stride + F.off;
    const int nbp = min(NB_INNER, F.k - j);
    const int je = j + nbp;
    const int u = threadIdx.x >> 5, q = threadIdx.x & 31;

    __shared__ double rowT[2][NB_INNER];
    __shared__ double rowG[2][NB_INNER];
    __shared__ double rowS[2], rowRD[2];
    __shared__ double Ds[NB_INNER];
    __shared__ double Ws[NB_INNER][NB_INNER];
    __shared__ double xs[NB_INNER][128];

    // row sums over everything to the right of the panel (tau column excluded)
    double s = 1.0;   // padding rows behave like isolated absorbing nodes
    if (u < nbp) {
        if (F.f - je > RS_SPLIT) {
            const int nch = (F.f - je + RS_CHUNK - 1) / RS_CHUNK;
            const double *rs = rsbuf + (((long long)blockIdx.y * maxfb + blockIdx.x) * maxchunks) * NB_INNER;
            s = 0.0;
            for (int ch = q; ch < nch; ch += 32) s += rs[ch * NB_INNER + u];
            s = warp_sum(s);
        } else {
            const double *row = M + (long long)(j + u) * F.ld;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int c = je + q;
            for (; c + 96 < F.f; c += 128) { s0 += row[c]; s1 += row[c + 32]; s2 += row[c + 64]; s3 += row[c + 96]; }
            for (; c < F.f; c += 32) s0 += row[c];
            s = warp_sum((s0 + s1) + (s2 + s3));
        }
    }
    double t = (u < nbp && q < nbp && u != q) ? M[(long long)(j + u) * F.ld + j + q] : 0.0;
    double g = (u == q) ? 1.0 : 0.0;
    if (u == 0) {
        const double d = warp_sum(t) + s;
        rowT[0][q] = t;
        rowG[0][q] = g;
        if (q == 0) {
            rowS[0] = s;
            rowRD[0] = 1.0 / d;
            Ds[0] = d;
            if (!(d > 0.0) || isinf(d)) atomicOr(errflag, 1);
        }
    }
    for (int p = 0; p < nbp; ++p) {
        __syncthreads();
        const int b = p & 1;
        if (u != p) {
            const double tp = rowT[b][q], gp = rowG[b][q], sp = rowS[b], rd = rowRD[b];
            const double tup = __shfl_sync(0xffffffffu, t, p);
            const double fup = (tup == 0.0) ? 0.0 : tup * rd;   // rows that do not reach p stay clean even if d == 0
            t = (q == p || q == u) ? 0.0 : fma(fup, tp, t);
            g = fma(fup, gp, g);
            s = fma(fup, sp, s);
            if (u == p + 1 && u < nbp) {          // publish the next pivot row one step ahead
                const double d = warp_sum(t) + s;
                rowT[b ^ 1][q] = t;
                rowG[b ^ 1][q] = g;
                if (q == 0) {
                    rowS[b ^ 1] = s;
                    rowRD[b ^ 1] = 1.0 / d;
                    Ds[u] = d;
                    if (!(d > 0.0) || isinf(d)) atomicOr(errflag, 1);
                }
            }
        }
    }
    __syncthreads();
    const double w = (u < nbp) ? g / Ds[u] : 0.0;
    Ws[u][q] = w;
    wbuf[((long long)blockIdx.y * maxfb + blockIdx.x + ko.wslot) * (NB_INNER * NB_INNER) + u * NB_INNER + q] = w;

    // fused row panel for narrow fronts: columns je..f (f = tau column) in slabs of 128 staged through smem
    const int ncols = F.f + 1 - je;
    if (ncols > ko.fuse_cols) return;
    const int cc = threadIdx.x & 127, r0 = threadIdx.x >> 7;
    for (int cs = 0; cs < ncols; cs += 128) {
        __syncthreads();
        const bool live = cs + cc < ncols;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + 8 * i;
            xs[r][cc] = (live && r < nbp) ? M[(long long)(j + r) * F.ld + je + cs + cc] : 0.0;
        }
        __syncthreads();
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int tt = 0; tt < NB_INNER; ++tt) {
            const double x = xs[tt][cc];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fma(Ws[r0 + 8 * i][tt], x, acc[i]);
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + 8 * i;
                if (r < nbp) M[(long long)(j + r) * F.ld + je + cs + cc] = acc[i];
            }
        }
    }
}

// =====================================================================================================
// throughput variant of the panel kernel for launches with many fronts (leaf levels, batched networks, phase-2
// tasks): 256 threads, each thread owns 4 columns of one row, so 4+ panels are resident per SM and their
// latency-bound Gauss-Jordan chains overlap.  Same arithmetic, same order of operations per element.
// =====================================================================================================
__global__ void __launch_bounds__(256, 4)
panel_tp_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j,
                double *wbuf, const double *rsbuf, int maxfb, int maxchunks, int *errflag, KOpt ko) {
    const DFront F = fronts[ids[blockIdx.x]];
    if (j >= F.k) return;
    double *M = pool + (long long)blockIdx.y * pool_stride + F.off;
    const int nbp = min(NB_INNER, F.k - j);
    const int je = j + nbp;
    const int u = threadIdx.x >> 3, cg = threadIdx.x & 7, q0 = cg * 4;
    const int lane = threadIdx.x & 31;

    __shared__ double rowT[2][NB_INNER];
    __shared__ double rowG[2][NB_INNER];
    __shared__ double rowS[2], rowRD[2];
    __shared__ double Ds[NB_INNER];
    __shared__ double Ws[NB_INNER][NB_INNER];
    __shared__ double xs[NB_INNER][64];

    auto group_sum = [](double v) {   // sum over the 8 lanes that share a row
        v += __shfl_xor_sync(0xffffffffu, v, 4, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 2, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 1, 8);
        return v;
    };

    // NOTE: four rows share a warp here, so every shuffle below is executed by all 32 lanes unconditionally
    // (rows beyond nbp contribute zeros); only the use of the result is predicated.
    double s = 0.0;
    {
        const bool live_row = u < nbp;
        if (F.f - je > RS_SPLIT) {
            const int nch = (F.f - je + RS_CHUNK - 1) / RS_CHUNK;
            const double *rs = rsbuf + (((long long)blockIdx.y * maxfb + blockIdx.x) * maxchunks) * NB_INNER;
            if (live_row)
                for (int ch = cg; ch < nch; ch += 8) s += rs[ch * NB_INNER + u];
        } else if (live_row) {
            const double *row = M + (long long)(j + u) * F.ld;
            double s0 = 0.0, s1 = 0.0;
            int c = je + cg;
            for (; c + 8 < F.f; c += 16) { s0 += row[c]; s1 += row[c + 8]; }
            for (; c < F.f; c += 8) s0 += row[c];
            s = s0 + s1;
        }
        s = group_sum(s);
        if (!live_row) s = 1.0;   // padding rows behave like isolated absorbing nodes
    }
    double t[4], g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int q = q0 + e;
        t[e] = (u < nbp && q < nbp && u != q) ? M[(long long)(j + u) * F.ld + j + q] : 0.0;
        g[e] = (u == q) ? 1.0 : 0.0;
    }
    {
        const double d = group_sum((t[0] + t[1]) + (t[2] + t[3])) + s;
        if (u == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { rowT[0][q0 + e] = t[e]; rowG[0][q0 + e] = g[e]; }
            if (cg == 0) {
                rowS[0] = s;
                rowRD[0] = 1.0 / d;
                Ds[0] = d;
                if (!(d > 0.0) || isinf(d)) atomicOr(errflag, 1);
            }
        }
    }
    for (int p = 0; p < nbp; ++p) {
        __syncthreads();
        const int b = p & 1;
        // T[u][p] lives in element p&3 of the lane with column group p>>2
        const int pe = p & 3;
        const double mine = pe == 0 ? t[0] : pe == 1 ? t[1] : pe == 2 ? t[2] : t[3];
        const double tup = __shfl_sync(0xffffffffu, mine, (lane & ~7) | (p >> 2));
        if (u != p) {
            const double sp = rowS[b], rd = rowRD[b];
            const double fup = (tup == 0.0) ? 0.0 : tup * rd;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int q = q0 + e;
                t[e] = (q == p || q == u) ? 0.0 : fma(fup, rowT[b][q], t[e]);
                g[e] = fma(fup, rowG[b][q], g[e]);
            }
            s = fma(fup, sp, s);
        }
        const double d = group_sum((t[0] + t[1]) + (t[2] + t[3])) + s;   // all lanes, see NOTE
        if (u == p + 1 && u < nbp) {          // publish the next pivot row one step ahead
#pragma unroll
            for (int e = 0; e < 4; ++e) { rowT[b ^ 1][q0 + e] = t[e]; rowG[b ^ 1][q0 + e] = g[e]; }
            if (cg == 0) {
                rowS[b ^ 1] = s;
                rowRD[b ^ 1] = 1.0 / d;
                Ds[u] = d;
                if (!(d > 0.0) || isinf(d)) atomicOr(errflag, 1);
            }
        }
    }
    __syncthreads();
    double *Wout = wbuf + ((long long)blockIdx.y * maxfb + blockIdx.x + ko.wslot) * (NB_INNER * NB_INNER);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double w = (u < nbp) ? g[e] / Ds[u] : 0.0;
        Ws[u][q0 + e] = w;
        Wout[u * NB_INNER + q0 + e] = w;
    }
    const int ncols = F.f + 1 - je;
    if (ncols > ko.fuse_cols) return;
    const int cc = threadIdx.x & 63, r0 = threadIdx.x >> 6;
    for (int cs = 0; cs < ncols; cs += 64) {
        __syncthreads();
        const bool live = cs + cc < ncols;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + 4 * i;
            xs[r][cc] = (live && r < nbp) ? M[(long long)(j + r) * F.ld + je + cs + cc] : 0.0;
        }
        __syncthreads();
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
        for (int tt = 0; tt < NB_INNER; ++tt) {
            const double x = xs[tt][cc];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fma(Ws[r0 + 4 * i][tt], x, acc[i]);
        }
        if (live) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 4 * i;
                if (r < nbp) M[(long long)(j + r) * F.ld + je + cs + cc] = acc[i];
            }
        }
    }
}

void launch_panel(const LaunchCtx &c, const int *ids, int nfb, int j) {
    dim3 grid(nfb, c.nbatch);
    if ((long long)nfb * c.nbatch >= 2 * 148)   // many panels: favour occupancy over single-panel latency
        panel_tp_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, c.wbuf, c.rsbuf, c.maxfb,
                                                   c.maxchunks, c.errflag, c.ko);
    else
        panel_kernel<<<grid, 1024, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, c.wbuf, c.rsbuf, c.maxfb,
                                                 c.maxchunks, c.errflag, c.ko);
}

// =====================================================================================================
// row panel for wide fronts: U' = W [P_XR | tau_X] in place, columns je..f (f = tau column)
// 256 threads handle 128 columns: thread (h = tid/128, col = tid%128) produces rows 16h..16h+15
// =====================================================================================================
__global__ void __launch_bounds__(256)
rowpanel_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j,
                const double *wbuf, int maxfb, KOpt ko) {
    const DFront F = fronts[ids[blockIdx.y]];
    if (j >= F.k) return;
    const int nbp = min(NB_INNER, F.k - j);
    const int je = j + nbp;
    if (F.f + 1 - je <= ko.fuse_cols) return;    // done inside the panel kernel
    const int col = je + blockIdx.x * 128 + (threadIdx.x & 127);
    if (je + (int)blockIdx.x * 128 > F.f) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;

    __shared__ double Ws[NB_INNER][NB_INNER];
    const double *W = wbuf + ((long long)blockIdx.z * maxfb + blockIdx.y + ko.wslot) * (NB_INNER * NB_INNER);
    for (int i = threadIdx.x; i < NB_INNER * NB_INNER; i += 256) Ws[i / NB_INNER][i % NB_INNER] = W[i];

    double x[NB_INNER];
    const bool live = col <= F.f && owned_col(ko, col, F.f);
#pragma unroll
    for (int r = 0; r < NB_INNER; ++r) x[r] = (live && r < nbp) ? M[(long long)(j + r) * F.ld + col] : 0.0;
    __syncthreads();
    const int h = threadIdx.x >> 7;
    if (live) {
#pragma unroll 4
        for (int rr = 0; rr < 16; ++rr) {
            const int r = h * 16 + rr;
            if (r >= nbp) break;
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < NB_INNER; ++t) acc += Ws[r][t] * x[t];
            M[(long long)(j + r) * F.ld + col] = acc;
        }
    }
}

void launch_rowpanel(const LaunchCtx &c, const int *ids, int nfb, int j, int max_cols) {
    dim3 grid((max_cols + 127) / 128, nfb, c.nbatch);
    if (grid.x == 0) grid.x = 1;
    rowpanel_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, c.wbuf, c.maxfb, c.ko);
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

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 8 : 0;   // 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 2)
update_dmma_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j, int mode_in,
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
    const int tiles_c = (R.c1 - R.c0 + TN - 1) / TN;
    const int tiles_r = (R.r1 - R.r0 + TM - 1) / TM;
    const int tr = tile / tiles_c, tc = tile % tiles_c;
    if (tr >= tiles_r) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    const double *MA = F.a_ld ? pool + (long long)blockIdx.z * pool_stride + F.a_off : M;
    const long long lda = F.a_ld ? F.a_ld : F.ld;
    const int rb = R.r0 + tr * TM, cb = R.c0 + tc * TN;
    const long long ld = F.ld;
    if (ko.own_G > 1 && !owned_col(ko, cb, F.f) && !owned_col(ko, min(cb + TN, R.c1) - 1, F.f)) return;

    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, t4 = lane & 3;

    // accumulators = C tile (issued first: their latency hides behind the pipeline prologue)
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

    const int nk = (R.k1 - R.k0 + TK - 1) / TK;
    // per-thread load slots: A element (row ar + 16 i, pivot kb + ak), B element (pivot kb + bk + 4 i, column bc)
    const int ar = threadIdx.x >> 4, ak = threadIdx.x & 15;
    const int bk = threadIdx.x >> 6, bc = threadIdx.x & 63;
    const double *pa = MA + (long long)(rb + ar) * lda + ak;
    const double *pb = M + (long long)bk * ld + cb + bc;
    unsigned amask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) amask |= (rb + ar + 16 * i < R.r1) ? (1u << i) : 0u;
    const bool bcol_ok = cb + bc < R.c1;
    double *sa = smem + ar * AS_LD + ak;
    double *sb = smem + TM * AS_LD + bk * BS_LD + bc;
    auto load_stage = [&](int stg, int kb) {
        const bool ka = kb + ak < R.k1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = ka && ((amask >> i) & 1u);
            cp_async8(sa + stg * TSTAGE_DOUBLES + i * 16 * AS_LD, ok ? pa + (long long)i * 16 * lda + kb : M, ok);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = bcol_ok && (kb + bk + 4 * i < R.k1);
            cp_async8(sb + stg * TSTAGE_DOUBLES + i * 4 * BS_LD, ok ? pb + (long long)(kb + 4 * i) * ld : M, ok);
        }
    };
#pragma unroll
    for (int st = 0; st < TSTG - 1; ++st) {
        if (st < nk) load_stage(st, R.k0 + st * TK);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<TSTG - 2>();
        __syncthreads();
        const int nxt = kt + TSTG - 1;
        if (nxt < nk) load_stage(nxt % TSTG, R.k0 + nxt * TK);
        cp_async_commit();
        const double *As = smem + (kt % TSTG) * TSTAGE_DOUBLES;
        const double *Bs = As + TM * AS_LD;
#pragma unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[(wm * 32 + mi * 8 + g) * AS_LD + k4 * 4 + t4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(k4 * 4 + t4) * BS_LD + wn * 32 + ni * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
    // epilogue: thread holds C[g][2*t4 + {0,1}] of every 8 x 8 tile
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

// =====================================================================================================
// Schur update fed by the bulk-copy engine (TMA, `cp.async.bulk` -> SASS UBLKCP) with mbarrier hand-over:
// warp 8 is the producer (it arms the stage's "full" mbarrier with the byte count and issues one bulk copy per tile
// row: 128 B of A per row, up to 512 B of B per pivot), warps 0-7 are the consumers (wait on "full", LDS + DMMA,
// arrive on "empty").  No thread of the math warps touches addresses or predicates of the global loads.
// Same tile geometry, shared-memory layout and accumulation order as update_dmma_kernel, so results are bitwise
// equal.  Used for single-front trailing updates whose operands satisfy the 16-byte alignment rules of bulk
// copies (full outer blocks; the host checks, see Engine::tma_eligible); everything else stays on cp.async.
// =====================================================================================================
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int TMA_THREADS = 288;   // 8 math warps + 1 producer warp

__global__ void __maxnreg__(112)
update_dmma_tma_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j, int mode_in,
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
    const int tiles_c = (R.c1 - R.c0 + TN - 1) / TN;
    const int tiles_r = (R.r1 - R.r0 + TM - 1) / TM;
    const int tr = tile / tiles_c, tc = tile % tiles_c;
    if (tr >= tiles_r) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    const double *MA = F.a_ld ? pool + (long long)blockIdx.z * pool_stride + F.a_off : M;
    const long long lda = F.a_ld ? F.a_ld : F.ld;
    const int rb = R.r0 + tr * TM, cb = R.c0 + tc * TN;
    const long long ld = F.ld;
    if (ko.own_G > 1 && !owned_col(ko, cb, F.f) && !owned_col(ko, min(cb + TN, R.c1) - 1, F.f)) return;

    extern __shared__ double smem[];
    __shared__ unsigned long long full_bar[TSTG], empty_bar[TSTG];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < TSTG; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nk = (R.k1 - R.k0) / TK;                      // host guarantees K % 16 == 0
    const int arows = min(TM, R.r1 - rb);                   // tile rows that exist
    const int bcols = min(TN, ((R.c1 - cb) + 1) & ~1);      // even count; row padding (ld >= f+1 rounded to 8) covers it
    const unsigned stage_bytes = (unsigned)(arows * TK * 8 + TK * bcols * 8);

    if (warp == 8) {
        // ===== producer =====
        for (int kt = 0; kt < nk; ++kt) {
            const int stg = kt % TSTG;
            if (kt >= TSTG) mbar_wait(&empty_bar[stg], ((kt / TSTG) - 1) & 1);
            const int kb = R.k0 + kt * TK;
            double *As = smem + stg * TSTAGE_DOUBLES;
            double *Bs = As + TM * AS_LD;
            if (lane == 0) mbar_expect_tx(&full_bar[stg], stage_bytes);
            __syncwarp();
            for (int r = lane; r < arows; r += 32)
                bulk_g2s(As + r * AS_LD, MA + (long long)(rb + r) * lda + kb, TK * 8, &full_bar[stg]);
            if (lane < TK)
                bulk_g2s(Bs + lane * BS_LD, M + (long long)(kb + lane) * ld + cb, bcols * 8, &full_bar[stg]);
        }
        return;
    }
    // ===== consumers =====
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
    for (int kt = 0; kt < nk; ++kt) {
        const int stg = kt % TSTG;
        mbar_wait(&full_bar[stg], (kt / TSTG) & 1);
        const double *As = smem + stg * TSTAGE_DOUBLES;
        const double *Bs = As + TM * AS_LD;
#pragma unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[(wm * 32 + mi * 8 + g) * AS_LD + k4 * 4 + t4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(k4 * 4 + t4) * BS_LD + wn * 32 + ni * 8 + g];
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
    update_dmma_tmap_kernel<<<grid, 256, smem, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, mode, 0, c.ko,
                                                                  *static_cast<const CUtensorMap *>(tmA),
                                                                  *static_cast<const CUtensorMap *>(tmB));
}

// =====================================================================================================
// Large-region Schur update: CTA tile 128 x 128, 16 warps as 4 (M) x 4 (N), warp tile 32 x 32 = 4 x 4 DMMA tiles,
// K staged 16 at a time through a 4-stage cp.async ring (global -> shared without touching registers), so loads
// of stage t+3 overlap the DMMAs of stage t.  The accumulators start from the C tile itself (loaded while the
// pipeline fills), so the epilogue is a plain store: C = C + A B with no read-modify-write tail.
// =====================================================================================================
constexpr int BM = 128, BN = 128, BK = 16, STG = 4;
constexpr int BA_LD = BK + 4, BB_LD = BN + 4;
constexpr int STAGE_DOUBLES = BM * BA_LD + BK * BB_LD;

__global__ void __launch_bounds__(512, 1)
update_dmma_big_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, int j, int mode_in,
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
    const int tiles_c = (R.c1 - R.c0 + BN - 1) / BN;
    const int tiles_r = (R.r1 - R.r0 + BM - 1) / BM;
    const int tr = tile / tiles_c, tc = tile % tiles_c;
    if (tr >= tiles_r) return;
    double *M = pool + (long long)blockIdx.z * pool_stride + F.off;
    const double *MA = F.a_ld ? pool + (long long)blockIdx.z * pool_stride + F.a_off : M;
    const long long lda = F.a_ld ? F.a_ld : F.ld;
    const int rb = R.r0 + tr * BM, cb = R.c0 + tc * BN;
    const long long ld = F.ld;
    if (ko.own_G > 1 && !owned_col(ko, cb, F.f) && !owned_col(ko, min(cb + BN, R.c1) - 1, F.f)) return;

    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 3, wn = warp >> 2;          // 16 warps as 4 (M) x 4 (N), warp tile 32 x 32
    const int g = lane >> 2, t4 = lane & 3;

    // accumulators = C tile (issued first: their latency hides behind the pipeline prologue)
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

    const int nk = (R.k1 - R.k0 + BK - 1) / BK;
    // per-thread load slots: A element (row ar + 32 i, pivot kb + ak), B element (pivot kb + bk + 4 i, column bc)
    const int ar = threadIdx.x >> 4, ak = threadIdx.x & 15;
    const int bk = threadIdx.x >> 7, bc = threadIdx.x & 127;
    const double *pa = MA + (long long)(rb + ar) * lda + ak;
    const double *pb = M + (long long)bk * ld + cb + bc;
    unsigned amask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) amask |= (rb + ar + 32 * i < R.r1) ? (1u << i) : 0u;
    const bool bcol_ok = cb + bc < R.c1;
    double *sa = smem + ar * BA_LD + ak;
    double *sb = smem + BM * BA_LD + bk * BB_LD + bc;
    auto load_stage = [&](int stg, int kb) {
        const bool ka = kb + ak < R.k1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = ka && ((amask >> i) & 1u);
            cp_async8(sa + stg * STAGE_DOUBLES + i * 32 * BA_LD, ok ? pa + (long long)i * 32 * lda + kb : M, ok);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = bcol_ok && (kb + bk + 4 * i < R.k1);
            cp_async8(sb + stg * STAGE_DOUBLES + i * 4 * BB_LD, ok ? pb + (long long)(kb + 4 * i) * ld : M, ok);
        }
    };
#pragma unroll
    for (int st = 0; st < STG - 1; ++st) {
        if (st < nk) load_stage(st, R.k0 + st * BK);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STG - 2>();
        __syncthreads();
        const int nxt = kt + STG - 1;
        if (nxt < nk) load_stage(nxt % STG, R.k0 + nxt * BK);
        cp_async_commit();
        const double *As = smem + (kt % STG) * STAGE_DOUBLES;
        const double *Bs = As + BM * BA_LD;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[(wm * 32 + mi * 8 + g) * BA_LD + k4 * 4 + t4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(k4 * 4 + t4) * BB_LD + wn * 32 + ni * 8 + g];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();
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

int update_tile_m(int gemm_path) { return gemm_path == 1 ? FT : (gemm_path == 2 ? BM : TM); }
int update_tile_n(int gemm_path) { return gemm_path == 1 ? FT : (gemm_path == 2 ? BN : TN); }
int update_tile_k() { return TK; }

void init_kernel_attributes() {
    cudaFuncSetAttribute(update_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)TSTG * TSTAGE_DOUBLES * sizeof(double)));
    cudaFuncSetAttribute(update_dmma_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)TSTG * TSTAGE_DOUBLES * sizeof(double)));
    cudaFuncSetAttribute(update_dmma_tmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)TMS * TM_STAGE_BYTES));
    cudaFuncSetAttribute(update_dmma_tmap_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);   // two 97 KB CTAs per SM

    cudaFuncSetAttribute(update_dmma_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)STG * STAGE_DOUBLES * sizeof(double)));
}

void launch_update(const LaunchCtx &c, const int *ids, int nfb, int j, int mode, int max_tiles, int split) {
    if (max_tiles <= 0) return;
    dim3 grid(max_tiles, nfb, c.nbatch);
    if (c.gemm_path == 1) {
        update_dfma_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, mode, split, c.ko);
    } else if (c.gemm_path == 2) {
        const size_t smem = (size_t)STG * STAGE_DOUBLES * sizeof(double);
        update_dmma_big_kernel<<<grid, 512, smem, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, mode, split, c.ko);
    } else if (c.gemm_path == 3) {
        const size_t smem = (size_t)TSTG * TSTAGE_DOUBLES * sizeof(double);
        update_dmma_tma_kernel<<<grid, TMA_THREADS, smem, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, mode, split,
                                                                     c.ko);
    } else {
        const size_t smem = (size_t)TSTG * TSTAGE_DOUBLES * sizeof(double);
        update_dmma_kernel<<<grid, 256, smem, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, j, mode, split, c.ko);
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
        const int u = row_node[row];
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
// assembly (extend-add), parent-centric gather: one warp per row of a parent front walks the parent's children in a
// fixed order and adds the matching child entries (inv maps a parent local index to the child's update index or -1):
//   parent[a, b] += child[k_c + inv[a], k_c + inv[b]],   parent[a, tau] += child[k_c + inv[a], tau]
// Every parent entry has exactly one writer, so there are no atomics and the sum order is deterministic; parent rows
// are written with unit stride.  grid (row tiles of 8, parents of the level, networks).
// =====================================================================================================
__global__ void __launch_bounds__(256)
assemble_kernel(double *pool, long long pool_stride, const DFront *fronts, const int *ids, const long long *child_ptr,
                const int *child_ids, const int *inv, const long long *inv_off) {
    const int pid = ids[blockIdx.y];
    const DFront P = fronts[pid];
    const int a = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (a >= P.f) return;
    const int lane = threadIdx.x & 31;
    double *base = pool + (long long)blockIdx.z * pool_stride;
    double *prow = base + P.off + (long long)a * P.ld;
    for (long long ci = child_ptr[pid]; ci < child_ptr[pid + 1]; ++ci) {
        const int cid = child_ids[ci];
        const int *iv = inv + inv_off[cid];
        const int ia = iv[a];
        if (ia < 0) continue;                      // this child does not touch row a (uniform over the warp)
        const DFront C = fronts[cid];
        const double *crow = base + C.off + (long long)(C.k + ia) * C.ld + C.k;
        for (int b = lane; b < P.f; b += 32) {
            const int ib = iv[b];
            if (ib >= 0) {
                const double v = crow[ib];
                if (v != 0.0) prow[b] += v;
            }
        }
        if (lane == 0) {
            const double v = crow[C.f - C.k];      // the child's delta-tau column
            if (v != 0.0) prow[P.f] += v;
        }
    }
}

void launch_assemble(const LaunchCtx &c, const int *ids, int nparents, int max_f, const long long *child_ptr,
                     const int *child_ids, const int *inv, const long long *inv_off) {
    if (nparents <= 0 || max_f <= 0) return;
    dim3 grid((max_f + 7) / 8, nparents, c.nbatch);
    assemble_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, c.fronts, ids, child_ptr, child_ids, inv, inv_off);
}

// =====================================================================================================
// phase 2 (reference ngt.hpp:362-431): for every a of a group, eliminate the rest of the group.
// Task front (order ng+1): rows/cols 0..ng-2 = group minus a, row ng-1 = a, column ng = all mass into the
// other group (collapsed), column ng+1 = tau.  Eliminating its ng-1 pivots with the ordinary front kernels
// leaves 1-P'_aa in (ng-1, ng) and tau'_a in (ng-1, ng+1).
// =====================================================================================================
__global__ void __launch_bounds__(256)
phase2_gather_kernel(const double *pool, long long pool_stride, DFront root, int g0, int ng, int o0, int no,
                     double *p2pool, long long p2_stride, const DFront *p2fronts, int task0, int rank, int nranks) {
    const int t = blockIdx.y;            // task within the group = index of a
    if (nranks > 1 && (t % nranks) != rank) return;
    const int lr = blockIdx.x;           // local row of the task front (0..ng-1)
    const DFront T = p2fronts[task0 + t];
    const double *R = pool + (long long)blockIdx.z * pool_stride + root.off;
    double *M = p2pool + (long long)blockIdx.z * p2_stride + T.off;
    // local index -> group member: members other than a keep root order, a goes last
    const int gr = (lr == ng - 1) ? t : (lr < t ? lr : lr + 1);
    const double *src = R + (long long)(g0 + gr) * root.ld;
    double *dst = M + (long long)lr * T.ld;
    for (int lc = threadIdx.x; lc < ng; lc += 256) {
        const int gc = (lc == ng - 1) ? t : (lc < t ? lc : lc + 1);
        dst[lc] = (lc == lr) ? 0.0 : src[g0 + gc];
    }
    // mass into the other group, summed in a fixed order by warp 0
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int c = threadIdx.x; c < no; c += 32) s += src[o0 + c];
        s = warp_sum(s);
        if (threadIdx.x == 0) {
            dst[ng] = s;
            dst[ng + 1] = src[root.f];
        }
    }
    // the collapsed "other" node is row ng of the task front: never read (it is not a pivot), keep zero
    if (lr == 0)
        for (int lc = threadIdx.x; lc < ng + 2; lc += 256) M[(long long)ng * T.ld + lc] = 0.0;
}

void launch_phase2_gather(const LaunchCtx &c, DFront root, int g0, int ng, int o0, int no, double *p2pool,
                          long long p2_stride, const DFront *p2fronts, int task0, int ntask, int rank, int nranks) {
    if (ntask <= 0) return;
    dim3 grid(ng, ntask, c.nbatch);
    phase2_gather_kernel<<<grid, 256, 0, c.stream>>>(c.pool, c.pool_stride, root, g0, ng, o0, no, p2pool, p2_stride,
                                                     p2fronts, task0, rank, nranks);
}

__global__ void phase2_collect_kernel(const double *p2pool, long long p2_stride, const DFront *p2fronts, int task0,
                                      int ntask, int ng, double *om, double *tau, int out_stride, int out_off,
                                      int rank, int nranks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntask) return;
    const long long b = blockIdx.y;
    double o = 0.0, tt = 0.0;
    if (nranks <= 1 || (t % nranks) == rank) {
        const DFront T = p2fronts[task0 + t];
        const double *M = p2pool + b * p2_stride + T.off;
        o = M[(long long)(ng - 1) * T.ld + ng];
        tt = M[(long long)(ng - 1) * T.ld + ng + 1];
    }
    om[b * out_stride + out_off + t] = o;
    tau[b * out_stride + out_off + t] = tt;
}

void launch_phase2_collect(const LaunchCtx &c, const double *p2pool, long long p2_stride, const DFront *p2fronts,
                           int task0, int ntask, int ng, double *om, double *tau, int out_stride, int out_off,
                           int rank, int nranks) {
    if (ntask <= 0) return;
    dim3 grid((ntask + 127) / 128, c.nbatch);
    phase2_collect_kernel<<<grid, 128, 0, c.stream>>>(p2pool, p2_stride, p2fronts, task0, ntask, ng, om, tau,
                                                     out_stride, out_off, rank, nranks);
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

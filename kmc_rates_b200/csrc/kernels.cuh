// Device-side data layout and kernel launchers of the multifrontal NGT engine (sm_100a).
//
// HBM layout
//   pool            one allocation of `nbatch * pool_stride` doubles.  Network b owns
//                   pool[b*pool_stride ...]; inside it front F is a row-major f x ld matrix at F.off:
//                     rows/cols 0..k-1      pivots (nodes this front eliminates, in elimination order)
//                     rows/cols k..f-1      update set (later nodes the pivots touch)
//                     column f              tau: tau_x of a pivot row, accumulated delta-tau of an update row
//                   ld = roundup(f+1, 8) so every row starts 64-byte aligned.
//   Eliminating the pivots of a front is exactly the reference's remove_node loop restricted to
//   that front (reference source/kmc_rates/ngt.hpp:295-333), blocked:
//     panel X of nb pivots:  W = (I - P_XX)^-1 by subtraction-free Gauss-Jordan, pivots
//                            1-P_xx = sum of the off-diagonal row (ngt.hpp:222-238 does this only for
//                            P_xx >= 0.99; Wales' GTH form is used always here)
//                            U' = W [P_XR | tau_X]                      (panel_kernel, rowpanel_kernel)
//     Schur update:          P_RR += P_RX U' ,  tau_R += P_RX tau'_X    (update kernels, FP64 DMMA)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>

namespace ngt {

// thrown by the launchers when the runtime rejects a launch (mapped to NGT_E_CUDA at the C ABI)
struct LaunchFailure : std::runtime_error {
    explicit LaunchFailure(const std::string &m) : std::runtime_error(m) {}
};

#ifndef NGT_NB_OUTER
#define NGT_NB_OUTER 256
#endif
#ifndef NGT_FUSE_COLS
#define NGT_FUSE_COLS 384
#endif
constexpr int NB_INNER = 32;             // pivots per panel (in-SM Gauss-Jordan block)
constexpr int NB_OUTER = NGT_NB_OUTER;   // pivots per delayed trailing update (K of the big Schur GEMM)

struct DFront {
    long long off;   // offset in the pool (doubles)
    int f, k, ld;    // order (columns; column f is tau), pivots, leading dimension
    int parent;      // front id or -1
    // Optional generalisations used by the multi-GPU dense path (0 = not used):
    int fr;          // number of rows when the front is rectangular (panel fronts); 0 -> f
    int a_ld;        // leading dimension of a separate A-operand matrix for the Schur updates; 0 -> A is this front
    long long a_off; // its pool offset, pre-shifted so that A(r, p) = pool[a_off + r*a_ld + p] in this front's indices
};

// per-launch options that only the multi-GPU dense path changes
struct KOpt {
    int own_G;       // column ownership: column c is owned iff ((c / NB_OUTER) % own_G) == own_g; 0 or 1 -> all owned
    int own_g;
    int wslot;       // slot offset into wbuf (one W per inner panel of an outer block)
    int fuse_cols;   // row-panel fusion threshold (FUSE_COLS; 0 forces the stand-alone row-panel kernel)
};

constexpr int PANEL_CLUSTER = 8; // CTAs per front in the cluster variant of the panel kernel
constexpr int PANEL_CLUSTER_MAX_COLS = 2048;   // widest row panel the cluster variant takes on
constexpr int RS_CHUNK = 512;   // columns per partial-row-sum CTA (one warp per panel row)
constexpr int RS_SPLIT = 4096;  // fronts with more columns right of the panel get multi-CTA row sums
constexpr int FUSE_COLS = NGT_FUSE_COLS; // fronts with at most this many columns apply W to their row panel inside the panel kernel

// Schur-update regions.  A/B: L-shaped strips inside an outer block (K = one panel).  D: the delayed trailing update
// (K = the whole outer block), split for look-ahead into D1 = the strips the NEXT outer block's panels need
// (D1A column strip, D1B row strip) and D2 = the rest, which may run on a second stream while the next panels run.
// AB and D1 are "merged" launch modes: CTAs [0, split) take the first region, the others the second.
enum UpdateMode { MODE_A = 0, MODE_B = 1, MODE_D = 2, MODE_AB = 3, MODE_D1A = 4, MODE_D1B = 5, MODE_D2 = 6, MODE_D1 = 7 };

struct Region { int r0, r1, c0, c1, k0, k1; };

// Rows / columns / pivot range touched by one Schur-update launch (see kernels.cuh).
//   MODE_A: column panel inside the outer block      rows [je,f)  cols [je,JE)   K = [j,je)
//   MODE_B: row panel of the remaining outer pivots  rows [je,JE) cols [JE,f]    K = [j,je)
//   MODE_D: delayed trailing update                  rows [JE,f)  cols [JE,f]    K = [J,JE)
// column f is the tau column, so "cols [..,f]" means c1 = f+1.
__host__ __device__ inline bool owned_col(const KOpt &o, int c, int f) {
    if (o.own_G <= 1) return true;
    return c < f && ((c / NB_OUTER) % o.own_G) == o.own_g;     // the tau column of the big matrix is not used there
}

__host__ __device__ inline bool get_region(int f, int fr, int k, int j, int mode, Region &R) {
    if (fr <= 0) fr = f;
    if (j >= k) return false;
    const int je = (j + NB_INNER < k ? j + NB_INNER : k);
    const int J = (j / NB_OUTER) * NB_OUTER;
    const int JE = (J + NB_OUTER < k ? J + NB_OUTER : k);
    if (mode == MODE_A) { R.r0 = je; R.r1 = fr; R.c0 = je; R.c1 = JE; R.k0 = j; R.k1 = je; }
    else if (mode == MODE_B) { R.r0 = je; R.r1 = (JE < fr ? JE : fr); R.c0 = JE; R.c1 = f + 1; R.k0 = j; R.k1 = je; }
    else {
        if (je != JE) return false;   // delayed updates run once, after the last panel of the outer block
        const int JE2 = (JE + NB_OUTER < k ? JE + NB_OUTER : k);
        R.k0 = J; R.k1 = JE;
        if (mode == MODE_D) { R.r0 = JE; R.r1 = fr; R.c0 = JE; R.c1 = f + 1; }
        else if (mode == MODE_D1A) { R.r0 = JE; R.r1 = fr; R.c0 = JE; R.c1 = JE2; }
        else if (mode == MODE_D1B) { R.r0 = JE; R.r1 = (JE2 < fr ? JE2 : fr); R.c0 = JE2; R.c1 = f + 1; }
        else { R.r0 = JE2; R.r1 = fr; R.c0 = JE2; R.c1 = f + 1; }
    }
    return R.r1 > R.r0 && R.c1 > R.c0;
}


struct LaunchCtx {
    double *pool;
    long long pool_stride;
    int nbatch;
    const DFront *fronts;     // device array of all fronts
    double *wbuf;             // nbatch * maxfb * NB_INNER*NB_INNER
    int maxfb;                // fronts per batch upper bound (stride of wbuf)
    double *rsbuf;            // nbatch * maxfb * maxchunks * NB_INNER partial row sums (wide fronts)
    int maxchunks;
    int *errflag;             // device int
    cudaStream_t stream;
    int gemm_path;            // 0 DMMA (FP64 tensor cores), 1 scalar DFMA (cross-check)
    KOpt ko;                  // {0,0,0,FUSE_COLS} outside the multi-GPU dense path
    int panel_cluster = 0;      // 1: panel kernel runs as clusters of PANEL_CLUSTER CTAs per front and includes the row
                                // sums and the row panel (no rowsum / rowpanel launch for this step)
    int tile_m = 0;             // DMMA Schur update tile height: 0 / 128 = throughput tile, 64 = latency tile
    const DFront *f0 = nullptr; // host copy of the front when the launch covers exactly one (passed by value), else null
};

// ids: device array of front ids in this batch (nfb of them).
void launch_rowsum(const LaunchCtx &c, const int *ids, int nfb, int j, int max_cols);
void launch_panel(const LaunchCtx &c, const int *ids, int nfb, int j);
void launch_rowpanel(const LaunchCtx &c, const int *ids, int nfb, int j, int max_cols);
// max_tiles: upper bound of tiles over the fronts of the batch for this (j, mode); computed by the host.
// mode MODE_AB runs both L-shaped regions in one launch: CTAs [0, split) do MODE_A, the rest MODE_B
void launch_update(const LaunchCtx &c, const int *ids, int nfb, int j, int mode, int max_tiles, int split);

// tensor-map (TMA) fed Schur update of ONE front: tmA / tmB point to CUtensorMap objects over the front's matrix with
// boxes of 16 x 128 and 16 x 16 doubles (128-byte swizzle); see Engine::make_tensor_maps
void launch_update_tmap(const LaunchCtx &c, const int *ids, int j, int mode, int max_tiles, const void *tmA,
                        const void *tmB);

// opt-in shared-memory sizes of the DMMA kernels (call once per process/device before the first launch)
void init_kernel_attributes();

// tile geometry used by the host to size grids
int update_tile_m(int gemm_path);
int update_tile_n(int gemm_path);
int update_tile_k();

// ---- ingest ---------------------------------------------------------------------------------------
// sparse: CSR by source node; korig[e] = index of entry e in the caller's k array; dest[e] = pool offset
void launch_ingest_sparse(double *pool, long long pool_stride, int nbatch, const double *k, long long k_stride,
                          int n_rows, const long long *row_ptr, const int *row_node, const int *korig,
                          const long long *dest, const long long *tau_dest, double *tau0, long long tau0_stride,
                          int *errflag, cudaStream_t s);
// rate -> slot table on the device (see kernels.cu): row_ptr / korig = the host's CSR-by-source edge index (row u =
// node u), dst = the caller's destination ids, pos / pos_front / ffirst / fidx / idx_off = the analysed structure
void launch_slot_table(int n, const long long *row_ptr, const int *korig, const long long *dst, const int *pos,
                       const int *pos_front, const DFront *fronts, const int *ffirst, const int *fidx,
                       const long long *idx_off, long long *dest, long long *tau_dest, int *errflag, cudaStream_t s);
// dense: K is nbatch x n x n row-major; pos = node -> elimination position; front 0 = all intermediates (k0 = nI),
// root = last front
void launch_ingest_dense(double *pool, long long pool_stride, int nbatch, const double *K, int n, const int *pos,
                         int nI, DFront f0, DFront root, double *tau0, int *errflag, cudaStream_t s);

// ---- whole-network kernel: one CTA eliminates one small complete network (ingest + phase 1 fused; see kernels.cu) ------
// K: nbatch x n x n rates; node_at: elimination position -> node id; k: pivots (the kept nodes follow); F0: the front of
// all intermediates (receives the factor rows U'), RF: the root front (receives the reduced A ∪ B graph incl. tau').
constexpr int DENSE_NET_MAX_N = 640;      // row block + Gauss-Jordan scratch of two CTAs fit one SM's shared memory up to here
size_t dense_network_smem(int n);
void launch_dense_network(const double *K, int n, int nbatch, const int *node_at, int k, double *pool, long long pool_stride,
                          DFront F0, DFront RF, double *tau0, int *errflag, cudaStream_t s);

// ---- assembly: child update blocks (+ delta-tau column) are added into their parents -----------------------
// ids: parents of one tree level; row_base / row_ptr / row_child / row_crow: per parent row, the (child, child update
// row) pairs that add into it; rel / rel_off: per child, child update index -> parent local index (ascending).
// Levels with a few large parents run ASM_WPR warps per parent row, each owning one ASM_WPR-th of the parent's columns
// (assemble_splits_rows tells which); child_split[c * (ASM_WPR-1) + i] = first update index of child c mapped at or
// beyond parent column (i+1) * ceil(f_parent / ASM_WPR), tabulated by the host for the children of such parents.
constexpr int ASM_WPR = 4;
bool assemble_splits_rows(int nparents, int nbatch, int max_f);
void launch_assemble(const LaunchCtx &c, const int *ids, int nparents, int max_f, const long long *row_base,
                     const long long *row_ptr, const int *row_child, const int *row_crow, const int *rel,
                     const long long *rel_off, const int *child_split);

// ---- phase 2: a binary tree of eliminations per group (see kernels.cu) -------------------------------
struct P2Task {       // one phase-2 front = "parent's kept members minus T are eliminated, T is kept"
    int src;          // source front (index into the phase-2 front array), or -1: the reduced root block
    int src_k;        // first kept row / column in the source (= its pivot count; root: first row of the group)
    int npk;          // members the source kept (the new front has order npk + 1: + the collapsed other group)
    int t_off, t_len; // T = source's kept members [t_off, t_off + t_len): a prefix or a suffix
    int o0, no;       // root source only: columns [o0, o0 + no) of the other group collapse into one
    int pad;
};
// build the fronts task0 .. task0+ntask-1 (one tree level) from their sources; max_f = largest new order
void launch_phase2_gather(const LaunchCtx &c, DFront root, double *p2pool, long long p2_stride, const DFront *p2fronts,
                          const P2Task *tasks, int task0, int ntask, int max_f);
// read 1-P'_aa and tau'_a from the fronts that kept a single member
void launch_phase2_collect(const LaunchCtx &c, const double *p2pool, long long p2_stride, const DFront *p2fronts,
                           const int *leaf_front, const int *leaf_out, int nleaf, double *om, double *tau,
                           int out_stride);

// ---- multi-GPU dense path helpers (see kernels.cu) --------------------------------------------------------
void launch_dist_rowsum(const double *M, int ld, int f, int J, int JE, KOpt ko, double *out, cudaStream_t s);
// owner: the block's column operations on the panel rows below the block (block columns + tau), one launch; the block
// rows of `pan` must hold the factored square block (see kernels.cu)
void launch_dist_colstrip_slab(double *pan, int ld_pan, int kb, int fr, cudaStream_t s);
// the block's row operations (all inner panels) on the own columns right of the block, one launch (see kernels.cu):
// M = the big matrix, pan = the broadcast panel (rows J.., ld_pan), W = the block's inner-panel inverses (wbuf slots)
void launch_dist_apply_slab(double *M, long long ld, int f, int J, int kb, const double *pan, int ld_pan, const double *W,
                            KOpt ko, cudaStream_t s);
void launch_dist_pack(const double *M, int ld, int J, int kb, int fr, double *pan, int ld_pan, const double *S,
                      const double *tau_glob, cudaStream_t s);
void launch_dist_tau(const double *pan, int ld_pan, int kb, int fr, int J, double *tau_glob, cudaStream_t s);
void launch_dist_tau_init(const double *tau0, const int *pos, int n, double *tau_glob, cudaStream_t s);
void launch_dist_root_contrib(const double *M, int ld, int f, int nI, int m, KOpt ko, double *out, cudaStream_t s);
void launch_dist_root_finish(double *root, int ld_root, int m, const double *contrib, const double *tau_glob, int nI,
                             cudaStream_t s);

// ---- graph validation: connected components (label = smallest node id of the component) ---------------------
void launch_connected_components(int *parent, const long long *src, const long long *dst, long long nnz, int n, int *errflag,
                                 cudaStream_t s);

// ---- committor / MFPT back-substitution over kept factors -------------------------------------------
// x is nbatch x npos x 2 (q, t interleaved) indexed by elimination position; one call handles the panel that
// starts at pivot j of every front in `ids`; the host sweeps levels top-down and panels in reverse.
void launch_backsub_step(const LaunchCtx &c, const int *ids, int nfb, const int *fidx, const long long *idx_off,
                         double *x, long long x_stride, int j);

}  // namespace ngt

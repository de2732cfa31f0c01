/*
 * ngt_b200.h — C ABI of the B200-native New Graph Transformation (NGT) engine.
 *
 * This is the drop-in boundary for ONE hot path of js850/kmc_rates: phase-1 node elimination
 * (P_uv += P_ux P_xv/(1-P_xx), tau_u += P_ux tau_x/(1-P_xx)) + the phase-2 per-a reduction ->
 * k_AB, k_BA, steady-state rates, committors.  Every entry point replaces a call the reference's
 * Cython binding makes into its C++ engine; a maintainer binds these instead.
 *
 *   entry point                          replaces (reference file:line)
 *   ----------------------------------   -----------------------------------------------------------
 *   ngt_create                           new cNGT(rate_map, A, B)            kmc_rates/ngt.pyx:93
 *                                        graph_ns::NGT::NGT(rate_map_t&,..)  source/kmc_rates/ngt.hpp:105-181
 *   ngt_create_dense / _dense_dev        same ctor for a complete rate matrix (no std::map)
 *   ngt_set_weights                      set_node_occupation_probabilities   ngt.pyx:105, ngt.hpp:183-185
 *   ngt_compute_rates                    compute_rates()                     ngt.pyx:119, ngt.hpp:436-439
 *   ngt_compute_rates_and_committors     compute_rates_and_committors()      ngt.pyx:129, ngt.hpp:616-630
 *   ngt_get_rates                        get_rate_AB/BA/AB_SS/BA_SS          ngt.pyx:132-146, ngt.hpp:463-513
 *   ngt_get_final                        final_omPxx / final_tau             ngt.hpp:45-53
 *   ngt_get_committors                   get_committors()                    ngt.pyx:148-154, ngt.hpp:98
 *   ngt_get_reduced                      the A∪B graph left in NGT::_graph   ngt.hpp:36 (GraphReduction.graph,
 *                                                                            kmc_rates/_kmc_rates.py:127-131)
 *   ngt_destroy                          del thisptr                         ngt.pyx:111-114
 *   ngt_batch_*                          a loop of the above over networks that share one topology
 *
 * Conventions: nodes are dense ids 0..n-1 (label -> id mapping stays on the host side, as in
 * ngt.pyx:68-73); plain pointers and sizes only; the caller owns every buffer; every call is
 * synchronous on return; every function returns 0 on success or an NGT_E_* code, with a message in
 * ngt_last_error().  There is NO CPU fallback: without a CUDA device every compute call fails with
 * NGT_E_CUDA.
 */
#ifndef NGT_B200_H_
#define NGT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ngt_handle_s *ngt_handle;

enum {
    NGT_OK = 0,
    NGT_E_INVALID = 1,   /* bad argument: id out of range, A∩B≠∅, empty B, node without out-rates (A may be empty:
                            MFPT-to-B mode, rates AB then read NaN) */
    NGT_E_CUDA = 2,      /* CUDA runtime error (no device, out of memory, launch failure) */
    NGT_E_STATE = 3,     /* call order: results requested before compute_* */
    NGT_E_NUMERIC = 4    /* a pivot 1-P_xx was zero or not finite (dead-end node) */
};

/* Options (0 = default everywhere). */
typedef struct ngt_options {
    int32_t device;          /* CUDA device ordinal */
    int32_t gemm_path;       /* 0 = FP64 tensor-core DMMA, tile and feed (cp.async / TMA) chosen per launch (default); 1 = scalar DFMA (cross-check); 2 = DMMA with the plain 128x64 cp.async kernel everywhere (cross-check) */
    int32_t leaf_size;       /* nested-dissection leaf size (0 -> 64) */
    int32_t host_sweeps;     /* 1 = every level-structure sweep of the ordering on the host (cross-check); 0 = the sweeps of
                                the largest subgraphs on the device (default; the ordering is the same either way) */
    int32_t sweep_min_nodes; /* subgraphs of the ordering with at least this many nodes are swept on the device (0 -> 65536) */
    int32_t reserved[11];    /* must be zero */
} ngt_options;

/* ---- construction ------------------------------------------------------------------------- */

/* Sparse network from COO rate constants k[e] : src[e] -> dst[e] (host pointers).  A duplicate (src,dst)
 * overwrites the earlier one, as the reference's std::map does.  Does ingest + ordering + symbolic analysis. */
int ngt_create(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, const double *k,
               int64_t nA, const int64_t *A, int64_t nB, const int64_t *B,
               const ngt_options *opt, ngt_handle *out);

/* Complete network from a dense row-major n x n rate matrix K (K[u*n+v] = k_uv, diagonal = self rate,
 * usually 0).  K is a HOST pointer for ngt_create_dense and a DEVICE pointer for ngt_create_dense_dev. */
int ngt_create_dense(int64_t n, const double *K_host, int64_t nA, const int64_t *A, int64_t nB, const int64_t *B,
                     const ngt_options *opt, ngt_handle *out);
int ngt_create_dense_dev(int64_t n, const double *K_dev, int64_t nA, const int64_t *A, int64_t nB, const int64_t *B,
                         const ngt_options *opt, ngt_handle *out);

/* Replace the rate constants of an existing handle, keeping topology, ordering and symbolic analysis
 * (temperature / rate sweeps).  k has the length and order given at creation.  Host or device pointer. */
int ngt_set_rates(ngt_handle h, const double *k_host);
int ngt_set_rates_dev(ngt_handle h, const double *k_dev);

int ngt_set_weights(ngt_handle h, int64_t nw, const int64_t *ids, const double *w);

void ngt_destroy(ngt_handle h);

/* ---- compute ------------------------------------------------------------------------------ */
int ngt_compute_rates(ngt_handle h);
int ngt_compute_rates_and_committors(ngt_handle h);
/* phase 1 only (device resident; no result download) — used by bench.py for the HBM-resident timing */
int ngt_phase_one(ngt_handle h);
int ngt_phase_two(ngt_handle h);

/* ---- results ------------------------------------------------------------------------------ */
int ngt_get_rates(ngt_handle h, double out4[4]);                 /* k_AB, k_BA, k_AB^SS, k_BA^SS */
int ngt_get_final(ngt_handle h, double *omPxx, double *tau);      /* nbatch x (|A|+|B|) each, A then B in the order passed */
int ngt_get_reduced(ngt_handle h, double *P, double *tau);        /* (|A|+|B|)^2 row-major, |A|+|B| */
int ngt_get_initial_tau(ngt_handle h, double *tau);               /* n */
/* node ids in elimination order: the intermediates, then A, then B (count = number of nodes named by a rate and not
 * pruned); writes at most cap entries and the count */
int ngt_get_elimination_order(ngt_handle h, int64_t *nodes, int64_t cap, int64_t *count);
int ngt_get_committors(ngt_handle h, double *q);                  /* n; q_A = 0, q_B = 1 */
int ngt_get_mfpt(ngt_handle h, double *t);                        /* n; mean first-passage time to A∪B */

/* ---- statistics for roofline accounting --------------------------------------------------- */
typedef struct ngt_stats {
    int64_t n_nodes, n_rates, n_eliminated;
    int64_t n_fronts, n_levels, n_launches;
    int64_t max_front, pool_doubles;
    double  alg_flops;          /* sum over eliminated nodes of 2 d_in d_out + 3 d_in + d_out */
    double  alg_bytes;          /* sum of 16 d_in d_out + 12 (d_in+d_out) + 16 d_in + 8 */
    double  schur_flops;        /* flops issued by the Schur-update GEMM kernel */
    double  ms_symbolic, ms_ingest, ms_phase1, ms_phase2, ms_schur;   /* last run; device times from CUDA events */
    double  reserved[8];
} ngt_stats;
int ngt_get_stats(ngt_handle h, ngt_stats *st);
/* Device-timed repetition for bench.py: `steps` x (ingest + phase 1 + phase 2) enqueued back to back on the
 * engine's stream, bracketed by CUDA events on that stream; results of the last step are then downloaded. */
int ngt_timed_steps(ngt_handle h, int32_t steps, double *ms_total);
/* One phase-1 pass with CUDA events around every Schur-update launch: total kernel time (ms), the flops those
 * launches performed, and their count.  Slower than a plain pass; never used inside a timed region. */
int ngt_profile_schur(ngt_handle h, double *ms_schur, double *flops, int64_t *launches);

/* ---- batched networks sharing one topology (temperature sweeps) --------------------------- */
/* k is nbatch x nnz (row-major): network b uses k[b*nnz .. (b+1)*nnz). */
int ngt_batch_create(int64_t nbatch, int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst,
                     const double *k, int64_t nA, const int64_t *A, int64_t nB, const int64_t *B,
                     const ngt_options *opt, ngt_handle *out);
/* dense twin: K is nbatch x n x n, HOST or DEVICE pointer */
int ngt_batch_create_dense(int64_t nbatch, int64_t n, const double *K_host, int64_t nA, const int64_t *A,
                           int64_t nB, const int64_t *B, const ngt_options *opt, ngt_handle *out);
int ngt_batch_create_dense_dev(int64_t nbatch, int64_t n, const double *K_dev, int64_t nA, const int64_t *A,
                               int64_t nB, const int64_t *B, const ngt_options *opt, ngt_handle *out);
/* rates: nbatch x 4 */
int ngt_batch_get_rates(ngt_handle h, double *out);

/* ---- multi-GPU phase 2: shard the per-a reductions ------------------
 * Phase 2 is a binary tree of eliminations per group (each a is still reached by eliminating group \ {a}; shared
 * prefixes are computed once, 4/3 |A|^3 flops instead of the reference's |A|^4).  A rank owns a contiguous block
 * of every group's members and builds only the tree branches that lead to them. */
/* Device pointer + size (doubles) of the reduced (|A|+|B|) x ld block incl. tau column, valid after
 * ngt_phase_one; a rank broadcasts it with NCCL into the same buffer of its peers. */
int ngt_reduced_block_dev(ngt_handle h, double **dptr, int64_t *ndoubles);
/* Run phase 2 only for this rank's members: member i of a group of n belongs to the rank whose block
 * [r*n/nranks ...) contains i (blocks differ by at most one; remainder to the low ranks).  Entries of other ranks
 * in the omPxx/tau result are 0, so an all-reduce(sum) merges the shards. */
int ngt_phase_two_shard(ngt_handle h, int32_t rank, int32_t nranks);
int ngt_set_final(ngt_handle h, const double *omPxx, const double *tau);   /* install merged results, nbatch x (|A|+|B|) each */
/* Device pointers of the phase-2 results (nbatch x (|A|+|B|) doubles each, valid after ngt_phase_two*): a rank
 * all-reduces them in place with NCCL and then calls ngt_adopt_final_dev, which downloads them and forms the rates. */
int ngt_final_dev(ngt_handle h, double **omPxx_dev, double **tau_dev, int64_t *ndoubles);
int ngt_adopt_final_dev(ngt_handle h);

/* ---- multi-GPU: ONE dense network over several GPUs (SURVEY §8e, config 3 at N > 1) ----------------------
 * 1-D block-cyclic column partition (block = *block of ngt_dist_info) with a panel broadcast per outer block.
 * Every rank creates the same dense handle (ngt_create_dense*), then drives, for J = 0, block, 2*block, ... < n_pivots:
 *     ngt_dist_rowsum(h, J, &p, &n)     all-reduce(sum) the n doubles at device pointer p over the ranks
 *     ngt_dist_factor(h, J)             only on the owner, rank (J / block) % world
 *     ngt_dist_panel(h, J, &p, &n)      broadcast the n doubles at p from the owner
 *     ngt_dist_apply(h, J)              every rank
 * and finally ngt_dist_root (all-reduce(sum) its buffer), ngt_dist_finish, then ngt_phase_two / getters as usual.
 * These staged calls are synchronous and leave the collectives to the caller (used by the emulated-rank tests);
 * ngt_dist_phase_one below runs the same stages from C++ with NCCL and no host synchronisation between blocks. */
int ngt_dist_begin(ngt_handle h, int32_t rank, int32_t world);
int ngt_dist_info(ngt_handle h, int64_t *n_pivots, int64_t *block);
int ngt_dist_rowsum(ngt_handle h, int64_t J, double **dptr, int64_t *ndoubles);
int ngt_dist_factor(ngt_handle h, int64_t J);
int ngt_dist_panel(ngt_handle h, int64_t J, double **dptr, int64_t *ndoubles);
int ngt_dist_apply(ngt_handle h, int64_t J);
int ngt_dist_root(ngt_handle h, double **dptr, int64_t *ndoubles);
int ngt_dist_finish(ngt_handle h);

/* ---- graph validation on the device ------------------------------------------------------
 * Connected components of the undirected structure of the rate graph (src[e] -- dst[e]) by a lock-free
 * union-find kernel: labels[u] = smallest node id of u's component, *ncomp (may be NULL) = number of components
 * (ids that appear in no rate are components of their own).  Replaces the networkx sweeps the reference runs
 * before it accepts A and B (kmc_rates/_kmc_rates.py:384-454) and in reduce_rates (rates_linalg.py:13-41).
 * Host pointers; labels has n entries. */
int ngt_connected_components(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, int32_t device,
                             int32_t *labels, int64_t *ncomp);

/* ---- one level-structure sweep of the ordering on the device (diagnostic; csrc/ordering.cu) -------------------
 * Breadth-first sweep of a connected, symmetric CSR graph (no self loops, no duplicate entries in a row) in the
 * serial queue order.  start >= 0: from that node; start < 0: from node 0, then again from the minimum-degree node of
 * the last level unless that is node 0 (the pseudo-peripheral step of the nested dissection, symbolic.cpp).
 * Returns NGT_OK with *handled = 1 and lvl[n], visit[n], *nlev filled, or *handled = 0 when the kernel declines
 * (graph not connected, levels thinner than ~24 nodes on average, a level above 16384 nodes, n above ~1.4e6): the
 * analysis sweeps such graphs on the host.  Host pointers. */
int ngt_level_sweep(int32_t device, int32_t n, const int32_t *ptr, const int32_t *adj, int32_t start, int32_t *lvl,
                    int32_t *visit, int32_t *nlev, int32_t *handled);

/* ---- PATHSAMPLE ingest on the device (reference kmc_rates/utils.py:29-95) ---------------------
 * Rate constants of a min.data / ts.data database: mindata is nmin x 3 (energy, fvib, point-group order), tsdata
 * nts x 3 (the same for the transition states), m1 / m2 the 0-based minima each transition state joins, T the
 * temperature.  Both directions of every transition state are evaluated in log space, shifted by the largest log
 * rate (returned in *log_rate_subtracted; rate_norm = exp(-that)), transition states between the same ordered
 * pair summed (log-sum-exp), self-connecting ones dropped.  Output: *n_out directed rates src[i] -> dst[i] (0-based,
 * sorted by (src, dst)) with rate constants k[i]; the three arrays must hold 2*nts entries.  Host pointers.
 * Errors are reported through ngt_pathsample_last_error(). */
int ngt_pathsample_rates(int64_t nmin, int64_t nts, const double *mindata, const double *tsdata, const int64_t *m1,
                         const int64_t *m2, double T, int32_t device, int64_t *n_out, int64_t *src, int64_t *dst,
                         double *k, double *log_rate_subtracted);
const char *ngt_pathsample_last_error(void);

/* ---- multi-GPU, driven from C++ with NCCL on the engine's own streams --------------------------------------
 * One communicator per process and GPU.  NCCL is bound at run time (dlopen of the libnccl.so.2 the host process
 * already uses); the 128-byte unique id is made on one rank (ngt_comm_unique_id) and handed to the others by
 * whatever the host has (torch.distributed, MPI, a file).  ngt_comm_create is collective over the ranks.
 *   ngt_dist_phase_one      phase 1 of ONE dense network over the ranks (every rank holds a handle created from the
 *                           same rate matrix): the whole block loop -- row-sum all-reduce, panel factorisation on
 *                           the owner, panel broadcast, trailing update of the own columns with look-ahead -- is
 *                           enqueued without a host synchronisation between blocks.  Then ngt_phase_two / getters.
 *   ngt_phase_two_sharded   after phase 1 on every rank (replicated, or rank 0's reduced block broadcast when
 *                           broadcast_root != 0): this rank's share of phase 2, one in-place all-reduce of the
 *                           (1-P'_aa, tau'_a) vectors, rates available on every rank.
 * ms_device (may be NULL) receives the device time of the call from CUDA events on the engine's main stream. */
typedef struct ngt_comm_s *ngt_comm;
int ngt_comm_unique_id(uint8_t id[128]);
int ngt_comm_create(int32_t device, const uint8_t id[128], int32_t rank, int32_t world, ngt_comm *out);
void ngt_comm_destroy(ngt_comm c);
int ngt_dist_phase_one(ngt_handle h, ngt_comm c, double *ms_device);
int ngt_phase_two_sharded(ngt_handle h, ngt_comm c, int32_t broadcast_root, double *ms_device);

/* Device buffers of destroyed handles are cached per process for reuse by later handles (cudaMalloc/cudaFree of the
 * front pools cost ~0.1 s), and so are the host temporaries of the sparse analysis (up to 1 GB; their first-touch page
 * faults cost more than the analysis passes that fill them).  This returns both to the driver / the allocator. */
void ngt_trim_device_cache(void);

const char *ngt_last_error(void);
const char *ngt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NGT_B200_H_ */

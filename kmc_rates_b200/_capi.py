"""ctypes binding of the C ABI in include/ngt_b200.h (kmc_rates_b200/csrc/libngt_b200.so).

This is the thin layer the north star asks for: Python host code -> C ABI -> hand-written sm_100a
kernels.  It plays the role of the reference's Cython shim (reference kmc_rates/ngt.pyx:12-28: the
`cdef extern` block and `except +` error translation).  There is no fallback: if the library is not
built, or no CUDA device is present, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# NGT_B200_LIB selects another build of the same library (A/B timing of kernel variants; scripts/build_variant.sh)
LIB_PATH = os.environ.get("NGT_B200_LIB") or os.path.join(_HERE, "csrc", "libngt_b200.so")

NGT_OK, NGT_E_INVALID, NGT_E_CUDA, NGT_E_STATE, NGT_E_NUMERIC = 0, 1, 2, 3, 4

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


class Options(ctypes.Structure):
    _fields_ = [
        ("device", ctypes.c_int32),
        ("gemm_path", ctypes.c_int32),
        ("leaf_size", ctypes.c_int32),
        ("host_sweeps", ctypes.c_int32),
        ("sweep_min_nodes", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 11),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("n_nodes", ctypes.c_int64), ("n_rates", ctypes.c_int64), ("n_eliminated", ctypes.c_int64),
        ("n_fronts", ctypes.c_int64), ("n_levels", ctypes.c_int64), ("n_launches", ctypes.c_int64),
        ("max_front", ctypes.c_int64), ("pool_doubles", ctypes.c_int64),
        ("alg_flops", ctypes.c_double), ("alg_bytes", ctypes.c_double), ("schur_flops", ctypes.c_double),
        ("ms_symbolic", ctypes.c_double), ("ms_ingest", ctypes.c_double), ("ms_phase1", ctypes.c_double),
        ("ms_phase2", ctypes.c_double), ("ms_schur", ctypes.c_double),
        ("reserved", ctypes.c_double * 8),
    ]

    def as_dict(self):
        d = {f: getattr(self, f) for f, _t in self._fields_ if f != "reserved"}
        d["schur_launches"] = self.reserved[0]
        d["ms_profiled_pass"] = self.reserved[1]
        d["h2d_bytes_total"] = int(self.reserved[2])   # process-wide counters of the library's host<->device copies
        d["d2h_bytes_total"] = int(self.reserved[3])
        return d


class NGTNumericError(ArithmeticError):
    """a pivot 1-P_xx was zero / not finite (dead-end node)"""


_lib = None


def lib():
    """Load the CUDA library.  Raises ImportError when it has not been built — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "kmc_rates_b200: CUDA extension %s is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    H = ctypes.c_void_p
    HP = ctypes.POINTER(ctypes.c_void_p)
    OP = ctypes.POINTER(Options)
    i64, vp = ctypes.c_int64, ctypes.c_void_p
    sig = {
        "ngt_create": [i64, i64, _i64p, _i64p, vp, i64, _i64p, i64, _i64p, OP, HP],
        "ngt_create_dense": [i64, vp, i64, _i64p, i64, _i64p, OP, HP],
        "ngt_create_dense_dev": [i64, vp, i64, _i64p, i64, _i64p, OP, HP],
        "ngt_set_rates": [H, vp],
        "ngt_set_rates_dev": [H, vp],
        "ngt_set_weights": [H, i64, _i64p, _f64p],
        "ngt_compute_rates": [H],
        "ngt_compute_rates_and_committors": [H],
        "ngt_phase_one": [H],
        "ngt_phase_two": [H],
        "ngt_get_rates": [H, _f64p],
        "ngt_get_final": [H, _f64p, _f64p],
        "ngt_get_reduced": [H, _f64p, _f64p],
        "ngt_get_initial_tau": [H, _f64p],
        "ngt_get_elimination_order": [H, _i64p, i64, _i64p],
        "ngt_level_sweep": [ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                            ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                            ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)],
        "ngt_get_committors": [H, _f64p],
        "ngt_get_mfpt": [H, _f64p],
        "ngt_get_stats": [H, ctypes.POINTER(Stats)],
        "ngt_timed_steps": [H, ctypes.c_int32, _f64p],
        "ngt_profile_schur": [H, _f64p, _f64p, _i64p],
        "ngt_batch_create": [i64, i64, i64, _i64p, _i64p, vp, i64, _i64p, i64, _i64p, OP, HP],
        "ngt_batch_create_dense": [i64, i64, vp, i64, _i64p, i64, _i64p, OP, HP],
        "ngt_batch_create_dense_dev": [i64, i64, vp, i64, _i64p, i64, _i64p, OP, HP],
        "ngt_batch_get_rates": [H, _f64p],
        "ngt_reduced_block_dev": [H, ctypes.POINTER(vp), _i64p],
        "ngt_phase_two_shard": [H, ctypes.c_int32, ctypes.c_int32],
        "ngt_set_final": [H, _f64p, _f64p],
        "ngt_final_dev": [H, ctypes.POINTER(vp), ctypes.POINTER(vp), _i64p],
        "ngt_adopt_final_dev": [H],
        "ngt_dist_begin": [H, ctypes.c_int32, ctypes.c_int32],
        "ngt_dist_info": [H, _i64p, _i64p],
        "ngt_dist_rowsum": [H, i64, ctypes.POINTER(vp), _i64p],
        "ngt_dist_factor": [H, i64],
        "ngt_dist_panel": [H, i64, ctypes.POINTER(vp), _i64p],
        "ngt_dist_apply": [H, i64],
        "ngt_dist_root": [H, ctypes.POINTER(vp), _i64p],
        "ngt_dist_finish": [H],
        "ngt_connected_components": [i64, i64, _i64p, _i64p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), _i64p],
        "ngt_pathsample_rates": [i64, i64, _f64p, _f64p, _i64p, _i64p, ctypes.c_double, ctypes.c_int32, _i64p, _i64p, _i64p,
                                 _f64p, _f64p],
        "ngt_comm_unique_id": [ctypes.POINTER(ctypes.c_uint8)],
        "ngt_comm_create": [ctypes.c_int32, ctypes.POINTER(ctypes.c_uint8), ctypes.c_int32, ctypes.c_int32, HP],
        "ngt_dist_phase_one": [H, H, _f64p],
        "ngt_phase_two_sharded": [H, H, ctypes.c_int32, _f64p],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = ctypes.c_int
    L.ngt_destroy.argtypes = [H]
    L.ngt_destroy.restype = None
    L.ngt_comm_destroy.argtypes = [H]
    L.ngt_comm_destroy.restype = None
    L.ngt_trim_device_cache.argtypes = []
    L.ngt_trim_device_cache.restype = None
    L.ngt_last_error.restype = ctypes.c_char_p
    L.ngt_pathsample_last_error.restype = ctypes.c_char_p
    L.ngt_version.restype = ctypes.c_char_p
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "ngt_create", "ngt_create_dense", "ngt_create_dense_dev", "ngt_set_rates", "ngt_set_rates_dev", "ngt_set_weights",
    "ngt_destroy", "ngt_compute_rates", "ngt_compute_rates_and_committors", "ngt_phase_one", "ngt_phase_two",
    "ngt_get_rates", "ngt_get_final", "ngt_get_reduced", "ngt_get_initial_tau", "ngt_get_elimination_order", "ngt_level_sweep", "ngt_get_committors", "ngt_get_mfpt",
    "ngt_get_stats", "ngt_timed_steps", "ngt_profile_schur", "ngt_batch_create", "ngt_batch_create_dense", "ngt_batch_create_dense_dev", "ngt_batch_get_rates",
    "ngt_reduced_block_dev", "ngt_phase_two_shard", "ngt_set_final", "ngt_final_dev", "ngt_adopt_final_dev", "ngt_dist_begin", "ngt_dist_info", "ngt_dist_rowsum",
    "ngt_dist_factor", "ngt_dist_panel", "ngt_dist_apply", "ngt_dist_root", "ngt_dist_finish", "ngt_trim_device_cache",
    "ngt_connected_components", "ngt_pathsample_rates", "ngt_pathsample_last_error", "ngt_comm_unique_id", "ngt_comm_create", "ngt_comm_destroy", "ngt_dist_phase_one", "ngt_phase_two_sharded",
    "ngt_last_error",
    "ngt_version",
]


def check(status):
    """Translate a status code the way ngt.pyx's `except +` translates C++ exceptions."""
    if status == NGT_OK:
        return
    msg = lib().ngt_last_error().decode("utf-8", "replace")
    if status == NGT_E_INVALID:
        raise ValueError(msg)
    if status == NGT_E_NUMERIC:
        raise NGTNumericError(msg)
    raise RuntimeError(msg)


def _ids(seq):
    a = np.ascontiguousarray(list(seq), np.int64)
    return a, a.ctypes.data_as(_i64p)


def _is_torch_cuda(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and x.is_cuda


def connected_components(n, src, dst, device=0):
    """(number of components, labels) of the undirected structure of the rate graph, computed by the union-find kernel
    (include/ngt_b200.h, ngt_connected_components); labels[u] = smallest node id of u's component."""
    src = np.ascontiguousarray(src, np.int64)
    dst = np.ascontiguousarray(dst, np.int64)
    if src.size != dst.size:
        raise ValueError("src and dst have different lengths")
    labels = np.empty(int(n), np.int32)
    ncomp = ctypes.c_int64()
    check(lib().ngt_connected_components(int(n), src.size, src.ctypes.data_as(_i64p), dst.ctypes.data_as(_i64p), int(device),
                                         labels.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.byref(ncomp)))
    return ncomp.value, labels


def level_sweep(ptr, adj, start=-1, device=0):
    """One level-structure sweep of the ordering on the device (include/ngt_b200.h, ngt_level_sweep): returns
    (lvl, visit, nlev), or None when the kernel declines the graph."""
    ptr = np.ascontiguousarray(ptr, np.int32)
    adj = np.ascontiguousarray(adj, np.int32)
    n = ptr.size - 1
    lvl = np.empty(n, np.int32)
    visit = np.empty(n, np.int32)
    nlev = ctypes.c_int32()
    handled = ctypes.c_int32()
    i32p = ctypes.POINTER(ctypes.c_int32)
    check(lib().ngt_level_sweep(int(device), n, ptr.ctypes.data_as(i32p), adj.ctypes.data_as(i32p), int(start),
                                lvl.ctypes.data_as(i32p), visit.ctypes.data_as(i32p), ctypes.byref(nlev),
                                ctypes.byref(handled)))
    return (lvl, visit, nlev.value) if handled.value else None


def pathsample_rates(mindata, tsdata, m1, m2, T, device=0):
    """PATHSAMPLE rate constants on the device (include/ngt_b200.h, ngt_pathsample_rates): returns
    (src, dst, k, log_rate_subtracted) with 0-based minima, sorted by (src, dst)"""
    mindata = np.ascontiguousarray(mindata, np.float64).reshape(-1, 3)
    tsdata = np.ascontiguousarray(tsdata, np.float64).reshape(-1, 3)
    m1 = np.ascontiguousarray(m1, np.int64)
    m2 = np.ascontiguousarray(m2, np.int64)
    nts = tsdata.shape[0]
    if m1.size != nts or m2.size != nts:
        raise ValueError("tsdata, m1 and m2 have different lengths")
    src, dst, k = np.empty(2 * nts, np.int64), np.empty(2 * nts, np.int64), np.empty(2 * nts, np.float64)
    n_out, shift = ctypes.c_int64(), ctypes.c_double()
    st = lib().ngt_pathsample_rates(mindata.shape[0], nts, mindata.ctypes.data_as(_f64p), tsdata.ctypes.data_as(_f64p),
                                    m1.ctypes.data_as(_i64p), m2.ctypes.data_as(_i64p), float(T), int(device),
                                    ctypes.byref(n_out), src.ctypes.data_as(_i64p), dst.ctypes.data_as(_i64p),
                                    k.ctypes.data_as(_f64p), ctypes.byref(shift))
    if st != NGT_OK:
        msg = lib().ngt_pathsample_last_error().decode("utf-8", "replace")
        raise (ValueError if st == NGT_E_INVALID else RuntimeError)(msg)
    n = n_out.value
    return src[:n].copy(), dst[:n].copy(), k[:n].copy(), shift.value


class Comm(object):
    """One NCCL communicator per process and GPU, created by the library itself (include/ngt_b200.h, ngt_comm_*).

    `exchange(id_bytes_or_None) -> id_bytes` hands rank 0's 128-byte unique id to the other ranks; from_torch() does it
    with torch.distributed (any backend), which is only plumbing here -- the collectives of the sharded paths are NCCL
    calls made by the C++ engine on its own streams."""

    def __init__(self, device, rank, world, exchange):
        self.c = ctypes.c_void_p()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        buf = (ctypes.c_uint8 * 128)()
        if rank == 0:
            check(lib().ngt_comm_unique_id(buf))
        got = exchange(bytes(buf) if rank == 0 else None)
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(got)
        check(lib().ngt_comm_create(self.device, buf, self.rank, self.world, ctypes.byref(self.c)))

    @classmethod
    def from_torch(cls, device=None, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()

        def exchange(b):
            box = [b]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            return box[0]
        return cls(device, rank, world, exchange)

    def close(self):
        if self.c:
            lib().ngt_comm_destroy(self.c)
            self.c = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Handle(object):
    """Owns one ngt_handle (like BaseNGT owns its cNGT*, reference ngt.pyx:62,93,111-114)."""

    def __init__(self):
        self.h = ctypes.c_void_p()
        self._keep = []   # buffers that must outlive the handle (device tensors borrowed by the engine)
        self.n = 0
        self.nA = self.nB = 0
        self.nbatch = 1

    @staticmethod
    def _opt(device=0, gemm_path=0, leaf_size=0, host_sweeps=0, sweep_min_nodes=0):
        o = Options()
        o.device, o.gemm_path, o.leaf_size, o.host_sweeps = int(device), int(gemm_path), int(leaf_size), int(host_sweeps)
        o.sweep_min_nodes = int(sweep_min_nodes)
        return o

    @classmethod
    def from_coo(cls, n, src, dst, k, A, B, nbatch=1, **opt):
        self = cls()
        src = np.ascontiguousarray(src, np.int64)
        dst = np.ascontiguousarray(dst, np.int64)
        k = np.ascontiguousarray(k, np.float64)
        if k.size != src.size * nbatch or dst.size != src.size:
            raise ValueError("src, dst, k have inconsistent lengths")
        a, ap = _ids(A)
        b, bp = _ids(B)
        o = cls._opt(**opt)
        L = lib()
        if nbatch == 1:
            st = L.ngt_create(int(n), src.size, src.ctypes.data_as(_i64p), dst.ctypes.data_as(_i64p),
                              k.ctypes.data, a.size, ap, b.size, bp, ctypes.byref(o), ctypes.byref(self.h))
        else:
            st = L.ngt_batch_create(int(nbatch), int(n), src.size, src.ctypes.data_as(_i64p), dst.ctypes.data_as(_i64p),
                                    k.ctypes.data, a.size, ap, b.size, bp, ctypes.byref(o), ctypes.byref(self.h))
        check(st)
        self.n, self.nA, self.nB, self.nbatch = int(n), a.size, b.size, int(nbatch)
        return self

    @classmethod
    def from_dense(cls, K, A, B, **opt):
        """K: (n, n) or (nbatch, n, n) rate matrix; numpy array (host) or torch CUDA float64 tensor (zero copy)."""
        self = cls()
        a, ap = _ids(A)
        b, bp = _ids(B)
        L = lib()
        shape = tuple(K.shape)
        if len(shape) == 2:
            nbatch, n = 1, shape[0]
        elif len(shape) == 3:
            nbatch, n = shape[0], shape[1]
        else:
            raise ValueError("K must be (n, n) or (nbatch, n, n)")
        if shape[-1] != n:
            raise ValueError("K must be square")
        if _is_torch_cuda(K):
            import torch
            if K.dtype != torch.float64:
                raise ValueError("device rate matrix must be float64")
            K = K.contiguous()
            torch.cuda.current_stream(K.device).synchronize()   # the producer of K must be done before the engine reads it
            opt = dict(opt)
            opt["device"] = K.device.index or 0
            ptr, dev = ctypes.c_void_p(K.data_ptr()), True
            self._keep.append(K)
        else:
            K = np.ascontiguousarray(K, np.float64)
            ptr, dev = ctypes.c_void_p(K.ctypes.data), False
            self._keep.append(K)
        o = cls._opt(**opt)
        if nbatch == 1:
            fn = L.ngt_create_dense_dev if dev else L.ngt_create_dense
            st = fn(int(n), ptr, a.size, ap, b.size, bp, ctypes.byref(o), ctypes.byref(self.h))
        else:
            fn = L.ngt_batch_create_dense_dev if dev else L.ngt_batch_create_dense
            st = fn(int(nbatch), int(n), ptr, a.size, ap, b.size, bp, ctypes.byref(o), ctypes.byref(self.h))
        check(st)
        if not dev:
            self._keep.clear()   # host data was copied to the device
        self.n, self.nA, self.nB, self.nbatch = int(n), a.size, b.size, int(nbatch)
        return self

    def close(self):
        if self.h:
            lib().ngt_destroy(self.h)
            self.h = ctypes.c_void_p()
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setters
    def set_rates(self, k):
        if _is_torch_cuda(k):
            import torch
            if k.dtype != torch.float64:
                raise ValueError("device rate constants must be float64")
            k = k.contiguous()
            torch.cuda.current_stream(k.device).synchronize()   # the producer of k must be done (engine streams are non-blocking)
            self._keep = [k]
            check(lib().ngt_set_rates_dev(self.h, ctypes.c_void_p(k.data_ptr())))
        else:
            k = np.ascontiguousarray(k, np.float64)
            check(lib().ngt_set_rates(self.h, ctypes.c_void_p(k.ctypes.data)))

    def set_weights(self, ids, w):
        ids = np.ascontiguousarray(ids, np.int64)
        w = np.ascontiguousarray(w, np.float64)
        check(lib().ngt_set_weights(self.h, ids.size, ids.ctypes.data_as(_i64p), w.ctypes.data_as(_f64p)))

    # ---- compute
    def compute_rates(self):
        check(lib().ngt_compute_rates(self.h))

    def compute_rates_and_committors(self):
        check(lib().ngt_compute_rates_and_committors(self.h))

    def phase_one(self):
        check(lib().ngt_phase_one(self.h))

    def phase_two(self):
        check(lib().ngt_phase_two(self.h))

    def phase_two_shard(self, rank, nranks):
        check(lib().ngt_phase_two_shard(self.h, int(rank), int(nranks)))

    # ---- results
    def rates(self):
        out = np.zeros((self.nbatch, 4))
        if self.nbatch == 1:
            check(lib().ngt_get_rates(self.h, out.ctypes.data_as(_f64p)))
            return out[0]
        check(lib().ngt_batch_get_rates(self.h, out.ctypes.data_as(_f64p)))
        return out

    def final(self):
        m = self.nA + self.nB
        om, tau = np.zeros(m * self.nbatch), np.zeros(m * self.nbatch)
        check(lib().ngt_get_final(self.h, om.ctypes.data_as(_f64p), tau.ctypes.data_as(_f64p)))
        if self.nbatch == 1:
            return om, tau
        return om.reshape(self.nbatch, m), tau.reshape(self.nbatch, m)

    def set_final(self, om, tau):
        om = np.ascontiguousarray(om, np.float64)
        tau = np.ascontiguousarray(tau, np.float64)
        if om.size != (self.nA + self.nB) * self.nbatch or tau.size != om.size:
            raise ValueError("set_final needs nbatch x (|A|+|B|) values of 1-P'_aa and of tau'_a")
        check(lib().ngt_set_final(self.h, om.ctypes.data_as(_f64p), tau.ctypes.data_as(_f64p)))

    def reduced(self):
        m = self.nA + self.nB
        P, tau = np.zeros((m, m)), np.zeros(m)
        check(lib().ngt_get_reduced(self.h, P.ctypes.data_as(_f64p), tau.ctypes.data_as(_f64p)))
        return P, tau

    def elimination_order(self):
        """node ids in elimination order (intermediates, then A, then B)"""
        cnt = ctypes.c_int64()
        check(lib().ngt_get_elimination_order(self.h, None, 0, ctypes.byref(cnt)))
        out = np.empty(cnt.value, np.int64)
        check(lib().ngt_get_elimination_order(self.h, out.ctypes.data_as(_i64p), out.size, ctypes.byref(cnt)))
        return out

    def initial_tau(self):
        t = np.zeros(self.n)
        check(lib().ngt_get_initial_tau(self.h, t.ctypes.data_as(_f64p)))
        return t

    def committors(self):
        q = np.zeros(self.n)
        check(lib().ngt_get_committors(self.h, q.ctypes.data_as(_f64p)))
        return q

    def mfpt(self):
        t = np.zeros(self.n)
        check(lib().ngt_get_mfpt(self.h, t.ctypes.data_as(_f64p)))
        return t

    def stats(self):
        s = Stats()
        check(lib().ngt_get_stats(self.h, ctypes.byref(s)))
        return s.as_dict()

    def timed_steps(self, steps):
        """device time (ms, CUDA events on the engine's stream) of `steps` x (ingest + phase 1 + phase 2)"""
        ms = ctypes.c_double()
        check(lib().ngt_timed_steps(self.h, int(steps), ctypes.byref(ms)))
        return ms.value

    def profile_schur(self):
        """(ms, flops, launches) of the Schur-update kernel over one phase-1 pass"""
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        check(lib().ngt_profile_schur(self.h, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)))
        return ms.value, fl.value, n.value

    # ---- one dense network over several GPUs (see include/ngt_b200.h) ------------------------------------
    def _ptr_call(self, fn, *args):
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        check(fn(self.h, *args, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def dist_begin(self, rank, world):
        check(lib().ngt_dist_begin(self.h, int(rank), int(world)))

    def dist_info(self):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        check(lib().ngt_dist_info(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def dist_rowsum(self, J):
        return self._ptr_call(lib().ngt_dist_rowsum, int(J))

    def dist_factor(self, J):
        check(lib().ngt_dist_factor(self.h, int(J)))

    def dist_panel(self, J):
        return self._ptr_call(lib().ngt_dist_panel, int(J))

    def dist_apply(self, J):
        check(lib().ngt_dist_apply(self.h, int(J)))

    def dist_root(self):
        return self._ptr_call(lib().ngt_dist_root)

    def dist_finish(self):
        check(lib().ngt_dist_finish(self.h))

    def final_dev(self):
        """device pointers (1-P'_aa, tau'_a) and their common length, for an in-place NCCL all-reduce"""
        a, b, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        check(lib().ngt_final_dev(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)))
        return a.value, b.value, n.value

    def adopt_final_dev(self):
        check(lib().ngt_adopt_final_dev(self.h))

    def dist_phase_one(self, comm):
        """phase 1 of one dense network over the ranks of `comm` (C++ block loop, NCCL on the engine's streams);
        returns the device time in ms"""
        ms = ctypes.c_double()
        check(lib().ngt_dist_phase_one(self.h, comm.c, ctypes.byref(ms)))
        return ms.value

    def phase_two_sharded(self, comm, broadcast_root=False):
        """this rank's share of phase 2 + in-place NCCL all-reduce of the results; returns the device time in ms"""
        ms = ctypes.c_double()
        check(lib().ngt_phase_two_sharded(self.h, comm.c, 1 if broadcast_root else 0, ctypes.byref(ms)))
        return ms.value

    def reduced_block_dev(self):
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        check(lib().ngt_reduced_block_dev(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

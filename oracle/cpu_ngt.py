"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

ctypes front end for the two CPU engines under oracle/:

  kind="port"       oracle/libngt_oracle.so   our C++ restatement (oracle/ngt_oracle.cpp)
  kind="reference"  oracle/_ref/libngt_ref.so the UNMODIFIED reference engine
                                              (reference source/kmc_rates/ngt.hpp) behind
                                              oracle/ref_shim.cpp

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  kmc_rates_b200/ never does.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": (os.path.join(_HERE, "libngt_oracle.so"), "ngto_"),
    "reference": (os.path.join(_HERE, "_ref", "libngt_ref.so"), "ngtr_"),
}
_LIBS = {}

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def available(kind):
    return os.path.exists(_PATHS[kind][0])


def best_kind():
    """'reference' when the compiled reference engine is present, else 'port'."""
    return "reference" if available("reference") else "port"


def _lib(kind):
    if kind not in _LIBS:
        path, pre = _PATHS[kind]
        lib = ctypes.CDLL(path)
        f = getattr(lib, pre + "create")
        f.restype = ctypes.c_void_p
        f.argtypes = [ctypes.c_int64, ctypes.c_int64, _i64p, _i64p, _f64p, ctypes.c_int64, _i64p, ctypes.c_int64, _i64p]
        getattr(lib, pre + "destroy").argtypes = [ctypes.c_void_p]
        getattr(lib, pre + "set_weights").argtypes = [ctypes.c_void_p, ctypes.c_int64, _i64p, _f64p]
        for name in ("compute_rates", "phase_one", "compute_rates_and_committors"):
            g = getattr(lib, pre + name)
            g.argtypes = [ctypes.c_void_p]
            g.restype = ctypes.c_int
        getattr(lib, pre + "get_rates").argtypes = [ctypes.c_void_p, _f64p]
        getattr(lib, pre + "get_final").argtypes = [ctypes.c_void_p, _f64p, _f64p]
        getattr(lib, pre + "get_committors").argtypes = [ctypes.c_void_p, _f64p]
        getattr(lib, pre + "get_reduced").argtypes = [ctypes.c_void_p, _f64p, _f64p]
        _LIBS[kind] = (lib, pre)
    return _LIBS[kind]


def _p(a, t):
    return a.ctypes.data_as(t)


class CpuNGT(object):
    """Same call sequence as the reference's Cython NGT (reference kmc_rates/ngt.pyx:65-154), over
    dense int ids and COO arrays."""

    def __init__(self, n, src, dst, k, A, B, weights=None, kind="port"):
        self.lib, self.pre = _lib(kind)
        self.n = int(n)
        self.src = np.ascontiguousarray(src, np.int64)
        self.dst = np.ascontiguousarray(dst, np.int64)
        self.k = np.ascontiguousarray(k, np.float64)
        self.A = np.ascontiguousarray(list(A), np.int64)
        self.B = np.ascontiguousarray(list(B), np.int64)
        self.h = self._f("create")(self.n, self.src.size, _p(self.src, _i64p), _p(self.dst, _i64p), _p(self.k, _f64p),
                                   self.A.size, _p(self.A, _i64p), self.B.size, _p(self.B, _i64p))
        if not self.h:
            raise ValueError("oracle rejected the input")
        if weights is not None:
            ids = np.ascontiguousarray(list(weights.keys()), np.int64)
            w = np.ascontiguousarray([weights[i] for i in weights.keys()], np.float64)
            self._f("set_weights")(self.h, ids.size, _p(ids, _i64p), _p(w, _f64p))

    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def __del__(self):
        if getattr(self, "h", None):
            self._f("destroy")(self.h)
            self.h = None

    def compute_rates(self):
        if self._f("compute_rates")(self.h):
            raise RuntimeError("oracle compute_rates failed")

    def phase_one(self):
        if self._f("phase_one")(self.h):
            raise RuntimeError("oracle phase_one failed")

    def compute_rates_and_committors(self):
        if self._f("compute_rates_and_committors")(self.h):
            raise RuntimeError("oracle compute_rates_and_committors failed")

    def rates(self):
        """(k_AB, k_BA, k_AB^SS, k_BA^SS)"""
        out = np.zeros(4)
        self._f("get_rates")(self.h, _p(out, _f64p))
        return out

    def final(self):
        """(1-P'_xx, tau'_x) over A then B in the order passed."""
        m = self.A.size + self.B.size
        om, tau = np.zeros(m), np.zeros(m)
        self._f("get_final")(self.h, _p(om, _f64p), _p(tau, _f64p))
        return om, tau

    def committors(self):
        q = np.zeros(self.n)
        self._f("get_committors")(self.h, _p(q, _f64p))
        return q

    def reduced(self):
        """dense P over (A then B) and tau of the phase-1-reduced graph."""
        m = self.A.size + self.B.size
        P, tau = np.zeros((m, m)), np.zeros(m)
        self._f("get_reduced")(self.h, _p(P, _f64p), _p(tau, _f64p))
        return P, tau

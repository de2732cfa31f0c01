"""`NGT`: drop-in for the reference's Cython class (reference kmc_rates/ngt.pyx:31-161), running on B200.

Same constructor, methods and error behaviour; the label -> dense-id mapping stays on the host as in
ngt.pyx:68-73, everything numeric goes through the C ABI (kmc_rates_b200/_capi.py) to the CUDA kernels.
"""
import time

import numpy as np

from . import _capi


def _label_maps(rate_constants):
    """ids in first-appearance order (the reference uses Python-2 set order, ngt.pyx:68-73; any order is valid)."""
    node2id = {}
    for (u, v) in rate_constants.keys():
        if u not in node2id:
            node2id[u] = len(node2id)
        if v not in node2id:
            node2id[v] = len(node2id)
    return node2id


def rates_to_coo(rate_constants, node2id):
    nnz = len(rate_constants)
    src = np.empty(nnz, np.int64)
    dst = np.empty(nnz, np.int64)
    k = np.empty(nnz, np.float64)
    for i, ((u, v), r) in enumerate(rate_constants.items()):
        src[i] = node2id[u]
        dst[i] = node2id[v]
        k[i] = r
    return src, dst, k


class BaseNGT(object):
    """Transition rates between two groups of nodes by graph transformation, computed on the GPU.

    rate_constants : dict mapping (u, v) -> rate constant of the move u -> v (any hashable node labels)
    A, B           : the reactant and the product group; the results are k_AB and k_BA
    weights        : optional dict node -> equilibrium occupation probability, used to weight the average of
                     inverse mean first-passage times over the members of A (and of B)
    debug          : accepted for compatibility with the reference signature
    device, gemm_path, leaf_size : B200-side options (CUDA ordinal; 0 = DMMA tensor cores, 1 = scalar DFMA
                     cross-check path; nested-dissection leaf size)

    k_AB is the inverse mean first-passage time from a to B averaged over a in A (Wales, J. Chem. Phys. 130,
    204111 (2009)); constructor and methods mirror reference kmc_rates/ngt.pyx:31-154.
    """
    time_solve = 0.

    def __init__(self, rate_constants, A, B, debug=False, weights=None, device=0, gemm_path=0, leaf_size=0):
        self.node2id = _label_maps(rate_constants)
        self.node_list = list(self.node2id.keys())
        src, dst, k = rates_to_coo(rate_constants, self.node2id)
        # unknown label -> KeyError, as ngt.pyx:84-90
        self._A = [self.node2id[u] for u in A]
        self._B = [self.node2id[u] for u in B]
        self.debug = debug
        self._h = _capi.Handle.from_coo(len(self.node_list), src, dst, k, self._A, self._B,
                                        device=device, gemm_path=gemm_path, leaf_size=leaf_size)
        if weights is not None:
            ids, w = [], []
            for u, p in weights.items():
                if u in self.node2id:   # unknown labels silently ignored, ngt.pyx:99-103
                    ids.append(self.node2id[u])
                    w.append(p)
            self._h.set_weights(ids, w)

    # ---- array-native constructors for networks too large for a Python dict -----------------------------
    @classmethod
    def _blank(cls, n, A, B):
        self = cls.__new__(cls)
        self.node2id = None   # labels are the dense ids
        self.node_list = None
        self._A, self._B = [int(a) for a in A], [int(b) for b in B]
        self.debug = False
        self._n = int(n)
        return self

    @classmethod
    def from_coo(cls, n, src, dst, k, A, B, weights=None, nbatch=1, **opt):
        """rates as flat arrays over dense ids 0..n-1; k may be (nbatch, nnz) for networks sharing a topology"""
        self = cls._blank(n, A, B)
        self._h = _capi.Handle.from_coo(n, src, dst, k, self._A, self._B, nbatch=nbatch, **opt)
        if weights is not None:
            self._h.set_weights(list(weights.keys()), list(weights.values()))
        return self

    @classmethod
    def from_dense(cls, K, A, B, weights=None, **opt):
        """complete network from an (n,n) or (nbatch,n,n) rate matrix: numpy (host) or torch CUDA tensor (zero copy)"""
        n = K.shape[-1]
        self = cls._blank(n, A, B)
        self._h = _capi.Handle.from_dense(K, self._A, self._B, **opt)
        if weights is not None:
            self._h.set_weights(list(weights.keys()), list(weights.values()))
        return self

    # ---- the reference's methods ---------------------------------------------------------------------
    def compute_rates(self):
        """compute the rates from A->B and B->A"""
        t0 = time.perf_counter()
        self._h.compute_rates()
        self.time_solve = time.perf_counter() - t0

    def compute_rates_and_committors(self):
        """rates in both directions plus the committor of every intermediate node

        The reference re-runs a full elimination per intermediate (ngt.hpp:553-608, "much slower"); here it is one
        back-substitution over the factors phase 1 leaves behind.
        """
        t0 = time.perf_counter()
        self._h.compute_rates_and_committors()
        self.time_solve = time.perf_counter() - t0

    def get_rate_AB(self):
        """return the rate from A->B"""
        return self._rates()[0]

    def get_rate_BA(self):
        """return the rate from B->A"""
        return self._rates()[1]

    def get_rate_AB_SS(self):
        """return the steady state rate from A->B"""
        return self._rates()[2]

    def get_rate_BA_SS(self):
        """return the steady state rate from B->A"""
        return self._rates()[3]

    def _rates(self):
        r = self._h.rates()
        return r if r.ndim == 1 else r.T   # batched: each getter returns an array over the networks

    def get_committors(self):
        """return a dictionary of the committor probabilities"""
        q = self._h.committors()
        if self.node2id is None:
            return {i: float(x) for i, x in enumerate(q)}
        return {node: float(q[nid]) for node, nid in self.node2id.items()}

    # ---- extras --------------------------------------------------------------------------------------
    def get_committor_array(self):
        return self._h.committors()

    def get_mfpt_array(self):
        """mean first-passage time from every node to A ∪ B"""
        return self._h.mfpt()

    def get_final(self):
        """(1-P'_xx, tau'_x) for A then B: a->B committor and tau, whose ratio is T_aB (reference README.rst:103-112)"""
        return self._h.final()

    def stats(self):
        return self._h.stats()


class NGT(BaseNGT):
    pass

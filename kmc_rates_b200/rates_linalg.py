"""`TwoStateRates`: GPU twin of the reference's linear-algebra class (reference kmc_rates/rates_linalg.py:389-477).

The reference solves  K_II t = -1  (mean first-passage times into B) and  K_II q = -k_IB  (committors) with scipy's
sparse LU.  Eliminating nodes with the NGT engine IS that factorisation in subtraction-free (GTH) form, so both
solves are one elimination + one back-substitution on the GPU (`ngt_get_mfpt`, `ngt_get_committors`):

  * mean first-passage times: a handle with A = {} and B absorbing -> t_x for every node x;
  * committors: a handle with (A, B) -> q_x.

Same constructor, methods, attributes (`mfptimes`, `committor_dict`, `sum_out_rates`) and formulas as the reference.
"""
from collections import defaultdict

import numpy as np

from . import _capi
from .ngt import _label_maps, rates_to_coo


class LinalgError(Exception):
    """raised when a solve returns unphysical values"""


def reduce_rates(rates, B, A=None, device=0):
    """drop every rate whose tail is not connected to B (reference rates_linalg.py:13-41)"""
    B = set(B)
    if A is not None:
        A = set(A)
        if A.intersection(B):
            raise Exception("A and B share", len(A.intersection(B)), "nodes")
    labels = _label_maps(rates)
    n = len(labels)
    src, dst, _k = rates_to_coo(rates, labels)
    _nc, comp = _capi.connected_components(n, src, dst, device=device)      # union-find kernel
    good = comp[labels[next(iter(B))]]
    connected = {u for u, i in labels.items() if comp[i] == good}
    if len(connected) != n:
        rates = dict((uv, rate) for uv, rate in rates.items() if uv[0] in connected)
        if B - connected:
            raise Exception("the nodes in B are not all connected")
        if A is not None and A - connected:
            raise Exception("the A nodes are not all connected to the B nodes")
    return rates


def compute_sum_out_rates(rates):
    """sum of outgoing rates per node, added smallest first like the reference (rates_linalg.py:43-59)"""
    lists = defaultdict(list)
    for (u, _v), rate in rates.items():
        lists[u].append(rate)
    return {u: sum(sorted(r)) for u, r in lists.items()}


class TwoStateRates(object):
    """compute committors and several different rates between two groups"""

    def __init__(self, rate_constants, A, B, weights=None, check_rates=True, device=0):
        if check_rates:
            self.rate_constants = reduce_rates(rate_constants, B, A=A, device=device)
        else:
            self.rate_constants = rate_constants
        self.A = A
        self.B = B
        self.weights = weights
        if self.weights is None:
            self.weights = dict([(a, 1.) for a in self.A])
        self.sum_out_rates = compute_sum_out_rates(self.rate_constants)
        self._device = device
        self._labels = _label_maps(self.rate_constants)
        self._coo = rates_to_coo(self.rate_constants, self._labels)

    def _handle(self, A, B):
        src, dst, k = self._coo
        # The reference's linear solves treat B (and A, for committors) as absorbing and never read rates out of them
        # (reference rates_linalg.py:246-285), so a product node may be a pure sink.  The elimination engine wants a
        # waiting time for every node: kept nodes without an outgoing rate get a self rate, which no result depends on.
        has_out = np.zeros(len(self._labels), bool)
        has_out[src] = True
        sinks = [self._labels[x] for x in list(A) + list(B) if not has_out[self._labels[x]]]
        if sinks:
            sinks = np.asarray(sinks, np.int64)
            src = np.concatenate([src, sinks])
            dst = np.concatenate([dst, sinks])
            k = np.concatenate([k, np.ones(sinks.size)])
        return _capi.Handle.from_coo(len(self._labels), src, dst, k, [self._labels[a] for a in A],
                                     [self._labels[b] for b in B], device=self._device)

    def compute_rates(self, **_ignored):
        """compute the mean first passage times (into B) of every node (reference :458-471)"""
        h = self._handle([], self.B)
        h.phase_one()
        t = h.mfpt()
        h.close()
        inB = set(self.B)
        self.mfptimes = {u: float(t[i]) for u, i in self._labels.items() if u not in inB}
        if any(v < 0 for v in self.mfptimes.values()):
            raise LinalgError("error the mean first passage times are not all greater than zero")
        return self.mfptimes

    def compute_committors(self):
        """compute the committors (reference :473-477, :161-207)"""
        h = self._handle(self.A, self.B)
        h.phase_one()
        q = h.committors()
        h.close()
        skip = set(self.A) | set(self.B)
        self.committor_dict = {u: float(q[i]) for u, i in self._labels.items() if u not in skip}
        eps = 1e-10
        vals = np.array(list(self.committor_dict.values())) if self.committor_dict else np.zeros(0)
        if vals.size and (np.nanmin(vals) < -eps or np.nanmax(vals) > 1 + eps):
            raise LinalgError("The committors are not all between 0 and 1.  max=%.18g, min=%.18g" % (vals.max(), vals.min()))
        return self.committor_dict

    def get_rate_AB(self):
        """the inverse mean first passage time averaged over the nodes in A (reference :409-417)"""
        rate = sum((self.weights[a] / self.mfptimes[a] for a in self.A))
        norm = sum((self.weights[a] for a in self.A))
        return rate / norm

    def get_alternate_committors(self, Agroup, Bgroup):
        """probability that a ends up in B before coming back to itself or another node of A (reference :419-434)"""
        a_committors = dict([(a, 0.) for a in Agroup])
        for (u, v), rate in self.rate_constants.items():
            if u in Agroup and v not in Agroup:
                vcomm = 1. if v in Bgroup else self.committor_dict[v]
                a_committors[u] += rate * vcomm
        for a in Agroup:
            a_committors[a] /= self.sum_out_rates[a]
        return a_committors

    def get_rate_AB_SS(self):
        """the steady state rate from A to B (reference :436-447)"""
        a_committors = self.get_alternate_committors(self.A, self.B)
        rate = sum((self.weights[a] * a_committors[a] * self.sum_out_rates[a] for a in self.A))
        norm = sum((self.weights[a] for a in self.A))
        return rate / norm

    def get_committor(self, x):
        """the probability that a trajectory starting from x reaches B before A (reference :449-456)"""
        if x in self.A:
            return 0.
        elif x in self.B:
            return 1.
        return self.committor_dict[x]

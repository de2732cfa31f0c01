"""PATHSAMPLE ingest: min.data / ts.data / min.A / min.B -> rate constants and equilibrium occupations
(Py3, vectorised twin of reference kmc_rates/utils.py:6-162), plus the in.A / in.B / in.Peq / in.rates text format
the reference's drivers exchange (reference scripts/optim_graph.py:28-41, source/kmc_rates/test_ngt.cpp:29-98).

Same names and return values as the reference (`make_rates(directory, T) -> (rate_constants, Peq, rate_norm)`,
`read_minA`, `RatesFromPathsampleDB`, `log_sum2`); additionally `RatesFromPathsampleDB.coo()` hands the network to
the engine as flat arrays without going through a Python dict (a 10^6-transition-state database is ~2 s as a dict).
"""
import os

import numpy as np


def log_sum2(a, b):
    """return log( exp(a) + exp(b) )   (reference utils.py:6-11)"""
    if a > b:
        return a + np.log(1.0 + np.exp(-a + b))
    return b + np.log(1.0 + np.exp(a - b))


class RatesFromPathsampleDB(object):
    """read minima and ts data from a pathsample database and compute rate constants (reference utils.py:13-106)

    min.data columns: energy, log-product of vibrational frequencies, point-group order, ...
    ts.data columns:  energy, fvib, pgorder, min1, min2, ...            (minima are numbered from 1)

    k(u->v) through a transition state = pgorder_u / (2 pi pgorder_ts) * exp((fvib_u - fvib_ts)/2) * exp(-(E_ts - E_u)/T),
    evaluated in log space, shifted by the largest log rate (`log_rate_subtracted`, `rate_norm = exp(-max)`),
    transition states between the same pair of minima summed, self-connecting transition states dropped.
    """

    def __init__(self, mindata_file, tsdata_file, T, device=None):
        """device: CUDA ordinal -> the log rates, the shift and the merge of parallel transition states run on the GPU
        (`_capi.pathsample_rates`, csrc/pathsample.cu); None -> the vectorised numpy path below (same results)"""
        self.T = T
        self.device = device
        self.run(mindata_file, tsdata_file)

    def _log_equilibrium_occupation_probabilities(self, mindata):
        energy, fvib, pgorder = mindata[:, 0], mindata[:, 1], mindata[:, 2]
        return -energy / self.T - np.log(pgorder) - 0.5 * fvib

    def _compute_log_rates(self, m1vals, tsdata, T):
        return (np.log(m1vals[:, 2] / (2. * np.pi * tsdata[:, 2]))
                + (m1vals[:, 1] - tsdata[:, 1]) / 2.
                - (tsdata[:, 0] - m1vals[:, 0]) / T)

    def compute_rates(self, mindata, tsdata, mindices):
        m1vals = mindata[mindices, :]
        assert m1vals.shape[0] == tsdata.shape[0]
        return self._compute_log_rates(m1vals, tsdata, self.T)

    def run(self, mindataf, tsdataf):
        mindata = np.atleast_2d(np.genfromtxt(mindataf, usecols=[0, 1, 2]))
        tsdata = np.atleast_2d(np.genfromtxt(tsdataf, usecols=[0, 1, 2, 3, 4]))
        m1 = tsdata[:, 3].astype(np.int64) - 1
        m2 = tsdata[:, 4].astype(np.int64) - 1
        nmin = mindata.shape[0]
        log_Peq = self._log_equilibrium_occupation_probabilities(mindata)
        self._Peq = np.exp(log_Peq - log_Peq.max())
        self._nmin = nmin
        if self.device is not None:
            from . import _capi
            self._src, self._dst, self._k, max_log_rate = _capi.pathsample_rates(mindata[:, :3], tsdata[:, :3], m1, m2, self.T,
                                                                               device=self.device)
            self.log_rate_subtracted = max_log_rate
            self.rate_norm = np.exp(-max_log_rate)
            return
        log_k12 = self.compute_rates(mindata, tsdata, m1)
        log_k21 = self.compute_rates(mindata, tsdata, m2)
        max_log_rate = max(log_k12.max(), log_k21.max())
        log_k12 -= max_log_rate
        log_k21 -= max_log_rate
        self.log_rate_subtracted = max_log_rate
        self.rate_norm = np.exp(-max_log_rate)

        # directed entries, self-connecting transition states dropped, parallel ones summed in log space
        keep = m1 != m2
        src = np.concatenate([m1[keep], m2[keep]])
        dst = np.concatenate([m2[keep], m1[keep]])
        logk = np.concatenate([log_k12[keep], log_k21[keep]])
        key = src * nmin + dst
        order = np.argsort(key, kind="stable")
        key, logk = key[order], logk[order]
        starts = np.flatnonzero(np.concatenate([[True], key[1:] != key[:-1]]))
        merged = np.logaddexp.reduceat(logk, starts)
        self._src = (key[starts] // nmin).astype(np.int64)
        self._dst = (key[starts] % nmin).astype(np.int64)
        self._k = np.exp(merged)

    # ---- the reference's attributes (1-based pathsample ids), built on demand -------------------------------
    @property
    def rate_constants(self):
        return {(int(u) + 1, int(v) + 1): float(r) for u, v, r in zip(self._src, self._dst, self._k)}

    @property
    def Peq(self):
        return {u + 1: float(p) for u, p in enumerate(self._Peq)}

    def coo(self):
        """(n, src, dst, k, Peq): 0-based flat arrays for NGT.from_coo (minimum i of the database is id i-1)"""
        return self._nmin, self._src, self._dst, self._k, self._Peq


def read_minA(fname):
    """load data from min.A or min.B: first line = count, then whitespace-separated ids (reference utils.py:141-154)"""
    with open(fname) as fin:
        ids = []
        nminima = None
        for i, line in enumerate(fin):
            if i == 0:
                nminima = int(line.split()[0])
            else:
                ids += [int(x) for x in line.split()]
    assert nminima == len(ids)
    return ids


def make_rates(directory, T, device=None):
    """rate constants, equilibrium occupation probabilities and the rate normalisation from a pathsample database
    (reference utils.py:156-162); device: CUDA ordinal for the on-device path"""
    gen = RatesFromPathsampleDB(os.path.join(directory, "min.data"), os.path.join(directory, "ts.data"), T, device=device)
    return gen.rate_constants, gen.Peq, gen.rate_norm


def write_dat_files(rates, A, B, Peq, directory="."):
    """in.A / in.B / in.Peq / in.rates (reference scripts/optim_graph.py:28-41)"""
    with open(os.path.join(directory, "in.A"), "w") as f:
        for a in A:
            f.write("%d\n" % a)
    with open(os.path.join(directory, "in.B"), "w") as f:
        for b in B:
            f.write("%d\n" % b)
    with open(os.path.join(directory, "in.Peq"), "w") as f:
        for u, p in Peq.items():
            f.write("%d %.16g\n" % (u, p))
    with open(os.path.join(directory, "in.rates"), "w") as f:
        for (u, v), p in rates.items():
            f.write("%d %d %.16g\n" % (u, v, p))


def read_dat_files(directory="."):
    """inverse of write_dat_files -> (rates, A, B, Peq)   (format read by reference source/kmc_rates/test_ngt.cpp:29-98)"""
    def ints(name):
        with open(os.path.join(directory, name)) as f:
            return [int(line.split()[0]) for line in f if line.strip()]
    Peq = {}
    with open(os.path.join(directory, "in.Peq")) as f:
        for line in f:
            if line.strip():
                u, p = line.split()
                Peq[int(u)] = float(p)
    rates = {}
    with open(os.path.join(directory, "in.rates")) as f:
        for line in f:
            if line.strip():
                u, v, p = line.split()
                rates[(int(u), int(v))] = float(p)
    return rates, ints("in.A"), ints("in.B"), Peq

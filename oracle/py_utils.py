"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

Loop-for-loop Python-3 restatement of the reference's PATHSAMPLE reader (reference kmc_rates/utils.py:13-106,
Python 2: print statements :74, izip :1, iteritems :66), used to check kmc_rates_b200/utils.py's vectorised version.
Parity status: PINNED.  The reference ships no fixture for this path (no min.data / ts.data in the repository and no
test of utils.py), so tests/golden/pathsample_golden.json holds outputs of the reference's own utils.py, executed
through the in-memory Python-2 shim of oracle/make_golden_utils.py; tests/test_utils.py checks this restatement (and
the product's host and device paths) against them.
"""
import numpy as np


def log_sum2(a, b):
    if a > b:
        return a + np.log(1.0 + np.exp(-a + b))
    return b + np.log(1.0 + np.exp(a - b))


def rates_from_pathsample(mindataf, tsdataf, T):
    mindata = np.atleast_2d(np.genfromtxt(mindataf, usecols=[0, 1, 2]))
    tsdata = np.atleast_2d(np.genfromtxt(tsdataf, usecols=[0, 1, 2, 3, 4]))
    m1 = tsdata[:, 3].astype(int) - 1
    m2 = tsdata[:, 4].astype(int) - 1

    def log_rates(mind):
        mv = mindata[mind, :]
        return (np.log(mv[:, 2] / (2. * np.pi * tsdata[:, 2])) + (mv[:, 1] - tsdata[:, 1]) / 2.
                - (tsdata[:, 0] - mv[:, 0]) / T)

    k12, k21 = log_rates(m1), log_rates(m2)
    mx = max(k12.max(), k21.max())
    k12 -= mx
    k21 -= mx
    lr = {}
    for u, v, a, b in zip(m1, m2, k12, k21):          # reference utils.py:56-65
        if u == v:
            continue
        if (u, v) in lr:
            lr[(u, v)] = log_sum2(lr[(u, v)], a)
            lr[(v, u)] = log_sum2(lr[(v, u)], b)
        else:
            lr[(u, v)] = a
            lr[(v, u)] = b
    rates = {(int(u) + 1, int(v) + 1): float(np.exp(k)) for (u, v), k in lr.items()}
    logp = -mindata[:, 0] / T - np.log(mindata[:, 2]) - 0.5 * mindata[:, 1]
    peq = np.exp(logp - logp.max())
    return rates, {i + 1: float(p) for i, p in enumerate(peq)}, float(np.exp(-mx))

"""Rates from a PATHSAMPLE database with NGT on B200 (Py3 twin of reference scripts/pathsample_rates.py:24-125, which
used the linear solver): reads min.data, ts.data, min.A, min.B from -d, prints k_AB, k_BA and the steady-state rates.

    python scripts/pathsample_rates.py -d <directory> -T <temperature>
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmc_rates_b200 import NGT, utils


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("-d", type=str, default=".", help="directory with the min.data, ts.data, min.A, min.B files")
    ap.add_argument("-T", type=float, default=1.0, help="temperature")
    args = ap.parse_args()
    t0 = time.perf_counter()
    gen = utils.RatesFromPathsampleDB(os.path.join(args.d, "min.data"), os.path.join(args.d, "ts.data"), args.T)
    A = [a - 1 for a in utils.read_minA(os.path.join(args.d, "min.A"))]
    B = [b - 1 for b in utils.read_minA(os.path.join(args.d, "min.B"))]
    n, src, dst, k, peq = gen.coo()
    ngt = NGT.from_coo(n, src, dst, k, A, B, weights={i: float(p) for i, p in enumerate(peq)})
    ngt.compute_rates()
    norm = gen.rate_norm
    print("temperature", args.T, " minima", n, " directed rates", src.size)
    print("k_AB    %.16g" % (ngt.get_rate_AB() / norm))
    print("k_BA    %.16g" % (ngt.get_rate_BA() / norm))
    print("k_AB_SS %.16g" % (ngt.get_rate_AB_SS() / norm))
    print("k_BA_SS %.16g" % (ngt.get_rate_BA_SS() / norm))
    print("total time %.3f s (NGT solve %.4f s)" % (time.perf_counter() - t0, ngt.time_solve))


if __name__ == "__main__":
    main()

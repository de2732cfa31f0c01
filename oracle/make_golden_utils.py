"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

Golden vectors for the PATHSAMPLE ingest path, produced by RUNNING THE REFERENCE'S OWN CODE:
/root/reference/kmc_rates/utils.py is Python 2 (print statements, izip, iteritems, map/filter lists), so its source
is read from where it lies, passed through a token-level shim IN MEMORY (nothing of it is copied into this
repository) and executed; `make_rates`, `read_minA` and `log_sum2` of that module then process small synthetic
min.data / ts.data / min.A files.  Inputs (the file texts) and outputs go to tests/golden/pathsample_golden.json.

    python oracle/make_golden_utils.py            # needs /root/reference; run in the build container only

The shim is mechanical and behaviour-preserving:
    print <expr-list>        -> print(<expr-list>)
    from itertools import izip / izip(  -> zip(
    .iteritems() / .iterkeys() -> .items() / .keys()
    filter(f, xs) / map(f, xs) used as lists -> list(...)
"""
import json
import os
import re
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/kmc_rates/utils.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pathsample_golden.json")


def load_reference_utils():
    src = open(REF).read()
    out = []
    for line in src.splitlines():
        m = re.match(r"^(\s*)print\s+(.*)$", line)
        if m and not m.group(2).startswith("("):
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        out.append(line)
    src = "\n".join(out)
    src = src.replace("from itertools import izip", "izip = zip")
    src = src.replace(".iteritems()", ".items()").replace(".iterkeys()", ".keys()")
    src = src.replace("Aclist = filter(", "Aclist = list(filter(").replace("len(c)>0, Aclist)", "len(c)>0, Aclist))")
    src = src.replace("ids += map(int, sline)", "ids += list(map(int, sline))")
    mod = types.ModuleType("reference_utils")
    exec(compile(src, REF, "exec"), mod.__dict__)
    return mod


def synthetic_database(nmin, nts, seed):
    """min.data / ts.data texts in PATHSAMPLE's column layout (energy fvib pgorder [min1 min2] I1 I2 I3), with
    repeated transition states between the same pair of minima, both orientations, and self-connecting ones"""
    rng = np.random.default_rng(seed)
    E = rng.uniform(-10.0, -5.0, nmin)
    fv = rng.uniform(80.0, 120.0, nmin)
    pg = rng.choice([1, 2, 3, 4, 6, 12], nmin)
    mind = "\n".join("%.10f %.10f %d %.4f %.4f %.4f" % (E[i], fv[i], pg[i], *rng.uniform(1, 9, 3)) for i in range(nmin)) + "\n"
    rows = []
    for t in range(nts):
        if t > 0 and rng.random() < 0.25:
            a, b = rows[int(rng.integers(0, len(rows)))][3:5]      # another TS between an existing pair
            if rng.random() < 0.5:
                a, b = b, a
        elif rng.random() < 0.05:
            a = b = int(rng.integers(1, nmin + 1))                  # self-connecting TS (dropped by the reference)
        else:
            a, b = (int(x) for x in rng.choice(nmin, 2, replace=False) + 1)
        ets = max(E[a - 1], E[b - 1]) + rng.uniform(0.05, 2.0)
        rows.append((ets, rng.uniform(78.0, 118.0), int(rng.choice([1, 2])), a, b))
    tsd = "\n".join("%.10f %.10f %d %d %d %.4f %.4f %.4f" % (r[0], r[1], r[2], r[3], r[4], *rng.uniform(1, 9, 3)) for r in rows) + "\n"
    ids = sorted(int(x) for x in rng.choice(nmin, min(5, nmin), replace=False) + 1)
    mina = "%d\n%s\n" % (len(ids), "\n".join(" ".join(str(x) for x in ids[i:i + 3]) for i in range(0, len(ids), 3)))
    return mind, tsd, mina


def main():
    ref = load_reference_utils()
    cases = []
    for name, nmin, nts, T, seed in [("tiny", 4, 6, 1.0, 1), ("small", 12, 40, 0.5, 2), ("cold", 30, 120, 0.1, 3),
                                     ("wide", 60, 400, 2.0, 4)]:
        mind, tsd, mina = synthetic_database(nmin, nts, seed)
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "min.data"), "w").write(mind)
            open(os.path.join(d, "ts.data"), "w").write(tsd)
            open(os.path.join(d, "min.A"), "w").write(mina)
            _stdout = sys.stdout
            sys.stdout = open(os.devnull, "w")
            try:
                rates, peq, norm = ref.make_rates(d, T)
                ids = ref.read_minA(os.path.join(d, "min.A"))
            finally:
                sys.stdout = _stdout
        cases.append({
            "name": name, "T": T, "min_data": mind, "ts_data": tsd, "min_A": mina,
            "rate_constants": [[int(u), int(v), repr(float(k))] for (u, v), k in sorted(rates.items())],
            "Peq": [[int(u), repr(float(p))] for u, p in sorted(peq.items())],
            "rate_norm": repr(float(norm)), "min_A_ids": [int(x) for x in ids],
        })
    pairs = [(-3.0, 1.5), (0.0, 0.0), (700.0, -700.0), (-50.25, -50.5), (1e-3, 2e-3)]
    golden = {"source": "outputs of /root/reference/kmc_rates/utils.py (make_rates, read_minA, log_sum2) executed through "
                        "oracle/make_golden_utils.py's in-memory Python-2 shim; floats as repr() strings",
              "log_sum2": [[repr(a), repr(b), repr(float(ref.log_sum2(a, b)))] for a, b in pairs],
              "cases": cases}
    with open(OUT, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", OUT, "with", len(cases), "cases,", sum(len(c["rate_constants"]) for c in cases), "rate constants")


if __name__ == "__main__":
    main()

"""ORACLE — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/ngt_golden.json by running the UNMODIFIED reference engine
(oracle/_ref/libngt_ref.so = reference source/kmc_rates/ngt.hpp behind oracle/ref_shim.cpp) on a
fixed list of cases.  Run in the build container, where /root/reference exists:

    make -C oracle && python -m oracle.make_golden

The GPU box has no /root/reference; there the committed JSON is the reference's voice.
Inputs of small cases are stored inline; larger ones are regenerated from
kmc_rates_b200.synth with the recorded parameters and verified by a SHA-256 of the COO arrays.
"""
import hashlib
import json
import os

import numpy as np

from kmc_rates_b200 import synth
from oracle.cpu_ngt import CpuNGT

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "ngt_golden.json")


def digest(src, dst, k):
    h = hashlib.sha256()
    for a in (np.ascontiguousarray(src, np.int64), np.ascontiguousarray(dst, np.int64), np.ascontiguousarray(k, np.float64)):
        h.update(a.tobytes())
    return h.hexdigest()


def build_input(spec):
    """spec -> (n, src, dst, k, A, B).  Shared with tests/conftest.py via import."""
    g = spec["gen"]
    if g == "inline":
        e = spec["edges"]
        src = np.array([x[0] for x in e], np.int64)
        dst = np.array([x[1] for x in e], np.int64)
        k = np.array([x[2] for x in e], np.float64)
        return spec["n"], src, dst, k, spec["A"], spec["B"]
    if g == "complete_deterministic":
        src, dst, k = synth.complete_deterministic(spec["n"])
        return spec["n"], src, dst, k, spec["A"], spec["B"]
    if g == "complete_random":
        src, dst, k = synth.complete_random(spec["n"], seed=spec["seed"])
        return spec["n"], src, dst, k, spec["A"], spec["B"]
    if g == "random_connected_graph":
        rates = synth.random_connected_graph(spec["nnodes"], spec["nedges"], spec["A"] + spec["B"], seed=spec["seed"])
        src, dst, k = synth.from_dict(rates)
        return spec["nnodes"], src, dst, k, spec["A"], spec["B"]
    if g == "landscape_2d":
        n, src, dst, k, A, B = synth.landscape_2d(spec["side"], T=spec["T"], seed=spec["seed"])
        return n, src, dst, k, A, B
    if g == "landscape_groups":
        return synth.landscape_groups(spec["side"], spec["nA"], spec["nB"], T=spec["T"], seed=spec["seed"])
    raise ValueError(g)


def cases():
    c = []
    c.append(dict(name="ngt3_self_rates", gen="inline", n=3, A=[0], B=[1], committors=True,
                  edges=[[i, j, 1.0] for i in range(3) for j in range(3)],
                  cite="cpp_tests/source/test_ngt.cpp:9-46"))
    c.append(dict(name="ngt3", gen="inline", n=3, A=[0], B=[1], committors=True,
                  edges=[[i, j, 1.0] for i in range(3) for j in range(3) if i != j],
                  cite="kmc_rates/tests/test_cpp_ngt.py:10-36"))
    c.append(dict(name="ngt10", gen="complete_deterministic", n=10, A=[0, 1, 2], B=[3, 4, 5], committors=True,
                  cite="cpp_tests/source/test_ngt.cpp:48-93"))
    c.append(dict(name="ngt10_weighted", gen="complete_deterministic", n=10, A=[0, 1, 2], B=[3, 4, 5],
                  weights={"0": 1.0, "1": 1.0, "2": 5e3, "3": 1.0, "4": 1.0, "5": 2.0},
                  cite="SURVEY.md §8(c) probe; scripts/optim_graph.py:78-84 passes weights=Peq"))
    c.append(dict(name="tuple3", gen="inline", n=3, A=[0], B=[1], committors=True,
                  edges=[[0, 1, 1.0], [1, 0, 2.0], [0, 2, 1.0], [2, 0, 1.0], [1, 2, 1.0], [2, 1, 1.0]],
                  cite="SURVEY.md §8(c) probe"))
    for i, (A, B) in enumerate([([0], [1]), ([0, 1, 2], [3]), ([0, 1, 2], [3, 4, 5, 6])]):
        c.append(dict(name="random20_%d" % i, gen="random_connected_graph", nnodes=20, nedges=20, seed=i, A=A, B=B,
                      committors=True, cite="kmc_rates/tests/test_cpp_ngt.py:66-106"))
    c.append(dict(name="random10_group8", gen="random_connected_graph", nnodes=10, nedges=20, seed=7,
                  A=list(range(8)), B=[9], committors=True, cite="kmc_rates/tests/test_kmc.py:93-96"))
    c.append(dict(name="c1_complete100", gen="complete_random", n=100, seed=0, A=[0], B=[1], committors=False,
                  cite="BASELINE config 1; examples/example_random_graph.py:7-21"))
    c.append(dict(name="complete6_example", gen="complete_random", n=6, seed=1, A=[0, 1], B=[4, 5], committors=True,
                  cite="examples/example_random_graph.py:25-47"))
    c.append(dict(name="complete64_g8", gen="complete_random", n=64, seed=2, A=list(range(8)), B=list(range(56, 64)),
                  committors=False, cite="multi-state phase 2"))
    c.append(dict(name="complete256", gen="complete_random", n=256, seed=3, A=[0], B=[255], committors=False,
                  cite="BASELINE config 5 unit"))
    c.append(dict(name="landscape12_T0.3", gen="landscape_2d", side=12, T=0.3, seed=0, committors=True, cite="C2 generator"))
    c.append(dict(name="landscape20_T0.3", gen="landscape_2d", side=20, T=0.3, seed=0, committors=False, cite="C2 generator",
                  truth_mp=True))
    # rtol: the reference's own run-to-run spread on ill-conditioned inputs is ~2e-10 (pointer-ordered
    # elimination + `1. - Pxx` below the 0.99 threshold); these cases carry an extended-precision truth.
    c.append(dict(name="landscape20_T0.03", gen="landscape_2d", side=20, T=0.03, seed=0, committors=False,
                  cite="C2 generator, ill-conditioned", rtol=5e-9, truth_mp=True))
    c.append(dict(name="landscape50_T0.3", gen="landscape_2d", side=50, T=0.3, seed=1, committors=False, cite="C2 generator"))
    c.append(dict(name="landscape50_T0.05", gen="landscape_2d", side=50, T=0.05, seed=1, committors=False,
                  cite="C2 generator, ill-conditioned", rtol=5e-9, truth_mp=True))
    c.append(dict(name="groups24_6x6", gen="landscape_groups", side=24, nA=6, nB=6, T=0.3, seed=0, committors=False,
                  cite="BASELINE config 4 shape"))
    return c


def truth_mp(n, src, dst, k, a, b, dps=60):
    """Extended-precision (mpmath) dense NGT for |A|=|B|=1: the exact answer an ill-conditioned case is
    judged against, because the reference itself wanders by ~2e-10 run to run there (its elimination
    order follows pointer values, reference source/kmc_rates/graph.hpp:67, ngt.hpp:192-200)."""
    import mpmath
    mpmath.mp.dps = dps
    mpf = mpmath.mpf
    sumk = [mpf(0)] * n
    for u, r in zip(src, k):
        sumk[u] += mpf(float(r))
    tau = [1 / s for s in sumk]
    tau0 = list(tau)
    P = [dict() for _ in range(n)]
    for u, v, r in zip(src, dst, k):
        P[u][int(v)] = mpf(float(r)) * tau[u]
    preds = [set() for _ in range(n)]
    for u in range(n):
        for v in P[u]:
            preds[v].add(u)
    order = [x for x in range(n) if x != a and x != b]
    alive = set(range(n))
    import heapq
    while order:
        x = min(order, key=lambda y: len(P[y]) + len(preds[y]))
        order.remove(x)
        row = {v: p for v, p in P[x].items() if v != x}
        om = sum(row.values())
        for u in list(preds[x]):
            if u == x:
                continue
            f = P[u].pop(x) / om
            tau[u] += f * tau[x]
            for v, p in row.items():
                if v not in P[u]:
                    P[u][v] = mpf(0)
                    preds[v].add(u)
                P[u][v] += f * p
        for v in row:
            preds[v].discard(x)
        alive.discard(x)
    res = {}
    res["kAB"] = float(P[a][b] / tau[a])
    res["kBA"] = float(P[b][a] / tau[b])
    res["kAB_SS"] = float(P[a][b] / tau0[a])
    res["kBA_SS"] = float(P[b][a] / tau0[b])
    return res


def main():
    out = []
    for spec in cases():
        n, src, dst, k, A, B = build_input(spec)
        w = spec.get("weights")
        wd = {int(a): float(b) for a, b in w.items()} if w else None
        ref = CpuNGT(n, src, dst, k, A, B, weights=wd, kind="reference")
        ref.phase_one()
        P, tau1 = ref.reduced()
        ref = CpuNGT(n, src, dst, k, A, B, weights=wd, kind="reference")
        if spec.get("committors"):
            ref.compute_rates_and_committors()
            q = ref.committors()
        else:
            ref.compute_rates()
            q = None
        rates = ref.rates()
        om, tau = ref.final()
        rec = dict(spec)
        rec["resolved_A"], rec["resolved_B"] = [int(a) for a in A], [int(b) for b in B]
        rec["n_nodes"], rec["n_rates"] = int(n), int(src.size)
        rec["sha256"] = digest(src, dst, k)
        rec["expect"] = dict(
            kAB=float(rates[0]), kBA=float(rates[1]), kAB_SS=float(rates[2]), kBA_SS=float(rates[3]),
            final_omPxx=[float(x) for x in om], final_tau=[float(x) for x in tau],
            reduced_P=[[float(x) for x in row] for row in P], reduced_tau=[float(x) for x in tau1],
        )
        if spec.get("truth_mp"):
            rec["truth_mp"] = truth_mp(n, [int(x) for x in src], [int(x) for x in dst], k, int(A[0]), int(B[0]))
        if q is not None:
            rec["expect"]["committors"] = [None if np.isnan(x) else float(x) for x in q]
        out.append(rec)
        print("%-22s n=%-6d nnz=%-7d kAB=%.17g kBA=%.17g" % (spec["name"], n, src.size, rates[0], rates[1]))
    with open(OUT, "w") as f:
        json.dump(dict(generator="oracle/make_golden.py", engine="reference source/kmc_rates/ngt.hpp via oracle/ref_shim.cpp, g++ -O2",
                       cases=out), f, indent=1)
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()

"""Randomised parity sweep on the GPU: many small random networks (sparse and dense, random group sizes, optional
self rates, weights, tiny leaf sizes to force deep assembly trees) against the CPU oracle, through the C ABI."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _random_network(rng):
    n = int(rng.integers(4, 70))
    kind = rng.choice(["dense", "sparse", "ring", "tree"])
    edges = set()
    if kind == "dense":
        edges = {(i, j) for i in range(n) for j in range(n) if i != j}
    elif kind == "ring":
        for i in range(n):
            edges.add((i, (i + 1) % n)); edges.add(((i + 1) % n, i))
        for _ in range(int(rng.integers(0, n))):
            u, v = rng.integers(0, n, 2)
            if u != v:
                edges.add((int(u), int(v))); edges.add((int(v), int(u)))
    elif kind == "tree":
        for i in range(1, n):
            p = int(rng.integers(0, i))
            edges.add((i, p)); edges.add((p, i))
    else:
        for i in range(1, n):                      # spanning chain keeps it connected
            edges.add((i - 1, i)); edges.add((i, i - 1))
        for _ in range(int(rng.integers(n, 4 * n))):
            u, v = rng.integers(0, n, 2)
            if u != v:
                edges.add((int(u), int(v))); edges.add((int(v), int(u)))
    edges = sorted(edges)
    src = np.array([e[0] for e in edges]); dst = np.array([e[1] for e in edges])
    k = np.exp(rng.uniform(-6, 2, len(edges)))     # rates over ~3.5 decades
    if rng.random() < 0.3:                          # a few self rates
        loops = rng.choice(n, size=max(1, n // 5), replace=False)
        src = np.concatenate([src, loops]); dst = np.concatenate([dst, loops]); k = np.concatenate([k, rng.random(loops.size)])
    nodes = rng.permutation(n)
    nA = int(rng.integers(1, max(2, min(6, n // 2))))
    nB = int(rng.integers(1, max(2, min(6, n - nA))))
    A = [int(x) for x in nodes[:nA]]; B = [int(x) for x in nodes[nA:nA + nB]]
    w = {x: float(rng.uniform(0.1, 1.0)) for x in A + B} if rng.random() < 0.5 else None
    return n, src, dst, k, A, B, w


@pytest.mark.parametrize("seed", range(40))
def test_random_network_matches_oracle(seed):
    from kmc_rates_b200 import NGT
    from oracle.cpu_ngt import CpuNGT, best_kind
    rng = np.random.default_rng(1000 + seed)
    n, src, dst, k, A, B, w = _random_network(rng)
    leaf = int(rng.choice([0, 4, 8, 16]))
    g = NGT.from_coo(n, src, dst, k, A, B, weights=w, leaf_size=leaf, gemm_path=int(rng.choice([0, 1, 2])))
    g.compute_rates_and_committors()
    cpu = CpuNGT(n, src, dst, k, A, B, weights=w, kind=best_kind())   # the compiled reference itself when it is present
    cpu.compute_rates_and_committors()
    got = [g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()]
    assert rel_err(got, cpu.rates()) < 1e-9
    om, tau = g.get_final()
    com, ctau = cpu.final()
    assert rel_err(om, com) < 1e-9 and rel_err(tau, ctau) < 1e-9
    q, cq = g.get_committor_array(), cpu.committors()
    ok = ~np.isnan(cq)
    assert np.max(np.abs(q[ok] - cq[ok])) < 1e-9

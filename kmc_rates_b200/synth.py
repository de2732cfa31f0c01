"""Synthetic rate networks for tests and bench.py.

All generators return flat COO arrays ``(src, dst, k)`` (int64, int64, float64) over dense
node ids ``0..n-1`` plus the reactant / product id lists, because the large configurations
(10^5 nodes, 16 384-node complete graphs) must never exist as Python dicts.  ``to_dict``
turns a small case into the ``{(u, v): k}`` mapping the reference API takes
(reference kmc_rates/_kmc_rates.py:18-31).
"""
import numpy as np


def to_dict(src, dst, k):
    return {(int(u), int(v)): float(r) for u, v, r in zip(src, dst, k)}


def from_dict(rates):
    """dict over dense int ids -> COO arrays (insertion order)."""
    n = len(rates)
    src = np.empty(n, np.int64)
    dst = np.empty(n, np.int64)
    k = np.empty(n, np.float64)
    for i, ((u, v), r) in enumerate(rates.items()):
        src[i], dst[i], k[i] = u, v, r
    return src, dst, k


def complete_deterministic(n):
    """k_ij = (i+j)/(i+1): the reference's NGT10 fixture at any size
    (reference kmc_rates/tests/test_graph_transformation.py:8-14, cpp_tests/source/test_ngt.cpp:55-60)."""
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    mask = i != j
    src = i[mask].astype(np.int64)
    dst = j[mask].astype(np.int64)
    k = (src + dst).astype(np.float64) / (src + 1)
    return src, dst, k


def complete_random(n, seed=0):
    """Random complete rate graph, k_ij ~ U(0,1) (reference examples/example_random_graph.py:7-21;
    BASELINE config 1 at n=100)."""
    rng = np.random.default_rng(seed)
    K = rng.random((n, n))
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    mask = i != j
    return i[mask].astype(np.int64), j[mask].astype(np.int64), K[mask].astype(np.float64)


_HALF_OFFSETS = [(dx, dy) for dx in range(-2, 3) for dy in range(-2, 3) if (dx, dy) > (0, 0)]


def landscape_2d(side, T=0.3, seed=0, mean_ts_per_node=10.0, shuffle=True):
    """Low-dimensional synthetic energy landscape (BASELINE config 2).

    ``side*side`` minima on a 2-D lattice; a transition state joins two minima within Chebyshev
    distance 2: the 4 axis neighbours always, the other 20 with the probability that gives
    ``mean_ts_per_node`` transition states per minimum (10 -> mean degree 20).  Energies
    E_i ~ U(0,1), E_ts = max(E_i, E_j) + U(0.1, 1), k_ij = exp(-(E_ts - E_i)/T).

    Returns (n, src, dst, k, A, B): A = the minimum at lattice corner (0,0), B = the opposite corner.
    Node ids are a seeded random permutation of the lattice index when ``shuffle`` so that no
    ordering can exploit id locality.
    """
    rng = np.random.default_rng(seed)
    n = side * side
    p = (mean_ts_per_node - 2.0) / 10.0
    p = min(max(p, 0.0), 1.0)
    E = rng.random(n)
    xs, ys = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    us, vs = [], []
    for dx, dy in _HALF_OFFSETS:
        ok = (xs + dx >= 0) & (xs + dx < side) & (ys + dy >= 0) & (ys + dy < side)
        u = (xs[ok] * side + ys[ok]).ravel()
        v = ((xs[ok] + dx) * side + (ys[ok] + dy)).ravel()
        if (abs(dx) + abs(dy)) != 1:
            keep = rng.random(u.size) < p
            u, v = u[keep], v[keep]
        us.append(u)
        vs.append(v)
    u = np.concatenate(us).astype(np.int64)
    v = np.concatenate(vs).astype(np.int64)
    Ets = np.maximum(E[u], E[v]) + rng.uniform(0.1, 1.0, u.size)
    kuv = np.exp(-(Ets - E[u]) / T)
    kvu = np.exp(-(Ets - E[v]) / T)
    if shuffle:
        perm = rng.permutation(n).astype(np.int64)
    else:
        perm = np.arange(n, dtype=np.int64)
    src = np.concatenate([perm[u], perm[v]])
    dst = np.concatenate([perm[v], perm[u]])
    k = np.concatenate([kuv, kvu])
    A = [int(perm[0])]
    B = [int(perm[n - 1])]
    return n, src, dst, k, A, B


def landscape_groups(side, nA, nB, T=0.3, seed=0, mean_ts_per_node=10.0):
    """C2 generator with multi-minimum reactant / product groups (BASELINE config 4):
    A = the nA lattice sites nearest corner (0,0), B = the nB nearest the opposite corner.
    Both groups are connected through axis neighbours, as the reference requires
    (reference kmc_rates/_kmc_rates.py:390-404)."""
    n, src, dst, k, _, _ = landscape_2d(side, T=T, seed=seed, mean_ts_per_node=mean_ts_per_node, shuffle=False)
    xs, ys = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    dA = (xs + ys).ravel()
    dB = ((side - 1 - xs) + (side - 1 - ys)).ravel()
    A = [int(i) for i in np.argsort(dA, kind="stable")[:nA]]
    B = [int(i) for i in np.argsort(dB, kind="stable")[:nB]]
    return n, src, dst, k, A, B


def random_connected_graph(nnodes=20, nedges=20, node_set=(0, 1), seed=0):
    """Py3 twin of the reference's _MakeRandomGraph (reference
    kmc_rates/tests/test_graph_transformation.py:17-56): add random bidirectional edges with rates
    U(0.5,1.5) until ``node_set`` is connected and at least ``nedges`` directed rates exist.
    Returns a dict {(u,v): k}; nodes never touched by an edge do not appear."""
    rng = np.random.RandomState(seed)
    nodes = np.arange(nnodes)
    parent = list(range(nnodes))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    rates = {}

    def add():
        rng.shuffle(nodes)
        u, v = int(nodes[0]), int(nodes[1])
        r1, r2 = rng.uniform(0.5, 1.5, 2)
        rates[(u, v)] = float(r1)
        rates[(v, u)] = float(r2)
        parent[find(u)] = find(v)

    node_set = list(node_set)
    while len({find(x) for x in node_set}) > 1:
        add()
    target = min(nedges, nnodes * (nnodes - 1))
    while len(rates) < target:
        add()
    return rates

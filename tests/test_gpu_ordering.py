"""The level-structure sweeps of the nested dissection on the device (csrc/ordering.cu) must report the serial
breadth-first queue order exactly, so that the elimination order does not depend on who sweeps."""
import numpy as np
import pytest

from kmc_rates_b200 import _capi, synth
from kmc_rates_b200._capi import Handle

pytestmark = pytest.mark.gpu


def _csr(n, src, dst, rng=None):
    """symmetric, de-duplicated CSR without self loops; rows optionally shuffled (the sweep order depends on them)"""
    u = np.concatenate([src, dst])
    v = np.concatenate([dst, src])
    keep = u != v
    e = np.unique(np.stack([u[keep], v[keep]], 1), axis=0)
    if rng is not None:
        e = e[rng.permutation(len(e))]
        e = e[np.argsort(e[:, 0], kind="stable")]
    ptr = np.zeros(n + 1, np.int64)
    np.add.at(ptr, e[:, 0] + 1, 1)
    return np.cumsum(ptr).astype(np.int32), e[:, 1].astype(np.int32)


def _serial_sweep(ptr, adj, start):
    n = ptr.size - 1
    lvl = np.full(n, -1, np.int32)
    visit = [start]
    lvl[start] = 0
    head = 0
    while head < len(visit):
        u = visit[head]
        head += 1
        for v in adj[ptr[u]:ptr[u + 1]]:
            if lvl[v] < 0:
                lvl[v] = lvl[u] + 1
                visit.append(int(v))
    return lvl, np.array(visit, np.int32)


def _serial_pseudo_peripheral(ptr, adj):
    lvl, visit = _serial_sweep(ptr, adj, 0)
    if visit.size != ptr.size - 1:
        return None
    last = visit[lvl[visit] == lvl.max()]
    deg = ptr[last + 1] - ptr[last]
    best = int(last[np.flatnonzero(deg == deg.min())[-1]])      # minimum degree, the latest visited among equals
    if best != 0:
        lvl, visit = _serial_sweep(ptr, adj, best)
    return lvl, visit


def _grid(side, reach, rng):
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    src, dst = [], []
    for di in range(-reach, reach + 1):
        for dj in range(-reach, reach + 1):
            if (di, dj) <= (0, 0) or di * di + dj * dj > reach * reach:
                continue
            ok = (ii + di >= 0) & (ii + di < side) & (jj + dj >= 0) & (jj + dj < side) & (rng.random((side, side)) < 0.8)
            src.append((ii * side + jj)[ok])
            dst.append(((ii + di) * side + jj + dj)[ok])
    return side * side, np.concatenate(src), np.concatenate(dst)


@pytest.mark.parametrize("side,reach,seed", [(12, 1, 0), (40, 2, 1), (90, 2, 2), (150, 3, 3)])
def test_sweep_matches_the_serial_queue_order(side, reach, seed):
    rng = np.random.default_rng(seed)
    n, src, dst = _grid(side, reach, rng)
    ptr, adj = _csr(n, src, dst, rng)
    want = _serial_pseudo_peripheral(ptr, adj)
    got = _capi.level_sweep(ptr, adj, -1)
    if want is None:
        assert got is None
        return
    assert got is not None
    lvl, visit, nlev = got
    assert np.array_equal(visit, want[1])
    assert np.array_equal(lvl, want[0])
    assert nlev == want[0].max() + 1
    # a single sweep from a given start
    start = int(rng.integers(0, n))
    want1 = _serial_sweep(ptr, adj, start)
    got1 = _capi.level_sweep(ptr, adj, start)
    assert got1 is not None
    assert np.array_equal(got1[1], want1[1]) and np.array_equal(got1[0], want1[0])


def test_sweep_with_long_rows_and_random_structure():
    """hubs with rows far above one warp's 32 lanes, small-world shortcuts"""
    rng = np.random.default_rng(7)
    n = 6000
    ring = np.arange(n)
    src = [ring, ring, rng.integers(0, n, 3 * n)]
    dst = [(ring + 1) % n, (ring + 7) % n, rng.integers(0, n, 3 * n)]
    for hub in (5, 1234):
        nb = rng.choice(n, 700, replace=False)
        src.append(np.full(nb.size, hub))
        dst.append(nb)
    ptr, adj = _csr(n, np.concatenate(src), np.concatenate(dst), rng)
    assert (ptr[1:] - ptr[:-1]).max() > 600
    want = _serial_pseudo_peripheral(ptr, adj)
    got = _capi.level_sweep(ptr, adj, -1)
    assert want is not None and got is not None
    assert np.array_equal(got[1], want[1]) and np.array_equal(got[0], want[0])


def test_sweep_declines_what_it_cannot_do():
    # two components
    ptr, adj = _csr(6, np.array([0, 1, 3, 4]), np.array([1, 2, 4, 5]))
    assert _capi.level_sweep(ptr, adj, -1) is None
    # a chain: one node per level, the host is faster
    n = 5000
    ptr, adj = _csr(n, np.arange(n - 1), np.arange(1, n))
    assert _capi.level_sweep(ptr, adj, -1) is None
    # a short chain is below the cut-off and handled
    ptr, adj = _csr(40, np.arange(39), np.arange(1, 40))
    lvl, visit, nlev = _capi.level_sweep(ptr, adj, -1)
    assert nlev == 40 and visit[0] == 39 and lvl[0] == 39


@pytest.mark.parametrize("side,sweep_min", [(120, 2000), (180, 8000)])
def test_elimination_order_does_not_depend_on_who_sweeps(side, sweep_min):
    """the threshold is lowered so that several levels of the dissection (both halves of a split, concurrently)
    are swept on the device"""
    n, src, dst, k, A, B = synth.landscape_2d(side, T=0.3, seed=4)
    dev = Handle.from_coo(n, src, dst, k, A, B, sweep_min_nodes=sweep_min)
    host = Handle.from_coo(n, src, dst, k, A, B, host_sweeps=1)
    assert np.array_equal(dev.elimination_order(), host.elimination_order())
    dev.compute_rates()
    host.compute_rates()
    assert np.array_equal(dev.rates(), host.rates())
    dev.close()
    host.close()

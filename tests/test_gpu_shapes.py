"""Awkward front shapes through the blocked elimination, against the CPU oracle (C ABI, GPU).

One dense front of order n exercises, depending on n: short last panels (n-|A|-|B| not a multiple of 32), several outer
blocks with the delayed D1/D2 updates and the look-ahead stream (k > 256), odd tile origins (16-byte paths off),
half-height update tiles, the cluster variant of the panel kernel (few fronts) and, with a batch of networks, the
many-front variants.  Group sizes > 1 add phase-2 task fronts of every small order."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _dense_case(n, nA, nB, seed, decades=3.0):
    rng = np.random.default_rng(seed)
    K = np.exp(rng.uniform(-decades * 2.3, 0.0, (n, n)))
    np.fill_diagonal(K, 0.0)
    nodes = rng.permutation(n)
    A = [int(x) for x in nodes[:nA]]
    B = [int(x) for x in nodes[nA:nA + nB]]
    return K, A, B


@pytest.mark.parametrize("n,nA,nB", [(35, 1, 1), (66, 2, 1), (97, 1, 3), (130, 3, 2), (257, 1, 1), (259, 2, 2),
                                      (289, 1, 1), (300, 5, 7), (321, 1, 2), (515, 1, 1), (580, 2, 1), (700, 3, 3)])
def test_dense_front_of_awkward_order(n, nA, nB):
    from kmc_rates_b200 import NGT
    from oracle.cpu_ngt import CpuNGT
    K, A, B = _dense_case(n, nA, nB, seed=n)
    with_committors = n <= 130     # the oracle re-runs an elimination per node for them (reference ngt.hpp:553-608)
    g = NGT.from_dense(K, A, B)
    i, j = np.nonzero(~np.eye(n, dtype=bool))
    cpu = CpuNGT(n, i, j, K[i, j], A, B, kind="port")
    if with_committors:
        g.compute_rates_and_committors()
        cpu.compute_rates_and_committors()
    else:
        g.compute_rates()
        cpu.compute_rates()
    got = [g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()]
    assert rel_err(got, cpu.rates()) < RTOL
    om, tau = g.get_final()
    com, ctau = cpu.final()
    assert rel_err(om, com) < RTOL and rel_err(tau, ctau) < RTOL
    if with_committors:
        assert np.max(np.abs(g.get_committor_array() - cpu.committors())) < RTOL


@pytest.mark.parametrize("n,nbatch", [(45, 3), (100, 5), (290, 2), (40, 320)])
def test_batched_dense_fronts(n, nbatch):
    """few networks -> cluster / latency panel variants with gridDim.y > 1; many networks -> the 128-thread variant"""
    from kmc_rates_b200 import NGT
    from oracle.cpu_ngt import CpuNGT
    rng = np.random.default_rng(100 + n)
    K = np.exp(rng.uniform(-5.0, 0.0, (nbatch, n, n)))
    for b in range(nbatch):
        np.fill_diagonal(K[b], 0.0)
    A, B = [0, 3], [n - 1]
    g = NGT.from_dense(K, A, B)
    g.compute_rates()
    kab, kba = g.get_rate_AB(), g.get_rate_BA()
    i, j = np.nonzero(~np.eye(n, dtype=bool))
    for b in sorted(set([0, 1, nbatch // 2, nbatch - 1])):
        cpu = CpuNGT(n, i, j, K[b][i, j], A, B, kind="port")
        cpu.compute_rates()
        r = cpu.rates()
        assert abs(kab[b] - r[0]) < RTOL * r[0] and abs(kba[b] - r[1]) < RTOL * r[1]


def test_gemm_paths_agree_on_a_multi_block_front():
    """DMMA (default), scalar DFMA and the 128x128 DMMA tile give the same rates on a front with three outer blocks"""
    from kmc_rates_b200 import NGT
    K, A, B = _dense_case(650, 2, 2, seed=4)
    out = []
    for path in (0, 1, 2):
        g = NGT.from_dense(K, A, B, gemm_path=path)
        g.compute_rates()
        out.append([g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()])
    assert rel_err(out[1], out[0]) < 1e-11 and rel_err(out[2], out[0]) < 1e-11

"""GPU tests at BASELINE.json's full sizes.  The elimination oracles cannot reach these sizes (the reference would
need hours), so parity rests on size-independent properties and independent solvers:

  * probability conservation: every row of the reduced A ∪ B graph sums to 1 (sum_v P_uv = 1 is invariant under
    node elimination, reference README.rst:57-59);
  * the independent linear solve K_II t = -1 (scipy SuperLU for the sparse config, cuSOLVER through
    torch.linalg.solve for the dense one) -- checkers only, never on the product path;
  * the closed-form phase-2 identity 1-P'_aa = 1/G_aa;
  * the reference port on a sample of the batched networks.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_sparse_landscape_1e5_nodes():
    from kmc_rates_b200 import NGT, synth
    from oracle.py_linalg import TwoStateRatesCOO
    n, src, dst, k, A, B = synth.landscape_2d(316, T=0.3, seed=0)
    g = NGT.from_coo(n, src, dst, k, A, B)
    g.compute_rates_and_committors()
    P, _tau = g._h.reduced()
    assert np.max(np.abs(P.sum(axis=1) - 1.0)) < 1e-12
    la = TwoStateRatesCOO(n, src, dst, k, A, B)
    la.compute_rates()
    la.compute_committors()
    assert abs(g.get_rate_AB() - la.get_rate_AB()) < 1e-9 * la.get_rate_AB()
    q = g.get_committor_array()
    assert np.nanmax(np.abs(q - la.q)) < 1e-9
    assert abs(g.get_rate_AB_SS() - la.get_rate_AB_SS()) < 1e-9 * la.get_rate_AB_SS()
    assert g.stats()["n_eliminated"] == n - 2


def test_config3_dense_16384_nodes():
    import torch
    from kmc_rates_b200 import NGT
    n = 16384
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    K = torch.rand(n, n, generator=gen, dtype=torch.float64, device="cuda")
    K.fill_diagonal_(0.0)
    g = NGT.from_dense(K, [0], [1])
    g.compute_rates()
    P, _tau = g._h.reduced()
    assert np.max(np.abs(P.sum(axis=1) - 1.0)) < 1e-11
    # independent check: mean first-passage time into B = {1} by a dense LU solve on the device (cuSOLVER)
    M = K.clone()
    M -= torch.diag(K.sum(dim=1))
    keep = torch.ones(n, dtype=torch.bool, device="cuda")
    keep[1] = False
    t = torch.linalg.solve(M[keep][:, keep], -torch.ones(n - 1, dtype=torch.float64, device="cuda"))
    kab = 1.0 / float(t[0])
    assert abs(g.get_rate_AB() - kab) < 1e-9 * kab


def test_config4_groups_of_512_closed_form_identity():
    from kmc_rates_b200 import synth
    from kmc_rates_b200._capi import Handle
    gsz = 512
    n, src, dst, k, A, B = synth.landscape_groups(141, gsz, gsz, T=0.3, seed=0)
    h = Handle.from_coo(n, src, dst, k, A, B)
    h.compute_rates()
    om, tau = h.final()
    P, t = h.reduced()
    assert np.max(np.abs(P.sum(axis=1) - 1.0)) < 1e-12
    for lo, hi in ((0, gsz), (gsz, 2 * gsz)):
        G = np.linalg.inv(np.eye(gsz) - P[lo:hi, lo:hi])
        assert np.max(np.abs(om[lo:hi] * np.diag(G) - 1.0)) < 1e-9
        assert np.max(np.abs(tau[lo:hi] * np.diag(G) / (G @ t[lo:hi]) - 1.0)) < 1e-9
    r = h.rates()
    assert np.all(np.isfinite(r)) and np.all(r > 0)


def test_config5_batched_sweep_4096_networks():
    import torch
    from kmc_rates_b200 import NGT
    from oracle.cpu_ngt import CpuNGT
    NB, N = 4096, 256
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    E = torch.rand(N, generator=gen, dtype=torch.float64, device="cuda")
    bar = torch.rand(N, N, generator=gen, dtype=torch.float64, device="cuda") * 0.9 + 0.1
    bar = torch.maximum(bar, bar.T)
    temps = torch.logspace(np.log10(0.05), 0.0, NB, dtype=torch.float64, device="cuda")
    K = torch.exp(-((torch.maximum(E[:, None], E[None, :]) + bar)[None] - E[None, :, None]) / temps[:, None, None])
    K.diagonal(dim1=1, dim2=2).zero_()
    g = NGT.from_dense(K, [0], [N - 1])
    g.compute_rates()
    kab, kba = g.get_rate_AB(), g.get_rate_BA()
    assert kab.shape == (NB,) and np.all(np.isfinite(kab)) and np.all(kab > 0) and np.all(kba > 0)
    assert np.all(np.diff(kab) > 0)          # rates grow with temperature along the sweep
    i, j = np.nonzero(~np.eye(N, dtype=bool))
    for b in (0, 1000, 2500, 4095):          # the reference port on a sample of the networks
        Kb = K[b].cpu().numpy()
        cpu = CpuNGT(N, i, j, Kb[i, j], [0], [N - 1], kind="port")
        cpu.compute_rates()
        r = cpu.rates()
        assert abs(kab[b] - r[0]) < 1e-9 * r[0] and abs(kba[b] - r[1]) < 1e-9 * r[1]


def test_config2_ill_conditioned_T003_full_size():
    """SURVEY §8(d)'s ill-conditioned variant of config 2 at full size: T = 0.03, rate constants down to ~1e-29, P_xx -> 1
    after a few eliminations.  No CPU elimination reaches this size and the LU-based linear solve loses digits here, so
    the evidence is: (1) conservation and range properties, (2) two different elimination orders and two different GEMM
    paths (tensor-core and scalar) agree to 1e-9 -- subtraction-free pivots make every order accurate, a cancelling
    1 - P_xx would not survive this -- and (3) the sparse LU solve where it still holds (1e-6)."""
    from kmc_rates_b200 import NGT, synth
    from oracle.py_linalg import TwoStateRatesCOO
    n, src, dst, k, A, B = synth.landscape_2d(316, T=0.03, seed=0)
    assert k.min() < 1e-25
    g = NGT.from_coo(n, src, dst, k, A, B)
    g.compute_rates_and_committors()
    P, _tau = g._h.reduced()
    assert np.max(np.abs(P.sum(axis=1) - 1.0)) < 1e-12
    q, t = g.get_committor_array(), g.get_mfpt_array()
    assert np.nanmin(q) >= 0.0 and np.nanmax(q) <= 1.0 + 1e-12
    assert np.all(t[np.isfinite(t)] >= 0.0)
    r = np.array([g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()])
    assert np.all(np.isfinite(r)) and np.all(r > 0)
    g2 = NGT.from_coo(n, src, dst, k, A, B, leaf_size=24, gemm_path=1)      # another elimination tree, scalar DFMA updates
    g2.compute_rates_and_committors()
    r2 = np.array([g2.get_rate_AB(), g2.get_rate_BA(), g2.get_rate_AB_SS(), g2.get_rate_BA_SS()])
    assert np.max(np.abs(r2 - r) / r) < 1e-9
    assert np.nanmax(np.abs(g2.get_committor_array() - q)) < 1e-9
    la = TwoStateRatesCOO(n, src, dst, k, A, B)
    la.compute_committors()
    assert np.nanmax(np.abs(q - la.q)) < 1e-6

"""Edge cases through the C ABI on the GPU: empty intermediate set, two-node graphs, pruned components, duplicate
COO entries (last wins, like the reference's std::map), self rates on the sparse path, labels of mixed types."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _cpu(n, src, dst, k, A, B, **kw):
    from oracle.cpu_ngt import CpuNGT
    o = CpuNGT(n, src, dst, k, A, B, **kw)
    o.compute_rates()
    return o


def test_two_node_graph_no_intermediates():
    from kmc_rates_b200 import NGT
    ngt = NGT({(0, 1): 2.0, (1, 0): 0.5}, [0], [1])
    ngt.compute_rates_and_committors()
    assert abs(ngt.get_rate_AB() - 2.0) < 1e-12 and abs(ngt.get_rate_BA() - 0.5) < 1e-12
    assert ngt.get_committors() == {0: 0.0, 1: 1.0}


def test_every_node_in_A_or_B():
    from kmc_rates_b200 import NGT, synth
    src, dst, k = synth.complete_random(9, seed=3)
    A, B = [0, 1, 2, 3], [4, 5, 6, 7, 8]
    ngt = NGT.from_coo(9, src, dst, k, A, B)
    ngt.compute_rates()
    cpu = _cpu(9, src, dst, k, A, B)
    got = [ngt.get_rate_AB(), ngt.get_rate_BA(), ngt.get_rate_AB_SS(), ngt.get_rate_BA_SS()]
    assert rel_err(got, cpu.rates()) < 1e-12


def test_single_intermediate():
    from kmc_rates_b200 import NGT
    rates = {(0, 2): 1.0, (2, 0): 3.0, (2, 1): 1.0, (1, 2): 2.0}
    ngt = NGT(rates, [0], [1])
    ngt.compute_rates_and_committors()
    # from 2: to B with prob 1/4 -> k_AB = 1 * 1/4 ... MFPT(0->1) = 1 + 4*(1/4) ... check against the oracle instead
    src = np.array([0, 2, 2, 1]); dst = np.array([2, 0, 1, 2]); k = np.array([1.0, 3.0, 1.0, 2.0])
    cpu = _cpu(3, src, dst, k, [0], [1])
    assert rel_err([ngt.get_rate_AB(), ngt.get_rate_BA()], cpu.rates()[:2]) < 1e-12
    assert abs(ngt.get_committors()[2] - 0.25) < 1e-12


def test_unreachable_component_is_pruned_and_reads_nan():
    from kmc_rates_b200 import NGT
    rates = {(0, 1): 1.0, (1, 0): 1.0, (1, 2): 1.0, (2, 1): 1.0, (7, 8): 1.0, (8, 7): 1.0, (8, 9): 2.0, (9, 8): 1.0}
    ngt = NGT(rates, [0], [2])
    ngt.compute_rates_and_committors()
    q = ngt.get_committors()
    assert abs(q[1] - 0.5) < 1e-12 and np.isnan(q[7]) and np.isnan(q[8]) and np.isnan(q[9])
    assert abs(ngt.get_rate_AB() - 1.0 / 3.0) < 1e-12   # chain 0-1-2, unit rates: MFPT 0->2 is 3
    assert ngt.stats()["n_eliminated"] == 1


def test_duplicate_coo_entries_last_wins():
    from kmc_rates_b200 import NGT, synth
    src, dst, k = synth.complete_random(12, seed=1)
    # prepend stale copies of the first 20 entries with different values
    src2 = np.concatenate([src[:20], src]); dst2 = np.concatenate([dst[:20], dst]); k2 = np.concatenate([k[:20] * 7 + 1, k])
    a = NGT.from_coo(12, src2, dst2, k2, [0], [11]); a.compute_rates()
    b = NGT.from_coo(12, src, dst, k, [0], [11]); b.compute_rates()
    assert a.get_rate_AB() == b.get_rate_AB() and a.get_rate_BA() == b.get_rate_BA()


def test_self_rates_on_sparse_and_dense_paths_agree_with_oracle():
    from kmc_rates_b200 import NGT, synth
    rng = np.random.default_rng(0)
    n = 30
    src, dst, k = synth.complete_random(n, seed=4)
    loops = np.arange(n)
    src = np.concatenate([src, loops]); dst = np.concatenate([dst, loops]); k = np.concatenate([k, rng.random(n)])
    A, B = [0, 1], [28, 29]
    cpu = _cpu(n, src, dst, k, A, B)
    s = NGT.from_coo(n, src, dst, k, A, B); s.compute_rates()
    K = np.zeros((n, n)); K[src, dst] = k
    d = NGT.from_dense(K, A, B); d.compute_rates()
    for g in (s, d):
        got = [g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()]
        assert rel_err(got, cpu.rates()) < 1e-11


def test_mixed_hashable_labels_and_weights():
    from kmc_rates_b200 import NGT, GraphReduction
    rates = {("a", 1): 1.0, (1, "a"): 2.0, ("a", (0, 0)): 1.0, ((0, 0), "a"): 1.0, (1, (0, 0)): 1.0, ((0, 0), 1): 1.0,
             ((0, 0), "z"): 0.5, ("z", (0, 0)): 0.25, ("z", 1): 1.5, (1, "z"): 0.75}
    w = {"a": 0.2, "z": 0.8, 1: 1.0}
    g = NGT(rates, ["a", "z"], [1], weights=w); g.compute_rates()
    r = GraphReduction(rates, ["a", "z"], [1], weights=w); r.compute_rates()
    assert abs(g.get_rate_AB() - r.get_rate_AB()) < 1e-13 * g.get_rate_AB()
    from oracle.py_ngt import PyGraphReduction
    o = PyGraphReduction(rates, ["a", "z"], [1], weights=w); o.compute_rates()
    assert abs(g.get_rate_AB() - o.get_rate_AB()) < 1e-12 * o.get_rate_AB()
    assert abs(g.get_rate_BA() - o.get_rate_BA()) < 1e-12 * o.get_rate_BA()


def test_set_rates_reuses_the_analysis():
    """temperature sweep: same topology, new rate constants, no new symbolic analysis"""
    from kmc_rates_b200 import synth
    from kmc_rates_b200._capi import Handle
    n, src, dst, k, A, B = synth.landscape_2d(20, T=0.3, seed=5)
    h = Handle.from_coo(n, src, dst, k, A, B)
    h.compute_rates()
    r1 = h.rates().copy()
    k2 = k ** 1.3
    h.set_rates(k2)
    h.compute_rates()
    h.compute_rates()      # third step runs from the captured CUDA graph
    r2 = h.rates().copy()
    cpu = _cpu(n, src, dst, k2, A, B)
    assert rel_err(r2, cpu.rates()) < 1e-9
    assert rel_err(r1, _cpu(n, src, dst, k, A, B).rates()) < 1e-9


def test_large_ill_conditioned_landscape_against_linear_solve_and_properties():
    """n = 10^4 at T = 0.05 (rates spanning ~17 decades): committors within [0,1], monotone bounds, and k_AB against
    the independent sparse linear solve where that solve is still accurate (1e-6)"""
    from kmc_rates_b200 import NGT, synth
    from oracle.py_linalg import TwoStateRatesCOO
    n, src, dst, k, A, B = synth.landscape_2d(100, T=0.05, seed=3)
    g = NGT.from_coo(n, src, dst, k, A, B)
    g.compute_rates_and_committors()
    q = g.get_committor_array()
    assert np.nanmin(q) >= 0.0 and np.nanmax(q) <= 1.0 + 1e-12
    t = g.get_mfpt_array()
    assert np.all(t[np.isfinite(t)] >= 0.0)
    la = TwoStateRatesCOO(n, src, dst, k, A, B)
    la.compute_committors()
    assert np.nanmax(np.abs(q - la.q)) < 1e-6


def test_set_rates_with_a_device_tensor_produced_on_the_callers_stream():
    """rate constants produced by torch kernels (on torch's stream) and handed over as a CUDA tensor: the engine's
    own non-blocking streams must see the finished values (sparse: copied in; dense: borrowed)"""
    import torch
    from kmc_rates_b200 import synth
    from kmc_rates_b200._capi import Handle
    n, src, dst, k, A, B = synth.landscape_2d(40, T=0.3, seed=4)
    h = Handle.from_coo(n, src, dst, k, A, B)
    h.compute_rates()
    base = h.rates().copy()
    kd = torch.from_numpy(k).cuda()
    for scale in (2.0, 0.25):
        big = torch.randn(4096, 4096, device="cuda") @ torch.randn(4096, 4096, device="cuda")   # keep torch's stream busy
        kd2 = kd * scale + 0.0 * big[0, 0].double()
        h.set_rates(kd2)
        h.compute_rates()
        assert np.allclose(h.rates(), base * scale, rtol=1e-12, atol=0)     # all rates are linear in the rate constants
    s2, d2, k2 = synth.complete_random(300, seed=2)
    K = torch.zeros(300, 300, dtype=torch.float64, device="cuda")
    K[torch.from_numpy(s2).cuda(), torch.from_numpy(d2).cuda()] = torch.from_numpy(k2).cuda()
    hd = Handle.from_dense(K, [0], [299])
    hd.compute_rates()
    r0 = hd.rates().copy()
    K3 = K.clone().mul_(3.0)
    hd.set_rates(K3)
    hd.compute_rates()
    assert np.allclose(hd.rates(), 3.0 * r0, rtol=1e-12, atol=0)


def test_batched_handle_final_and_single_network_getters():
    """final() of a batched handle has one row per network; getters that return one network refuse a batch"""
    from kmc_rates_b200 import synth
    from kmc_rates_b200._capi import Handle
    from oracle.cpu_ngt import CpuNGT
    n, src, dst, k, A, B = synth.landscape_groups(12, 3, 2, T=0.3, seed=9)
    ks = np.stack([k, 2.0 * k, k ** 1.1])
    h = Handle.from_coo(n, src, dst, ks.ravel(), A, B, nbatch=3)
    h.compute_rates()
    om, tau = h.final()
    assert om.shape == (3, 5) and tau.shape == (3, 5)
    for b in range(3):
        cpu = CpuNGT(n, src, dst, ks[b], A, B, kind="port")
        cpu.compute_rates()
        o, t = cpu.final()
        assert np.allclose(om[b], o, rtol=1e-9, atol=0) and np.allclose(tau[b], t, rtol=1e-9, atol=0)
        assert np.allclose(h.rates()[b], cpu.rates(), rtol=1e-9, atol=0)
    h.set_final(om, tau)                       # symmetric with final(): nbatch x (|A|+|B|)
    with pytest.raises(ValueError):
        h.set_final(om[0], tau[0])
    for getter in (h.reduced, h.committors, h.mfpt, h.initial_tau):
        with pytest.raises(ValueError):
            getter()


def test_phase_two_tree_odd_group_sizes_and_weights_against_the_reference():
    """phase 2 (a binary tree of eliminations here, a loop over a in the reference, ngt.hpp:362-431) on group sizes
    that are not powers of two, including one-member groups, against the compiled reference when present"""
    from kmc_rates_b200 import NGT, synth
    from oracle.cpu_ngt import CpuNGT, best_kind
    for nA, nB, seed in [(1, 7, 1), (5, 1, 2), (13, 11, 3), (33, 2, 4), (64, 65, 5)]:
        n, src, dst, k, A, B = synth.landscape_groups(18, nA, nB, T=0.3, seed=seed)
        rng = np.random.default_rng(seed)
        w = {int(x): float(rng.uniform(0.1, 1.0)) for x in A + B}
        g = NGT.from_coo(n, src, dst, k, A, B, weights=w)
        g.compute_rates()
        cpu = CpuNGT(n, src, dst, k, A, B, weights=w, kind=best_kind())
        cpu.compute_rates()
        got = np.array([g.get_rate_AB(), g.get_rate_BA(), g.get_rate_AB_SS(), g.get_rate_BA_SS()])
        assert np.allclose(got, cpu.rates(), rtol=1e-9, atol=0), (nA, nB)
        om, tau = g._h.final()
        o, t = cpu.final()
        assert np.allclose(om, o, rtol=1e-9, atol=0) and np.allclose(tau, t, rtol=1e-9, atol=0), (nA, nB)

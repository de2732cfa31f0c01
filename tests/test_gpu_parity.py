"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(kmc_rates_b200/_capi.py -> libngt_b200.so); the oracle is only the checker.

Tolerance: BASELINE.json's north star asks for rel 1e-9 on rates, committors and MFPTs.  Ill-conditioned golden
cases carry their own rtol (the reference's own run-to-run spread, see oracle/make_golden.py) and an
extended-precision truth against which the GPU result must hold 1e-9."""
import numpy as np
import pytest

from conftest import golden_cases, golden_input, rel_err

pytestmark = pytest.mark.gpu

CASES = golden_cases()
IDS = [c["name"] for c in CASES]
RTOL = 1e-9


def _expect_rates(e):
    return [e["kAB"], e["kBA"], e["kAB_SS"], e["kBA_SS"]]


def _dense_K(n, src, dst, k):
    K = np.zeros((n, n))
    K[src, dst] = k
    return K


@pytest.mark.parametrize("gemm_path", [0, 1, 2], ids=["dmma", "dfma", "dmma128"])
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_sparse_path_matches_reference_golden(case, gemm_path):
    from kmc_rates_b200 import NGT
    n, src, dst, k, A, B, w = golden_input(case)
    ngt = NGT.from_coo(n, src, dst, k, A, B, weights=w, gemm_path=gemm_path, leaf_size=case.get("leaf", 0))
    ngt.compute_rates()
    e = case["expect"]
    tol = max(RTOL, case.get("rtol", 0))
    got = [ngt.get_rate_AB(), ngt.get_rate_BA(), ngt.get_rate_AB_SS(), ngt.get_rate_BA_SS()]
    assert rel_err(got, _expect_rates(e)) < tol
    om, tau = ngt.get_final()
    assert rel_err(om, e["final_omPxx"]) < tol
    assert rel_err(tau, e["final_tau"]) < tol
    P, t1 = ngt._h.reduced()
    assert rel_err(t1, e["reduced_tau"]) < tol
    Pe = np.array(e["reduced_P"])
    off = ~np.eye(len(Pe), dtype=bool)
    assert np.max(np.abs(P[off] - Pe[off]) / np.maximum(np.abs(Pe[off]), 1e-300) * (Pe[off] > 1e-300)) < 10 * tol
    assert np.max(np.abs(np.diag(P) - np.diag(Pe))) < 1e-9
    if "truth_mp" in case:
        t = case["truth_mp"]
        assert rel_err(got, [t["kAB"], t["kBA"], t["kAB_SS"], t["kBA_SS"]]) < RTOL


COMM = [c for c in CASES if c.get("committors")]


@pytest.mark.parametrize("case", COMM, ids=[c["name"] for c in COMM])
def test_committors_match_reference_golden(case):
    """compute_rates_and_committors (reference ngt.hpp:616-630) by back-substitution"""
    from kmc_rates_b200 import NGT
    n, src, dst, k, A, B, w = golden_input(case)
    ngt = NGT.from_coo(n, src, dst, k, A, B, weights=w)
    ngt.compute_rates_and_committors()
    q = ngt.get_committor_array()
    want = np.array([np.nan if x is None else x for x in case["expect"]["committors"]])
    ok = ~np.isnan(want)
    assert np.max(np.abs(q[ok] - want[ok])) < 1e-9
    assert rel_err([ngt.get_rate_AB(), ngt.get_rate_BA()], [case["expect"]["kAB"], case["expect"]["kBA"]]) < RTOL


DENSE = [c for c in CASES if c["gen"] in ("complete_deterministic", "complete_random")]


@pytest.mark.parametrize("gemm_path", [0, 1, 2], ids=["dmma", "dfma", "dmma128"])
@pytest.mark.parametrize("case", DENSE, ids=[c["name"] for c in DENSE])
def test_dense_path_matches_reference_golden(case, gemm_path):
    from kmc_rates_b200 import NGT
    n, src, dst, k, A, B, w = golden_input(case)
    ngt = NGT.from_dense(_dense_K(n, src, dst, k), A, B, weights=w, gemm_path=gemm_path)
    ngt.compute_rates()
    e = case["expect"]
    got = [ngt.get_rate_AB(), ngt.get_rate_BA(), ngt.get_rate_AB_SS(), ngt.get_rate_BA_SS()]
    assert rel_err(got, _expect_rates(e)) < RTOL
    om, tau = ngt.get_final()
    assert rel_err(om, e["final_omPxx"]) < RTOL
    assert rel_err(tau, e["final_tau"]) < RTOL


def test_dense_device_tensor_zero_copy():
    """rate matrix already resident in HBM as a torch float64 tensor"""
    import torch
    from kmc_rates_b200 import NGT
    case = [c for c in CASES if c["name"] == "c1_complete100"][0]
    n, src, dst, k, A, B, w = golden_input(case)
    K = torch.from_numpy(_dense_K(n, src, dst, k)).cuda()
    ngt = NGT.from_dense(K, A, B)
    ngt.compute_rates()
    assert rel_err([ngt.get_rate_AB(), ngt.get_rate_BA()], [case["expect"]["kAB"], case["expect"]["kBA"]]) < RTOL


def test_self_rates_ngt3():
    """reference cpp_tests/source/test_ngt.cpp:9-46: all 9 rates = 1 including self rates"""
    from kmc_rates_b200 import NGT
    rates = {(i, j): 1.0 for i in range(3) for j in range(3)}
    ngt = NGT(rates, [0], [1])
    ngt.compute_rates_and_committors()
    assert abs(ngt.get_rate_AB() - 1) < 1e-9 and abs(ngt.get_rate_BA() - 1) < 1e-9
    assert abs(ngt.get_rate_AB_SS() - 1.5) < 1e-9 and abs(ngt.get_rate_BA_SS() - 1.5) < 1e-9
    assert abs(ngt.get_committors()[2] - 0.5) < 1e-9


def test_medium_landscape_against_cpu_oracle_and_linear_solve():
    """n = 2500 landscape with a small leaf size so the assembly tree is deep; port oracle (rel 1e-9) and
    the independent linear-solve oracle (reference rates_linalg.TwoStateRates restated)."""
    from kmc_rates_b200 import NGT, synth
    from oracle.cpu_ngt import CpuNGT
    from oracle.py_linalg import TwoStateRatesCOO
    n, src, dst, k, A, B = synth.landscape_2d(50, T=0.3, seed=11)
    for leaf in (16, 64):
        ngt = NGT.from_coo(n, src, dst, k, A, B, leaf_size=leaf)
        ngt.compute_rates_and_committors()
        got = np.array([ngt.get_rate_AB(), ngt.get_rate_BA(), ngt.get_rate_AB_SS(), ngt.get_rate_BA_SS()])
        cpu = CpuNGT(n, src, dst, k, A, B, kind="port")
        cpu.compute_rates()
        assert rel_err(got, cpu.rates()) < RTOL
        la = TwoStateRatesCOO(n, src, dst, k, A, B)
        la.compute_rates()
        la.compute_committors()
        assert abs(got[0] - la.get_rate_AB()) < 1e-8 * got[0]
        q = ngt.get_committor_array()
        assert np.nanmax(np.abs(q - la.q)) < 1e-8
        # MFPT to A ∪ B against a direct solve with A ∪ B absorbing
        la2 = TwoStateRatesCOO(n, src, dst, k, A, A + B)
        t = la2.compute_rates()
        tt = ngt.get_mfpt_array()
        inter = np.ones(n, bool)
        inter[A + B] = False
        assert np.max(np.abs(tt[inter] - t[inter]) / t[inter]) < 1e-8


def test_groups_phase_two_against_oracle():
    """multi-state phase 2 (BASELINE config 4 shape at test size): |A| = |B| = 40 on a 30x30 landscape"""
    from kmc_rates_b200 import NGT, synth
    from oracle.cpu_ngt import CpuNGT
    n, src, dst, k, A, B = synth.landscape_groups(30, 40, 40, T=0.3, seed=4)
    rng = np.random.default_rng(0)
    w = {int(x): float(rng.random()) for x in A + B}
    ngt = NGT.from_coo(n, src, dst, k, A, B, weights=w)
    ngt.compute_rates()
    cpu = CpuNGT(n, src, dst, k, A, B, weights=w, kind="port")
    cpu.compute_rates()
    got = np.array([ngt.get_rate_AB(), ngt.get_rate_BA(), ngt.get_rate_AB_SS(), ngt.get_rate_BA_SS()])
    assert rel_err(got, cpu.rates()) < RTOL
    om, tau = ngt.get_final()
    com, ctau = cpu.final()
    assert rel_err(om, com) < RTOL and rel_err(tau, ctau) < RTOL


def test_batched_networks_match_single_runs():
    """BASELINE config 5 shape: networks sharing a topology, different rates (temperature sweep)"""
    from kmc_rates_b200 import NGT, synth
    from oracle.cpu_ngt import CpuNGT
    n = 64
    rng = np.random.default_rng(5)
    E = rng.random(n)
    Ets = np.maximum(E[:, None], E[None, :]) + rng.uniform(0.1, 1.0, (n, n))
    Ets = np.maximum(Ets, Ets.T)
    temps = np.geomspace(0.1, 1.0, 7)
    K = np.stack([np.exp(-(Ets - E[:, None]) / T) for T in temps])
    for b in range(len(temps)):
        np.fill_diagonal(K[b], 0.0)
    A, B = [0], [n - 1]
    ngt = NGT.from_dense(K, A, B)
    ngt.compute_rates()
    kab, kba = ngt.get_rate_AB(), ngt.get_rate_BA()
    i, j = np.nonzero(~np.eye(n, dtype=bool))
    for b in range(len(temps)):
        cpu = CpuNGT(n, i, j, K[b][i, j], A, B, kind="port")
        cpu.compute_rates()
        r = cpu.rates()
        assert abs(kab[b] - r[0]) < RTOL * r[0] and abs(kba[b] - r[1]) < RTOL * r[1]
    # sparse batched twin
    n2, src, dst, k, A2, B2 = synth.landscape_2d(16, T=0.3, seed=2)
    ks = np.stack([k, k * 2.0, k ** 1.5])
    ngt2 = NGT.from_coo(n2, src, dst, ks, A2, B2, nbatch=3)
    ngt2.compute_rates()
    kab2 = ngt2.get_rate_AB()
    for b in range(3):
        cpu = CpuNGT(n2, src, dst, ks[b], A2, B2, kind="port")
        cpu.compute_rates()
        assert abs(kab2[b] - cpu.rates()[0]) < RTOL * cpu.rates()[0]


def test_linearity_and_time_rescaling_properties():
    """size-independent properties: scaling every rate by c scales k_AB by c and leaves committors unchanged"""
    from kmc_rates_b200 import NGT, synth
    n, src, dst, k, A, B = synth.landscape_2d(40, T=0.2, seed=9)
    a = NGT.from_coo(n, src, dst, k, A, B)
    a.compute_rates_and_committors()
    b = NGT.from_coo(n, src, dst, 3.5 * k, A, B)
    b.compute_rates_and_committors()
    assert abs(b.get_rate_AB() / a.get_rate_AB() - 3.5) < 1e-9
    assert np.nanmax(np.abs(a.get_committor_array() - b.get_committor_array())) < 1e-10


def test_dense_1024_against_port_oracle():
    from kmc_rates_b200 import NGT, synth
    from oracle.cpu_ngt import CpuNGT
    n = 700
    src, dst, k = synth.complete_random(n, seed=8)
    K = _dense_K(n, src, dst, k)
    ngt = NGT.from_dense(K, [0], [1])
    ngt.compute_rates()
    cpu = CpuNGT(n, src, dst, k, [0], [1], kind="port")
    cpu.compute_rates()
    got = np.array([ngt.get_rate_AB(), ngt.get_rate_BA(), ngt.get_rate_AB_SS(), ngt.get_rate_BA_SS()])
    assert rel_err(got, cpu.rates()) < RTOL


def test_errors():
    from kmc_rates_b200 import NGT, synth
    rates = synth.to_dict(*synth.complete_deterministic(5))
    with pytest.raises(ValueError):
        NGT(rates, [0, 1], [1, 2])       # A ∩ B
    with pytest.raises(KeyError):
        NGT(rates, [0], [17])            # unknown label (reference ngt.pyx:86)
    ngt = NGT(rates, [0], [1])
    with pytest.raises(RuntimeError):
        ngt.get_rate_AB()                # before compute_rates

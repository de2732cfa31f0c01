"""CPU tests: pin the oracles (oracle/) against the reference's golden values and against the frozen
outputs of the unmodified reference engine (tests/golden/ngt_golden.json)."""
import numpy as np
import pytest

from conftest import golden_cases, golden_input, rel_err
from oracle import cpu_ngt
from oracle.py_linalg import TwoStateRatesCOO
from oracle.py_ngt import PyGraphReduction
from kmc_rates_b200 import synth

CASES = golden_cases()
IDS = [c["name"] for c in CASES]
TOL = 1e-10  # elimination order differs from the reference only in ties -> rounding-level differences


def _run(kind, case, committors=False):
    n, src, dst, k, A, B, w = golden_input(case)
    o = cpu_ngt.CpuNGT(n, src, dst, k, A, B, weights=w, kind=kind)
    if committors:
        o.compute_rates_and_committors()
    else:
        o.compute_rates()
    return o


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_port_matches_reference_golden(case):
    o = _run("port", case, committors=bool(case.get("committors")))
    e = case["expect"]
    TOL = case.get("rtol", 1e-10)  # ill-conditioned cases: the reference's own run-to-run spread (see make_golden.py)
    r = o.rates()
    assert rel_err(r, [e["kAB"], e["kBA"], e["kAB_SS"], e["kBA_SS"]]) < TOL
    om, tau = o.final()
    assert rel_err(om, e["final_omPxx"]) < TOL
    assert rel_err(tau, e["final_tau"]) < TOL
    P, t1 = o.reduced()
    assert rel_err(t1, e["reduced_tau"]) < TOL
    assert rel_err(P, e["reduced_P"]) < max(TOL, 1e-9) or np.max(np.abs(P - np.array(e["reduced_P"]))) < 1e-12
    if case.get("committors"):
        q = o.committors()
        want = np.array([np.nan if x is None else x for x in e["committors"]])
        ok = ~np.isnan(want)
        assert np.max(np.abs(q[ok] - want[ok])) < 1e-12


@pytest.mark.skipif(not cpu_ngt.available("reference"), reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("case", [c for c in CASES if c["n_nodes"] <= 600], ids=[c["name"] for c in CASES if c["n_nodes"] <= 600])
def test_compiled_reference_reproduces_golden(case):
    o = _run("reference", case)
    e = case["expect"]
    # not bitwise: the reference's elimination order follows heap pointer values, so it differs run to run
    assert rel_err(o.rates(), [e["kAB"], e["kBA"], e["kAB_SS"], e["kBA_SS"]]) < case.get("rtol", 1e-12)


TRUTH = [c for c in CASES if "truth_mp" in c]


@pytest.mark.parametrize("case", TRUTH, ids=[c["name"] for c in TRUTH])
def test_port_against_extended_precision_truth(case):
    o = _run("port", case)
    t = case["truth_mp"]
    assert rel_err(o.rates(), [t["kAB"], t["kBA"], t["kAB_SS"], t["kBA_SS"]]) < 1e-9


def test_reference_published_goldens():
    """Values hard-coded in the reference's own tests (cpp_tests/source/test_ngt.cpp:27-46, 68-93)."""
    by = {c["name"]: c for c in CASES}
    o = _run("port", by["ngt3_self_rates"], committors=True)
    assert np.allclose(o.rates(), [1, 1, 1.5, 1.5], rtol=0, atol=1e-9)
    assert abs(o.committors()[2] - 0.5) < 1e-9
    o = _run("port", by["ngt10"])
    assert np.allclose(o.rates(), [5.1013138820442565, 3, 19.933950145409426, 6.970856553435547], rtol=0, atol=1e-9)
    o = _run("port", by["ngt10_weighted"])
    assert abs(o.rates()[0] - 4.8677045605461533) < 1e-9
    o = _run("port", by["tuple3"])
    assert np.allclose(o.rates()[:2], [1.0, 1.6666666666666667], atol=1e-9)


SMALL = [c for c in CASES if c["n_nodes"] <= 150]


@pytest.mark.parametrize("case", SMALL, ids=[c["name"] for c in SMALL])
def test_python_restatement_matches_golden(case):
    """oracle/py_ngt.py (restated _kmc_rates.py GraphReduction) against the reference engine's values."""
    n, src, dst, k, A, B, w = golden_input(case)
    red = PyGraphReduction(synth.to_dict(src, dst, k), A, B, weights=w)
    inter = [x for x in red.graph.nodes() if x not in A and x not in B]
    q = red.compute_committor_probabilities(inter) if case.get("committors") and inter else {}
    red.compute_rates()
    e = case["expect"]
    assert abs(red.get_rate_AB() - e["kAB"]) <= 1e-9 * abs(e["kAB"])
    assert abs(red.get_rate_BA() - e["kBA"]) <= 1e-9 * abs(e["kBA"])
    assert abs(red.get_rate_AB_SS() - e["kAB_SS"]) <= 1e-9 * abs(e["kAB_SS"])
    assert red.graph.number_of_nodes() == len(A) + len(B)
    for x, qx in q.items():
        want = e["committors"][x]
        if want is not None:
            assert abs(qx - want) < 1e-9
    for i, a in enumerate(A):
        assert abs(red.get_committor_probabilityAB(a) - e["final_omPxx"][i]) < 1e-9


WELL = [c for c in CASES if "T0.0" not in c["name"]]


@pytest.mark.parametrize("case", WELL, ids=[c["name"] for c in WELL])
def test_linear_solve_oracle_matches_golden(case):
    """oracle/py_linalg.py (restated rates_linalg.TwoStateRates): the independent cross-check
    (reference kmc_rates/tests/test_kmc.py:46,59 use 1e-5; well-conditioned cases do far better)."""
    n, src, dst, k, A, B, w = golden_input(case)
    la = TwoStateRatesCOO(n, src, dst, k, A, B, weights=w)
    la.compute_rates()
    la.compute_committors()
    e = case["expect"]
    assert abs(la.get_rate_AB() - e["kAB"]) <= 1e-8 * abs(e["kAB"])
    assert abs(la.get_rate_AB_SS() - e["kAB_SS"]) <= 1e-8 * abs(e["kAB_SS"])
    if case.get("committors"):
        want = np.array([np.nan if x is None else x for x in e["committors"]])
        ok = ~np.isnan(want) & ~np.isnan(la.q)
        assert np.max(np.abs(la.q[ok] - want[ok])) < 1e-8


def test_python_restatement_errors():
    rates = synth.to_dict(*synth.complete_deterministic(5))
    with pytest.raises(ValueError):
        PyGraphReduction(rates, [0, 1], [1, 2])
    with pytest.raises(ValueError):
        PyGraphReduction(rates, [0], [17])
    red = PyGraphReduction(rates, [0], [1])
    with pytest.raises(RuntimeError):
        red.get_committor_probabilityAB(0)
    red.compute_rates()
    with pytest.raises(ValueError):
        red.get_committor_probabilityAB(3)

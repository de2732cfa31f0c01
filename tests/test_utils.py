"""CPU tests of the PATHSAMPLE ingest (kmc_rates_b200/utils.py) against the loop-for-loop restatement of the
reference's reader, and of the in.* text format round trip."""
import os

import numpy as np
import pytest


def _write_db(d, nmin=40, nts=150, seed=0):
    rng = np.random.default_rng(seed)
    E = rng.uniform(-5, 0, nmin)
    with open(os.path.join(d, "min.data"), "w") as f:
        for i in range(nmin):
            f.write("%.10f %.10f %d %.3f %.3f %.3f\n" % (E[i], rng.uniform(50, 60), rng.choice([1, 2, 4]), 1, 1, 1))
    with open(os.path.join(d, "ts.data"), "w") as f:
        for t in range(nts):
            if t < nmin - 1:
                u, v = t, t + 1                      # a chain keeps everything connected
            else:
                u, v = rng.integers(0, nmin, 2)      # includes parallel and self-connecting transition states
            ets = max(E[u], E[v]) + rng.uniform(0.1, 2.0)
            f.write("%.10f %.10f %d %d %d %.3f %.3f %.3f\n" % (ets, rng.uniform(50, 60), rng.choice([1, 2]), u + 1, v + 1, 1, 1, 1))
    with open(os.path.join(d, "min.A"), "w") as f:
        f.write("2\n1 2\n")
    with open(os.path.join(d, "min.B"), "w") as f:
        f.write("3\n%d\n%d %d\n" % (nmin, nmin - 1, nmin - 2))


@pytest.mark.parametrize("T", [1.0, 0.2])
def test_make_rates_matches_reference_restatement(tmp_path, T):
    from kmc_rates_b200 import utils
    from oracle.py_utils import rates_from_pathsample
    d = str(tmp_path)
    _write_db(d)
    rates, Peq, norm = utils.make_rates(d, T)
    want_rates, want_Peq, want_norm = rates_from_pathsample(os.path.join(d, "min.data"), os.path.join(d, "ts.data"), T)
    assert set(rates) == set(want_rates)
    for uv, k in want_rates.items():
        assert abs(rates[uv] - k) <= 1e-12 * k
    assert set(Peq) == set(want_Peq)
    for u, p in want_Peq.items():
        assert abs(Peq[u] - p) <= 1e-14 * max(p, 1e-300)
    assert abs(norm - want_norm) <= 1e-14 * want_norm
    assert max(rates.values()) <= 1.0 + 1e-12          # shifted by the largest log rate
    assert all(u != v for (u, v) in rates)
    assert utils.read_minA(os.path.join(d, "min.A")) == [1, 2]
    assert utils.read_minA(os.path.join(d, "min.B")) == [40, 39, 38]


def test_coo_view_and_dat_files_round_trip(tmp_path):
    from kmc_rates_b200 import utils
    d = str(tmp_path)
    _write_db(d, nmin=25, nts=80, seed=3)
    gen = utils.RatesFromPathsampleDB(os.path.join(d, "min.data"), os.path.join(d, "ts.data"), 0.5)
    n, src, dst, k, peq = gen.coo()
    assert n == 25 and src.size == dst.size == k.size == len(gen.rate_constants)
    assert {(int(u) + 1, int(v) + 1) for u, v in zip(src, dst)} == set(gen.rate_constants)
    utils.write_dat_files(gen.rate_constants, [1, 2], [25], gen.Peq, directory=d)
    rates, A, B, Peq = utils.read_dat_files(d)
    assert A == [1, 2] and B == [25]
    assert set(rates) == set(gen.rate_constants)
    for uv, kk in gen.rate_constants.items():
        assert abs(rates[uv] - kk) <= 1e-15 * kk
    assert abs(Peq[7] - gen.Peq[7]) <= 1e-15 * gen.Peq[7]


def test_log_sum2():
    from kmc_rates_b200.utils import log_sum2
    assert abs(log_sum2(-1000.0, -1001.0) - (-1000.0 + np.log1p(np.exp(-1.0)))) < 1e-12
    assert abs(log_sum2(2.0, 3.0) - np.log(np.exp(2.0) + np.exp(3.0))) < 1e-12


@pytest.mark.gpu
def test_pathsample_database_to_rates_on_gpu(tmp_path):
    """the whole chain a `pathsample_rates`-style driver runs: database -> rates -> NGT with weights = Peq"""
    from kmc_rates_b200 import NGT, utils
    from oracle.cpu_ngt import CpuNGT
    d = str(tmp_path)
    _write_db(d, nmin=60, nts=260, seed=5)
    rates, Peq, _norm = utils.make_rates(d, 0.3)
    A, B = utils.read_minA(os.path.join(d, "min.A")), utils.read_minA(os.path.join(d, "min.B"))
    ngt = NGT(rates, A, B, weights=Peq)
    ngt.compute_rates()
    gen = utils.RatesFromPathsampleDB(os.path.join(d, "min.data"), os.path.join(d, "ts.data"), 0.3)
    n, src, dst, k, peq = gen.coo()
    cpu = CpuNGT(n, src, dst, k, [a - 1 for a in A], [b - 1 for b in B], weights={i: float(p) for i, p in enumerate(peq)})
    cpu.compute_rates()
    r = cpu.rates()
    assert abs(ngt.get_rate_AB() - r[0]) <= 1e-9 * r[0] and abs(ngt.get_rate_BA() - r[1]) <= 1e-9 * r[1]
    arr = NGT.from_coo(n, src, dst, k, [a - 1 for a in A], [b - 1 for b in B], weights={i: float(p) for i, p in enumerate(peq)})
    arr.compute_rates()
    assert abs(arr.get_rate_AB() - r[0]) <= 1e-9 * r[0]


# ---- pinned against outputs of the reference's own utils.py (tests/golden/pathsample_golden.json, made by
# ---- oracle/make_golden_utils.py) -----------------------------------------------------------------------------
def _pathsample_golden():
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "pathsample_golden.json")) as f:
        return json.load(f)


def _write_case(d, case):
    for name, key in (("min.data", "min_data"), ("ts.data", "ts_data"), ("min.A", "min_A")):
        with open(os.path.join(d, name), "w") as f:
            f.write(case[key])


def _check_against_golden(case, rates, Peq, norm, rtol):
    want = {(u, v): float(k) for u, v, k in case["rate_constants"]}
    assert set(rates) == set(want)
    for uv, k in want.items():
        assert abs(rates[uv] - k) <= rtol * k, (case["name"], uv, rates[uv], k)
    for u, p in case["Peq"]:
        assert abs(Peq[u] - float(p)) <= 1e-14 * max(float(p), 1e-300)
    assert abs(norm - float(case["rate_norm"])) <= 1e-14 * float(case["rate_norm"])


@pytest.mark.parametrize("idx", range(4))
def test_oracle_restatement_reproduces_the_reference_outputs(tmp_path, idx):
    from oracle.py_utils import rates_from_pathsample
    case = _pathsample_golden()["cases"][idx]
    _write_case(str(tmp_path), case)
    rates, Peq, norm = rates_from_pathsample(str(tmp_path / "min.data"), str(tmp_path / "ts.data"), case["T"])
    _check_against_golden(case, rates, Peq, norm, rtol=0.0 if idx < 4 else 1e-15)      # same loop, same order: bit-exact


@pytest.mark.parametrize("idx", range(4))
def test_host_ingest_reproduces_the_reference_outputs(tmp_path, idx):
    from kmc_rates_b200 import utils
    case = _pathsample_golden()["cases"][idx]
    _write_case(str(tmp_path), case)
    rates, Peq, norm = utils.make_rates(str(tmp_path), case["T"])
    _check_against_golden(case, rates, Peq, norm, rtol=1e-13)      # parallel transition states are summed in another order
    assert utils.read_minA(str(tmp_path / "min.A")) == case["min_A_ids"]


def test_log_sum2_golden():
    from kmc_rates_b200 import utils
    for a, b, want in _pathsample_golden()["log_sum2"]:
        assert utils.log_sum2(float(a), float(b)) == float(want)


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(4))
def test_device_ingest_reproduces_the_reference_outputs(tmp_path, idx):
    """csrc/pathsample.cu through the C ABI (ngt_pathsample_rates) against the reference's own outputs"""
    from kmc_rates_b200 import utils
    case = _pathsample_golden()["cases"][idx]
    _write_case(str(tmp_path), case)
    rates, Peq, norm = utils.make_rates(str(tmp_path), case["T"], device=0)
    _check_against_golden(case, rates, Peq, norm, rtol=1e-13)
    host = utils.RatesFromPathsampleDB(str(tmp_path / "min.data"), str(tmp_path / "ts.data"), case["T"])
    dev = utils.RatesFromPathsampleDB(str(tmp_path / "min.data"), str(tmp_path / "ts.data"), case["T"], device=0)
    n0, s0, d0, k0, p0 = host.coo()
    n1, s1, d1, k1, p1 = dev.coo()
    assert n0 == n1 and np.array_equal(s0, s1) and np.array_equal(d0, d1) and np.allclose(k0, k1, rtol=1e-13, atol=0)

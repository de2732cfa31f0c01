"""Host-side symbolic analysis without a GPU: the ordering and the whole front structure must not depend on the number
of worker threads (nested dissection branches, symbolic merge of dissection subtrees, relative maps all run threaded),
and must be repeatable.  Uses kmc_rates_b200/csrc/tools/analysis_bench.cpp, built here with g++."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "kmc_rates_b200", "csrc")


@pytest.fixture(scope="module")
def analysis_bench(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tools") / "analysis_bench")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", CSRC, os.path.join(CSRC, "tools", "analysis_bench.cpp"),
                           os.path.join(CSRC, "symbolic.cpp"), "-o", out])
    return out


def _dump(path, n, src, dst, A, B):
    with open(path, "wb") as f:
        np.array([n, src.size, len(A), len(B)], np.int64).tofile(f)
        np.asarray(src, np.int64).tofile(f)
        np.asarray(dst, np.int64).tofile(f)
        np.asarray(A, np.int64).tofile(f)
        np.asarray(B, np.int64).tofile(f)


def _hashes(exe, graph, env_extra, reps=2):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([exe, graph, str(reps)], env=env, check=True, capture_output=True, text=True).stdout
    return [line.split()[-1] for line in out.splitlines() if "structure hash" in line]


@pytest.mark.parametrize("side", [40, 150])
def test_analysis_is_independent_of_thread_count(analysis_bench, tmp_path, side):
    sys.path.insert(0, ROOT)
    from kmc_rates_b200 import synth
    n, src, dst, _k, A, B = synth.landscape_2d(side, T=0.3, seed=1)
    graph = str(tmp_path / "g.bin")
    _dump(graph, n, src, dst, A, B)
    serial = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "1", "NGT_B200_ND_THREADS_DEPTH": "0"})
    threaded = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "8", "NGT_B200_ND_THREADS_DEPTH": "4"})
    assert len(serial) == 2 and len(threaded) == 2
    assert len(set(serial + threaded)) == 1, (serial, threaded)
    # the sweeps have a scalar and (on hosts that support it) an AVX-512 version: same visiting order
    scalar = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "4", "NGT_B200_NO_AVX512": "1"}, reps=1)
    assert scalar[0] == serial[0], (scalar, serial)


def test_analysis_of_disconnected_and_irregular_graph(analysis_bench, tmp_path):
    """components of very different sizes, a clique, a long chain: same structure serial and threaded"""
    rng = np.random.default_rng(3)
    edges = []
    base = 0
    for size in (3000, 700, 40, 5):                      # random sparse components
        for i in range(1, size):
            j = int(rng.integers(max(0, i - 50), i))
            edges.append((base + i, base + j))
        for _ in range(2 * size):
            u, v = rng.integers(0, size, 2)
            if u != v:
                edges.append((base + int(u), base + int(v)))
        base += size
    for i in range(60):                                  # a clique
        for j in range(i):
            edges.append((base + i, base + j))
    base += 60
    for i in range(1, 2000):                             # a chain
        edges.append((base + i, base + i - 1))
    base += 2000
    e = np.array(edges, np.int64)
    src = np.concatenate([e[:, 0], e[:, 1]])
    dst = np.concatenate([e[:, 1], e[:, 0]])
    graph = str(tmp_path / "g.bin")
    _dump(graph, base, src, dst, [0], [1])
    serial = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "1", "NGT_B200_ND_THREADS_DEPTH": "0"}, reps=1)
    threaded = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "8", "NGT_B200_ND_THREADS_DEPTH": "4"}, reps=2)
    assert len(set(serial + threaded)) == 1, (serial, threaded)
    # the same through the accelerator interface (stand-in on the host): components, the clique and the chain are
    # declined or handed back, the structure stays the same
    for mode in ("1", "2"):
        accel = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "4", "NGT_ANALYSIS_MOCK_ACCEL": mode}, reps=1)
        assert accel == serial, (mode, accel, serial)
    # scalar sweeps and row filters against the AVX-512 ones (rows of every length from 1 to 59 and beyond)
    scalar = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "4", "NGT_B200_NO_AVX512": "1"}, reps=1)
    assert scalar == serial, (scalar, serial)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_analysis_does_not_depend_on_who_sweeps(analysis_bench, tmp_path, mode):
    """the dissector hands the sweeps of large subgraphs to an accelerator (the CUDA kernel of csrc/ordering.cu in the
    product; here a host stand-in behind the same interface, mode 2 declining every third call): same structure"""
    sys.path.insert(0, ROOT)
    from kmc_rates_b200 import synth
    n, src, dst, _k, A, B = synth.landscape_2d(90, T=0.3, seed=2)
    graph = str(tmp_path / "g.bin")
    _dump(graph, n, src, dst, A, B)
    host = _hashes(analysis_bench, graph, {"NGT_B200_HOST_THREADS": "4"}, reps=1)
    env = dict(os.environ, NGT_B200_HOST_THREADS="4", NGT_ANALYSIS_MOCK_ACCEL=mode)
    out = subprocess.run([analysis_bench, graph, "1"], env=env, check=True, capture_output=True, text=True).stdout
    accel = [line.split()[-1] for line in out.splitlines() if "structure hash" in line]
    offered = [int(line.split()[2]) for line in out.splitlines() if "sweeps offered" in line]
    assert offered and offered[0] >= 4, out
    assert host == accel, (host, accel)

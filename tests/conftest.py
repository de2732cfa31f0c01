import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_oracle_built():
    so = os.path.join(ROOT, "oracle", "libngt_oracle.so")
    src = os.path.join(ROOT, "oracle", "ngt_oracle.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


_ensure_oracle_built()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "ngt_golden.json")) as f:
        return json.load(f)["cases"]


def golden_cases():
    with open(os.path.join(ROOT, "tests", "golden", "ngt_golden.json")) as f:
        return json.load(f)["cases"]


def golden_input(case):
    """Rebuild the COO input of a golden case and verify its checksum."""
    from oracle.make_golden import build_input, digest
    n, src, dst, k, A, B = build_input(case)
    assert digest(src, dst, k) == case["sha256"], "synthetic generator drifted for %s" % case["name"]
    assert [int(a) for a in A] == case["resolved_A"] and [int(b) for b in B] == case["resolved_B"]
    w = case.get("weights")
    w = {int(a): float(b) for a, b in w.items()} if w else None
    return n, src, dst, k, list(A), list(B), w


def rel_err(got, want):
    got, want = np.asarray(got, float), np.asarray(want, float)
    denom = np.maximum(np.abs(want), 1e-300)
    return float(np.max(np.abs(got - want) / denom)) if got.size else 0.0

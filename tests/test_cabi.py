"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
fails loudly without a GPU, and the host-side symbolic analysis is sane."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    from kmc_rates_b200 import _capi
    L = _capi.lib()
    header = open(os.path.join(ROOT, "include", "ngt_b200.h")).read()
    declared = set(re.findall(r"\b(ngt_[a-z0-9_]+)\s*\(", header))
    declared -= {"ngt_handle_s"}
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_capi.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in L.ngt_version()


def test_no_cpu_fallback():
    """without a CUDA device construction must raise, never compute on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _ensure_built()
    from kmc_rates_b200 import NGT, synth
    rates = synth.to_dict(*synth.complete_deterministic(6))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        NGT(rates, [0], [1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kmc_rates_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, "%s mentions the oracle" % f


def test_input_validation_happens_before_cuda():
    _ensure_built()
    from kmc_rates_b200 import _capi
    src = np.array([0, 1, 1, 2], np.int64)
    dst = np.array([1, 0, 2, 1], np.int64)
    k = np.ones(4)
    with pytest.raises(ValueError, match="in common"):
        _capi.Handle.from_coo(3, src, dst, k, [0, 1], [1])
    with pytest.raises(ValueError, match="not in the graph"):
        _capi.Handle.from_coo(3, src, dst, k, [0], [7])
    with pytest.raises(ValueError, match="out of range"):
        _capi.Handle.from_coo(2, src, dst, k, [0], [1])

"""kmc_rates_b200 — B200-native New Graph Transformation (NGT) behind the kmc_rates API.

Drop-in names (reference kmc_rates/__init__.py:1-3 and kmc_rates/ngt.pyx):

    from kmc_rates_b200 import GraphReduction, kmcgraph_from_rates, NGT

Everything numeric runs in hand-written sm_100a CUDA kernels behind the C ABI of include/ngt_b200.h
(kmc_rates_b200/csrc/libngt_b200.so).  There is no CPU fallback: constructing a reducer without the built
library raises ImportError, and without a CUDA device RuntimeError.
"""
from .graph_reduction import GraphReduction, kmcgraph_from_rates
from .ngt import NGT, BaseNGT
from .rates_linalg import TwoStateRates

__all__ = ["GraphReduction", "kmcgraph_from_rates", "NGT", "BaseNGT", "TwoStateRates"]

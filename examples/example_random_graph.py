"""BASELINE config 1: random complete rate graph (reference examples/example_random_graph.py, README.rst:20-33), on B200.

    python examples/example_random_graph.py [nnodes]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from kmc_rates_b200 import GraphReduction, kmcgraph_from_rates


def main(nnodes=100):
    rng = np.random.default_rng(0)
    rates = {(i, j): rng.random() for i in range(nnodes) for j in range(nnodes) if i != j}
    A, B = [0], [1]
    reducer = GraphReduction(kmcgraph_from_rates(rates), A, B)      # a graph or the rates dict both work
    x = 2
    PxB = reducer.compute_committor_probability(x)
    print("committor probability %d -> %s before %s: %.12g" % (x, B, A, PxB))
    reducer.compute_rates()
    print("the transition rate from nodes", A, "to nodes", B, "is", reducer.get_rate_AB())
    print("the transition rate from nodes", B, "to nodes", A, "is", reducer.get_rate_BA())
    # the multi-state variant of the reference example
    reducer = GraphReduction(rates, [0, 1], [4, 5])
    reducer.compute_rates()
    print("A=[0,1] B=[4,5]: k_AB = %.12g  k_BA = %.12g  k_AB^SS = %.12g" % (reducer.get_rate_AB(), reducer.get_rate_BA(), reducer.get_rate_AB_SS()))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 100)

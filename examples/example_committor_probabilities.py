"""Committor map of a 2-D lattice with tuple node labels (reference examples/example_committor_probabilities.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networkx as nx
import numpy as np

from kmc_rates_b200 import GraphReduction


def main(dim=(10, 10)):
    rng = np.random.default_rng(0)
    grid = nx.grid_graph(list(dim))
    A, B = [(0, 0)], [(dim[0] - 1, dim[1] - 1)]
    rates = {}
    for u, v in grid.edges():
        rates[(u, v)] = rng.random()
        rates[(v, u)] = rng.random()
    reducer = GraphReduction(rates, A, B)
    com = reducer.compute_committor_probabilities(list(grid.nodes()))
    P = np.zeros(dim)
    for (x, y), p in com.items():
        P[x, y] = p
    np.set_printoptions(precision=3, suppress=True, linewidth=160)
    print("probability of reaching", B[0], "before", A[0])
    print(P)


if __name__ == "__main__":
    main()

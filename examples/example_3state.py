"""Three-state network: rates and committor by graph transformation on the GPU (twin of the reference's
examples/example_3state.py).  The reference cross-checks against its stochastic KMC simulator, which is outside this
repository's scope; the cross-check here is the linear-solve twin `TwoStateRates`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmc_rates_b200 import GraphReduction, kmcgraph_from_rates
from kmc_rates_b200.rates_linalg import TwoStateRates


def main():
    nnodes = 3
    transition_matrix = [[0., 1., 1.],
                         [1., 0., 1.],
                         [1., 1., 0.]]
    print("Computing rates and committor probabilities for transition matrix")
    print(transition_matrix)
    rates = {(i, j): transition_matrix[i][j] for i in range(nnodes) for j in range(nnodes) if i != j}

    A, B, x = [0], [1], 2          # rates from node 0 to node 1; committor of the intermediate node 2
    reducer = GraphReduction(rates, A, B)
    assert kmcgraph_from_rates(rates).number_of_nodes() == nnodes

    print("\ncomputing the probability for a trajectory to start at", x, "and reach", B, "before reaching", A)
    PxB = reducer.compute_committor_probability(x)
    print("the committor probability computed by graph transformation is", PxB)

    print("\ncomputing the transition rate from nodes", A, "to nodes", B)
    print("the exact rate for this network is 1.0")
    reducer.compute_rates()
    print("the transition rate computed by graph transformation is", reducer.get_rate_AB())

    la = TwoStateRates(rates, A, B)
    la.compute_rates()
    la.compute_committors()
    print("linear-solve cross-check: rate", la.get_rate_AB(), " committor of", x, "=", la.get_committor(x))


if __name__ == "__main__":
    main()

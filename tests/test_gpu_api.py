"""GPU tests of the drop-in Python API; they read like the reference's own tests
(reference kmc_rates/tests/test_graph_transformation.py, test_cpp_ngt.py, test_kmc.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _three_state_rates():
    return {(i, j): 1.0 for i in range(3) for j in range(3) if i != j}


@pytest.mark.parametrize("i,j", [(0, 1), (1, 2), (0, 2)])
def test_graph_reduction_three_state(i, j):
    """reference test_graph_transformation.py:73-98"""
    from kmc_rates_b200 import GraphReduction
    reducer = GraphReduction(_three_state_rates(), [i], [j], debug=True)
    reducer.check_graph()
    reducer.compute_rates()
    rAB = reducer.get_rate_AB()
    rBA = reducer.get_rate_BA()
    reducer.check_graph()
    assert reducer.graph.number_of_nodes() == 2
    assert reducer.graph.number_of_edges() == 4
    assert abs(rAB - 1.0) < 1e-7 and abs(rBA - 1.0) < 1e-7
    assert abs(reducer.get_rate_AB_SS() - 1.5) < 1e-7


@pytest.mark.parametrize("A,B", [([0], [1]), ([0, 1, 2], [3]), ([0, 1, 2], [3, 4, 5, 6])])
def test_graph_reduction_random(A, B):
    """reference test_graph_transformation.py:100-124, with the oracle as the value check"""
    from kmc_rates_b200 import GraphReduction, synth
    from oracle.py_ngt import PyGraphReduction
    rates = synth.random_connected_graph(20, 20, A + B, seed=len(A) + 10 * len(B))
    reducer = GraphReduction(rates, A, B)
    reducer.check_graph()
    reducer.compute_rates()
    reducer.check_graph()
    assert reducer.graph.number_of_nodes() == len(A) + len(B)
    if len(A) == 1 and len(B) == 1:
        assert reducer.graph.number_of_edges() <= 4
    ref = PyGraphReduction(rates, A, B)
    ref.compute_rates()
    assert abs(reducer.get_rate_AB() - ref.get_rate_AB()) < 1e-9 * ref.get_rate_AB()
    assert abs(reducer.get_rate_BA() - ref.get_rate_BA()) < 1e-9 * ref.get_rate_BA()
    assert abs(reducer.get_rate_AB_SS() - ref.get_rate_AB_SS()) < 1e-9 * ref.get_rate_AB_SS()
    for a in A:
        assert abs(reducer.get_committor_probabilityAB(a) - ref.get_committor_probabilityAB(a)) < 1e-9


def test_graph_reduction_committors_weights_and_groups():
    """reference test_kmc.py:74-115: groups, weights, committors of nodes inside and outside A ∪ B"""
    from kmc_rates_b200 import GraphReduction, synth
    from oracle.py_ngt import PyGraphReduction
    A, B = [0, 1, 2, 3], [8, 9]
    rates = synth.random_connected_graph(10, 20, A + B + [5], seed=3)
    rng = np.random.RandomState(0)
    weights = {x: rng.uniform(0, 1) for x in A + B}
    reducer = GraphReduction(rates, A, B, weights=weights)
    ref = PyGraphReduction(rates, A, B, weights=weights)
    nodes = set(A + B + [5])
    got = reducer.compute_committor_probabilities(nodes)
    want = ref.compute_committor_probabilities(nodes)
    for x in nodes:
        assert abs(got[x] - want[x]) < 1e-9
    assert abs(reducer.compute_committor_probability(5) - want[5]) < 1e-9
    reducer.compute_rates()
    ref.compute_rates()
    assert abs(reducer.get_rate_AB() - ref.get_rate_AB()) < 1e-9 * ref.get_rate_AB()
    assert abs(reducer.get_rate_BA() - ref.get_rate_BA()) < 1e-9 * ref.get_rate_BA()


def test_graph_reduction_accepts_graph_and_tuple_labels():
    """README passes a graph (reference README.rst:28-33); the lattice example uses tuple labels
    (reference examples/example_committor_probabilities.py:8-21)"""
    import networkx as nx
    from kmc_rates_b200 import GraphReduction, kmcgraph_from_rates
    from oracle.py_ngt import PyGraphReduction
    rng = np.random.RandomState(1)
    grid = nx.grid_graph([5, 5])
    rates = {}
    for u, v in grid.edges():
        rates[(u, v)] = rng.rand()
        rates[(v, u)] = rng.rand()
    A, B = [(0, 0)], [(4, 4)]
    reducer = GraphReduction(kmcgraph_from_rates(rates), A, B)
    com = reducer.compute_committor_probabilities(list(grid.nodes()))
    ref = PyGraphReduction(rates, A, B)
    want = ref.compute_committor_probabilities(list(grid.nodes()))
    for x in grid.nodes():
        assert abs(com[x] - want[x]) < 1e-9
    reducer.compute_rates()
    ref.compute_rates()
    assert abs(reducer.get_rate_AB() - ref.get_rate_AB()) < 1e-9 * ref.get_rate_AB()


def test_graph_reduction_errors():
    """error conventions of reference _kmc_rates.py:148-154, 434-447"""
    from kmc_rates_b200 import GraphReduction, synth
    rates = synth.to_dict(*synth.complete_deterministic(5))
    with pytest.raises(ValueError):
        GraphReduction(rates, [0, 1], [1, 2])
    with pytest.raises(ValueError):
        GraphReduction(rates, [0], [17])
    red = GraphReduction(rates, [0], [1])
    with pytest.raises(RuntimeError):
        red.get_committor_probabilityAB(0)
    red.compute_rates()
    with pytest.raises(ValueError):
        red.get_committor_probabilityAB(3)
    # disconnected A and B
    r2 = {(0, 1): 1.0, (1, 0): 1.0, (2, 3): 1.0, (3, 2): 1.0}
    with pytest.raises(ValueError):
        GraphReduction(r2, [0], [2])


def test_ngt_class_matches_reference_tests():
    """reference kmc_rates/tests/test_cpp_ngt.py:37-63 (NGT10 through the NGT class, dict input)"""
    from kmc_rates_b200 import NGT, synth
    rates = synth.to_dict(*synth.complete_deterministic(10))
    reducer = NGT(rates, [0, 1, 2], [3, 4, 5], debug=False)
    reducer.compute_rates()
    assert abs(reducer.get_rate_AB() - 5.1013138820442565) < 1e-7
    assert abs(reducer.get_rate_BA() - 3) < 1e-7
    assert abs(reducer.get_rate_AB_SS() - 19.933950145409426) < 1e-7
    assert abs(reducer.get_rate_BA_SS() - 6.970856553435547) < 1e-7
    assert reducer.time_solve > 0


@pytest.mark.parametrize("A,B,weighted", [([0], [1], False), ([0, 1], [2, 3], False), ([0, 1], [2, 3], True), (list(range(8)), [9], False)])
def test_two_state_rates_twin_matches_linear_solve(A, B, weighted):
    """GPU twin of rates_linalg.TwoStateRates against the restated scipy version (reference tests/test_kmc.py:40-59,
    tests/test_linalg.py:12-21)"""
    from kmc_rates_b200 import TwoStateRates, synth
    from oracle.py_linalg import TwoStateRatesCOO
    rates = synth.random_connected_graph(10, 20, A + B, seed=11 + len(A))
    rng = np.random.RandomState(3)
    weights = {x: rng.uniform(0, 1) for x in A} if weighted else None
    lin = TwoStateRates(rates, A, B, weights=weights)
    lin.compute_rates()
    lin.compute_committors()
    nodes = sorted({u for uv in lin.rate_constants for u in uv})
    idx = {u: i for i, u in enumerate(nodes)}
    src = np.array([idx[u] for (u, v) in lin.rate_constants]); dst = np.array([idx[v] for (u, v) in lin.rate_constants])
    k = np.array(list(lin.rate_constants.values()))
    ref = TwoStateRatesCOO(len(nodes), src, dst, k, [idx[a] for a in A], [idx[b] for b in B],
                           weights={idx[a]: w for a, w in weights.items()} if weights else None)
    ref.compute_rates(); ref.compute_committors()
    assert abs(lin.get_rate_AB() - ref.get_rate_AB()) < 1e-9 * ref.get_rate_AB()
    assert abs(lin.get_rate_AB_SS() - ref.get_rate_AB_SS()) < 1e-9 * ref.get_rate_AB_SS()
    for u, t in lin.mfptimes.items():
        assert abs(t - ref.mfpt[idx[u]]) < 1e-9 * ref.mfpt[idx[u]]
    for u, q in lin.committor_dict.items():
        assert abs(q - ref.q[idx[u]]) < 1e-9
    assert lin.get_committor(A[0]) == 0.0 and lin.get_committor(B[0]) == 1.0


def test_two_state_rates_three_state():
    """reference tests/test_linalg.py:6-28: k = 1.0, k_SS = 1.5"""
    from kmc_rates_b200 import TwoStateRates
    for i, j in ((0, 1), (1, 2), (0, 2)):
        red = TwoStateRates(_three_state_rates(), [i], [j])
        red.compute_rates()
        assert abs(red.get_rate_AB() - 1.0) < 1e-7
        red.compute_committors()
        assert abs(red.get_rate_AB_SS() - 1.5) < 1e-7


def test_graph_reduction_node_that_cannot_reach_the_other_group():
    """0 can only shuttle between itself and 2: after phase 2 nothing but its self loop is left, and the reference
    raises a bare Exception (reference _kmc_rates.py:227-228)"""
    from kmc_rates_b200 import GraphReduction
    rates = {(1, 0): 1.0, (0, 2): 1.0, (2, 0): 1.0, (1, 2): 1.0}
    reducer = GraphReduction(rates, [0], [1])
    with pytest.raises(Exception, match="node 0 is not connected"):
        reducer.compute_rates()


def test_graph_reduction_weights_covering_only_A():
    """the reference reads weights lazily per group (reference _kmc_rates.py:168-172): a plain dict over A is enough for
    get_rate_AB, and get_rate_BA then fails with KeyError"""
    from kmc_rates_b200 import GraphReduction, synth
    from oracle.py_ngt import PyGraphReduction
    A, B = [0, 1, 2], [7, 8]
    rates = synth.random_connected_graph(12, 24, A + B, seed=5)
    w = {0: 0.2, 1: 0.5, 2: 0.3}
    reducer = GraphReduction(rates, A, B, weights=w)
    reducer.compute_rates()
    ref = PyGraphReduction(rates, A, B, weights=w)
    ref.compute_rates()
    assert abs(reducer.get_rate_AB() - ref.get_rate_AB()) < 1e-9 * ref.get_rate_AB()
    with pytest.raises(KeyError):
        reducer.get_rate_BA()


def test_two_state_rates_with_absorbing_product_nodes():
    """B without outgoing rates: the reference's linear-algebra classes treat B as absorbing and never read rates out
    of it (reference rates_linalg.py:246-285)"""
    from kmc_rates_b200 import synth
    from kmc_rates_b200.rates_linalg import TwoStateRates
    from oracle.py_linalg import TwoStateRatesCOO
    A, B = [0, 1], [8, 9]
    rates = synth.random_connected_graph(10, 24, A + B, seed=11)
    rates = {uv: r for uv, r in rates.items() if uv[0] not in B}      # products become pure sinks
    assert any(v in B for (_u, v) in rates)
    twin = TwoStateRates(rates, A, B)
    twin.compute_rates()
    twin.compute_committors()
    labels = sorted({u for uv in rates for u in uv})
    idx = {u: i for i, u in enumerate(labels)}
    src = np.array([idx[u] for (u, _v) in rates]); dst = np.array([idx[v] for (_u, v) in rates])
    k = np.array(list(rates.values()))
    la = TwoStateRatesCOO(len(labels), src, dst, k, [idx[a] for a in A], [idx[b] for b in B])
    la.compute_rates()
    la.compute_committors()
    assert abs(twin.get_rate_AB() - la.get_rate_AB()) < 1e-9 * la.get_rate_AB()
    for u in labels:
        if u not in A and u not in B:
            assert abs(twin.get_committor(u) - la.q[idx[u]]) < 1e-9


def test_connected_components_kernel_matches_scipy():
    """graph validation on the device: union-find labels against scipy's connected components"""
    import scipy.sparse
    import scipy.sparse.csgraph
    from kmc_rates_b200 import _capi
    rng = np.random.default_rng(7)
    for n, nnz in [(1, 0), (5, 0), (50, 30), (2000, 1500), (20000, 60000), (300000, 280000)]:
        src = rng.integers(0, n, nnz)
        dst = rng.integers(0, n, nnz)
        nc, lab = _capi.connected_components(n, src, dst)
        adj = scipy.sparse.coo_matrix((np.ones(nnz), (src, dst)), shape=(n, n))
        nc_ref, lab_ref = scipy.sparse.csgraph.connected_components(adj, directed=False)
        assert nc == nc_ref
        # same partition, and the label is the smallest member
        first = {}
        for u, l in enumerate(lab_ref):
            first.setdefault(l, u)
        assert np.array_equal(lab, np.array([first[l] for l in lab_ref], np.int32))
    with pytest.raises(ValueError):
        _capi.connected_components(4, [0, 5], [1, 2])

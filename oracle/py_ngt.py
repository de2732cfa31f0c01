"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

Python-3 / networkx-3 restatement of the reference's pure-Python NGT
(reference kmc_rates/_kmc_rates.py, Python 2 + networkx 1, which cannot be imported here:
print statements :234, iteritems :48, graph.node[...] :62, out_edges_iter :180,
successors()+predecessors() list concat :334, connected_components as a list :451).

It is the semantic oracle for the parts of the drop-in API that the C++ engine does not have:
the static degree ordering, `1. - Pxx` (never the off-diagonal sum), the reduced `graph`
attribute, `compute_committor_probabilities`, and the error conventions.

Parity status: PINNED against the reference's golden values (tests/test_oracle.py):
3-state rates = 1.0 (reference kmc_rates/tests/test_graph_transformation.py:73-101), NGT10 values
(reference kmc_rates/tests/test_cpp_ngt.py:41-60) and, through those, the compiled reference engine.
"""
from collections import defaultdict

import networkx as nx


def kmcgraph_from_rates(rates):
    """reference _kmc_rates.py:18-66: tau_u = 1/sum_v k_uv, P_uu = 0, P_uv = k_uv tau_u."""
    total = defaultdict(float)
    for (u, _v), k in rates.items():
        total[u] += k
    g = nx.DiGraph()
    for u, s in total.items():
        g.add_node(u, tau=1.0 / s)
        g.add_edge(u, u, P=0.0)
    for (u, v), k in rates.items():
        g.add_edge(u, v, P=k * g.nodes[u]["tau"])
    return g


class PyGraphReduction(object):
    """reference _kmc_rates.py:70-546"""

    def __init__(self, rate_constants, A, B, debug=False, weights=None):
        self.graph = kmcgraph_from_rates(rate_constants)
        self.A, self.B = set(A), set(B)
        self.weights = defaultdict(lambda: 1.0) if weights is None else weights
        self._final_Pxx, self._final_tau = {}, {}
        self.debug = debug
        self.initial_check_graph()
        self.check_graph()
        self._initial_tau = {u: d["tau"] for u, d in self.graph.nodes(data=True)}

    # ---- elimination (reference :118-131, :274-355) ---------------------------------------
    def _degree(self, x):
        return self.graph.in_degree(x) + self.graph.out_degree(x)

    def _remove_nodes(self, nodes):
        order = sorted(nodes, key=self._degree)  # one static sort, reference :122
        for x in order:
            self._remove_node(x)

    def _remove_node(self, x):
        g = self.graph
        neibs = set(g.successors(x)) | set(g.predecessors(x))
        neibs.discard(x)
        tau_x = g.nodes[x]["tau"]
        Pxx = g[x][x]["P"]
        om = 1.0 - Pxx  # reference :302,325 — never the off-diagonal sum
        for u in neibs:
            if g.has_edge(u, x):  # reference :316 would KeyError on a pure successor; guard = no-op there
                g.nodes[u]["tau"] += g[u][x]["P"] * tau_x / om
        for u in neibs:
            if not g.has_edge(u, x):
                continue
            Pux = g[u][x]["P"]
            for v in neibs:
                if not g.has_edge(x, v):
                    continue
                if not g.has_edge(u, v):
                    g.add_edge(u, v, P=0.0)
                g[u][v]["P"] += Pux * g[x][v]["P"] / om
        g.remove_node(x)

    def _phase_one_remove_intermediates(self):
        self._remove_nodes(set(self.graph.nodes()) - self.A - self.B)

    # ---- phase two (reference :188-263) ----------------------------------------------------
    def _reduce_all_iterator(self, nodes, restore_graph=True):
        if len(nodes) == 0:
            return
        full = self.graph.copy() if restore_graph else None
        nodes = sorted(nodes, key=self._degree)
        while True:
            if len(nodes) == 1:
                yield nodes[0]
                break
            saved = self.graph.copy()
            u = nodes.pop(0)
            self._remove_nodes(nodes)
            yield u
            self.graph = saved
            self._remove_node(u)
        if restore_graph:
            self.graph = full

    def _phase_two_group(self, group):
        for a in self._reduce_all_iterator(group):
            if self.graph.out_degree(a) <= 1:
                raise Exception("node %s is not connected" % (a,))
            self._final_Pxx[a] = self.graph[a][a]["P"]
            self._final_tau[a] = self.graph.nodes[a]["tau"]

    def _phase_two(self):
        full = self.graph.copy()
        self._phase_two_group(self.A)
        self._phase_two_group(self.B)
        self.graph = full

    def compute_rates(self):
        self._phase_one_remove_intermediates()
        self._phase_two()

    # ---- results (reference :133-186) ------------------------------------------------------
    def _get_final_rate(self, group):
        rate = sum((1.0 - self._final_Pxx[x]) / self._final_tau[x] * self.weights[x] for x in group)
        return rate / sum(self.weights[x] for x in group)

    def get_rate_AB(self):
        return self._get_final_rate(self.A)

    def get_rate_BA(self):
        return self._get_final_rate(self.B)

    def get_rate_AB_SS(self):
        rate = 0.0
        for a in self.A:
            PaB = sum(d["P"] for _a, b, d in self.graph.out_edges(a, data=True) if b in self.B)
            rate += PaB * self.weights[a] / self._initial_tau[a]
        return rate / sum(self.weights[x] for x in self.A)

    def get_committor_probabilityAB(self, x):
        if len(self._final_Pxx) == 0:
            raise RuntimeError("you must call compute_rates before calling this function")
        try:
            return 1.0 - self._final_Pxx[x]
        except KeyError:
            if x not in self.A and x not in self.B:
                raise ValueError("x is not in A or in B.  Use compute_committor_probability() if x is an intermediate")

    # ---- committors of arbitrary nodes (reference :468-546) -------------------------------
    def _get_committor_probability(self, x):
        PxA = sum(d["P"] for _u, v, d in self.graph.out_edges(x, data=True) if v in self.A)
        PxB = sum(d["P"] for _u, v, d in self.graph.out_edges(x, data=True) if v in self.B)
        if PxA + PxB == 0.0:
            raise Exception
        return PxB / (PxA + PxB)

    def compute_committor_probability(self, x):
        return self.compute_committor_probabilities([x])[x]

    def compute_committor_probabilities(self, nodes):
        nodes = set(nodes)
        if nodes - set(self.graph.nodes()):
            raise ValueError("At least on of the nodes is not in the graph.")
        backup = self.graph.copy()
        self._remove_nodes(set(self.graph.nodes()) - nodes - self.A - self.B)
        PxB = {}
        for x in self._reduce_all_iterator(nodes - self.A - self.B, restore_graph=False):
            PxB[x] = self._get_committor_probability(x)
        for x in set(self.graph.nodes()) - self.A - self.B:
            self._remove_node(x)
        for x in nodes & (self.A | self.B):
            PxB[x] = self._get_committor_probability(x)
        self.graph = backup
        return PxB

    # ---- validation (reference :371-466) ---------------------------------------------------
    def _check_node(self, u):
        assert self.graph.nodes[u]["tau"] >= 0
        total = 0.0
        for _x, _v, d in self.graph.out_edges(u, data=True):
            assert 1 >= d["P"] >= 0
            total += d["P"]
        assert abs(total - 1.0) < 1e-6, "%s: total_prob %g" % (str(u), total)

    def check_graph(self):
        for u in self.graph.nodes():
            self._check_node(u)

    def _check_A_B_connected(self, components):
        comps = [set(c) for c in components]
        inA = [c & self.A for c in comps]
        inB = [c & self.B for c in comps]
        if len([c for c in inA if c]) != 1:
            raise ValueError("the reactant set (A) is not fully connected")
        if len([c for c in inB if c]) != 1:
            raise ValueError("the product set (B) is not fully connected")
        if not any(a and b for a, b in zip(inA, inB)):
            raise ValueError("product and reactant sets (A and B) are not connected")
        stray = set()
        for c in comps:
            if not (c & self.A) and not (c & self.B):
                stray |= c
        if stray:
            self.graph.remove_nodes_from(stray)

    def initial_check_graph(self):
        for a in self.A:
            if not self.graph.has_node(a):
                raise ValueError("an element in the reactant set (A) is not in the graph")
        for b in self.B:
            if not self.graph.has_node(b):
                raise ValueError("an element in the product set (B) is not in the graph")
        for u in list(self.graph.nodes()):
            if not self.graph.has_edge(u, u):
                self.graph.add_edge(u, u, P=0.0)
        if self.A & self.B:
            raise ValueError("A and B have at least one node in common")
        cc = list(nx.connected_components(self.graph.to_undirected()))
        if len(cc) != 1:
            self._check_A_B_connected(cc)

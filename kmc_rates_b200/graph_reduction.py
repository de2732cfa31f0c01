"""`GraphReduction` and `kmcgraph_from_rates`: drop-in for the reference's exported Python API
(reference kmc_rates/_kmc_rates.py, kmc_rates/__init__.py:1-3), with the arithmetic on B200.

What stays on the host (as in the reference): label handling, the per-node sanity checks and the final weighted
averages.  What moves to the GPU: every P / tau update (reference _kmc_rates.py:274-355), phase 1, phase 2, the
committor solves, and the connectivity analysis behind the A / B checks and the pruning of components that touch
neither A nor B (reference _kmc_rates.py:384-454; a union-find kernel, `_capi.connected_components`).
"""
from collections import defaultdict

import networkx as nx
import numpy as np

from . import _capi
from .ngt import _label_maps, rates_to_coo


def kmcgraph_from_rates(rates):
    """Build the kinetic graph of a rate dictionary {(u, v): k_uv}.

    Returns a networkx.DiGraph whose nodes carry "tau" = 1 / sum_v k_uv (mean waiting time) and whose edges carry
    "P" = k_uv * tau_u (branching probability); every node also gets a self loop with P = 0, the slot the
    transformation later fills (same construction as reference _kmc_rates.py:18-66).
    """
    graph = nx.DiGraph()
    sumk = defaultdict(float)
    for (u, _v), rate in rates.items():
        sumk[u] += rate
    for u, s in sumk.items():
        graph.add_node(u, tau=1. / s)
        graph.add_edge(u, u, P=0.)
    for (u, v), rate in rates.items():
        graph.add_edge(u, v, P=rate * graph.nodes[u]["tau"])
    return graph


def _rates_from_graph(graph):
    """invert kmcgraph_from_rates: k_uv = P_uv / tau_u (the README passes a graph, reference README.rst:28-33)"""
    rates = {}
    for u, v, data in graph.edges(data=True):
        p = data["P"]
        if u == v and p == 0.:
            continue
        rates[(u, v)] = p / graph.nodes[u]["tau"]
    return rates


class GraphReduction(object):
    """Drop-in for the reference's GraphReduction (reference _kmc_rates.py:70-116) with the arithmetic on B200.

    rate_constants : dict {(u, v): k_uv}, or a graph made by kmcgraph_from_rates
    A, B           : reactant and product groups (iterables of node labels)
    weights        : optional dict node -> equilibrium occupation probability for the final weighted averages
    debug          : accepted for compatibility

    After compute_rates(): get_rate_AB(), get_rate_BA(), get_rate_AB_SS(), get_committor_probabilityAB(x);
    compute_committor_probabilities(nodes) works before or after.  `graph` is the full kinetic graph before
    compute_rates() and the graph reduced to A ∪ B afterwards, as in the reference.
    """

    def __init__(self, rate_constants, A, B, debug=False, weights=None, device=0, gemm_path=0, leaf_size=0):
        if isinstance(rate_constants, nx.DiGraph):
            rate_constants = _rates_from_graph(rate_constants)
        self._rates = dict(rate_constants)
        self.A = set(A)
        self.B = set(B)
        if weights is None:
            self.weights = defaultdict(lambda: 1.)
        else:
            self.weights = weights
        self._final_Pxx = dict()
        self._final_tau = dict()
        self._final_omPxx = dict()
        self.debug = debug
        self._opt = dict(device=device, gemm_path=gemm_path, leaf_size=leaf_size)
        self._graph = None
        self._h = None
        self._computed = False
        self._committors = None

        self.initial_check_graph()
        self.check_graph()
        self._build_engine()
        self._initial_tau = dict(self._tau0)

    # ---- host-side structure ---------------------------------------------------------------------------
    def _build_engine(self):
        self._node2id = _label_maps(self._rates)
        self._nodes = list(self._node2id.keys())
        src, dst, k = rates_to_coo(self._rates, self._node2id)
        self._Aids = [self._node2id[a] for a in self.A]
        self._Bids = [self._node2id[b] for b in self.B]
        self._Alist, self._Blist = list(self.A), list(self.B)
        self._h = _capi.Handle.from_coo(len(self._nodes), src, dst, k, self._Aids, self._Bids, **self._opt)
        # the reference reads weights lazily, per group and only in the getter that needs them (_kmc_rates.py:168-172):
        # a plain dict that covers A alone is enough for get_rate_AB().  Push what exists; missing entries surface as
        # KeyError in the getter that asks for them, as in the reference.
        known = [x for x in self._Alist + self._Blist if isinstance(self.weights, defaultdict) or x in self.weights]
        if known:
            self._h.set_weights([self._node2id[x] for x in known], [self.weights[x] for x in known])

    @property
    def _tau0(self):
        sumk = defaultdict(float)
        for (u, _v), r in self._rates.items():
            sumk[u] += r
        return {u: 1. / s for u, s in sumk.items()}

    @property
    def graph(self):
        """Before compute_rates: the full kmc graph.  After: the graph reduced to A ∪ B
        (the side effect the reference's tests rely on, kmc_rates/tests/test_graph_transformation.py:86-87,110-112)."""
        if self._graph is None:
            if self._computed:
                self._graph = self._reduced_graph()
            else:
                self._graph = kmcgraph_from_rates(self._rates)
        return self._graph

    @graph.setter
    def graph(self, g):
        self._graph = g

    def _reduced_graph(self):
        P, tau = self._h.reduced()
        keep = self._Alist + self._Blist
        g = nx.DiGraph()
        for i, u in enumerate(keep):
            g.add_node(u, tau=float(tau[i]))
        for i, u in enumerate(keep):
            for j, v in enumerate(keep):
                if i == j or P[i, j] != 0.:
                    g.add_edge(u, v, P=float(P[i, j]))
        return g

    # ---- the reference's public methods --------------------------------------------------------------
    def compute_rates(self):
        """Run phase 1 and phase 2 on the GPU and cache 1-P'_xx and tau'_x of every x in A and B."""
        self._h.compute_rates()
        om, tau = self._h.final()
        for i, x in enumerate(self._Alist + self._Blist):
            if not om[i] > 0.:
                # x keeps no edge besides its self loop once the rest of its group is gone
                # (the reference's check, _kmc_rates.py:227-228)
                raise Exception("node %s is not connected" % (x,))
            self._final_omPxx[x] = float(om[i])
            self._final_Pxx[x] = 1. - float(om[i])
            self._final_tau[x] = float(tau[i])
        self._computed = True
        self._graph = None

    def get_rate_AB(self):
        """k_AB: the weighted mean over a in A of 1 / (mean first-passage time from a into B)."""
        return self._get_final_rate(self._Alist)

    def get_rate_BA(self):
        """k_BA: the same average taken over the members of B, for passages into A."""
        return self._get_final_rate(self._Blist)

    def _get_final_rate(self, group):
        rate = sum(self._final_omPxx[x] / self._final_tau[x] * self.weights[x] for x in group)
        norm = sum(self.weights[x] for x in group)
        return rate / norm

    def get_rate_AB_SS(self):
        """steady state rate sum_a w_a P_aB / tau_a^initial / sum_a w_a on the reduced graph"""
        if not self._computed:
            raise RuntimeError("you must call compute_rates before calling this function")
        P, _tau = self._h.reduced()
        nA = len(self._Alist)
        rate = 0.
        for i, a in enumerate(self._Alist):
            rate += P[i, nA:].sum() * self.weights[a] / self._initial_tau[a]
        return rate / sum(self.weights[x] for x in self._Alist)

    def get_committor_probabilityAB(self, x):
        """1 - P'_xx of an end-point node x after phase 2: for x in A the chance that a walker leaving x arrives in B
        before it is back at x; for x in B the same with A as the target.  Only defined for members of A or B."""
        if len(self._final_Pxx) == 0:
            raise RuntimeError("you must call compute_rates before calling this function")
        try:
            return self._final_omPxx[x]
        except KeyError:
            if x not in self.A and x not in self.B:
                raise ValueError("x is not in A or in B.  Use compute_committor_probability() if x is an intermediate")

    def compute_committor_probability(self, x):
        """Committor of a single node: the chance of hitting B before A when starting at x."""
        return self.compute_committor_probabilities([x])[x]

    def compute_committor_probabilities(self, nodes):
        """
        Committors (probability of reaching B before A) of all `nodes`, as a dict node -> value.

        The reference must be asked before compute_rates() because it destroys its graph
        (reference _kmc_rates.py:496-546); here the order does not matter.  Nodes of A ∪ B get
        P_xB / (P_xA + P_xB) read from the reduced graph, like the reference (:468-484, :538-541).
        """
        nodes = set(nodes)
        if nodes - set(self._node2id.keys()):
            raise ValueError("At least on of the nodes is not in the graph.")
        if self._committors is None:
            self._h.phase_one()
            self._committors = self._h.committors()
            self._reduced_for_committor = self._h.reduced()[0]
        P = self._reduced_for_committor
        keep = self._Alist + self._Blist
        nA = len(self._Alist)
        out = {}
        for x in nodes:
            if x in self.A or x in self.B:
                i = keep.index(x)
                PxA, PxB = P[i, :nA].sum(), P[i, nA:].sum()
                if PxA + PxB == 0.:
                    raise Exception("node %s has no path to A or B" % (x,))
                out[x] = float(PxB / (PxA + PxB))
            else:
                out[x] = float(self._committors[self._node2id[x]])
        return out

    # ---- validation (host; reference _kmc_rates.py:371-466) ------------------------------------------
    def _check_A_B_connected(self, labels, comp):
        compsA = {comp[labels[a]] for a in self.A}
        compsB = {comp[labels[b]] for b in self.B}
        if len(compsA) != 1:
            raise ValueError("the reactant set (A) is not fully connected")
        if len(compsB) != 1:
            raise ValueError("the product set (B) is not fully connected")
        if compsA != compsB:
            raise ValueError("product and reactant sets (A and B) are not connected")
        good = next(iter(compsA))
        before = len(self._rates)
        self._rates = {(u, v): r for (u, v), r in self._rates.items() if comp[labels[u]] == good}
        if self.debug and before != len(self._rates):
            print("removed rates of nodes that are not connected to A or to B")

    def initial_check_graph(self):
        labels = _label_maps(self._rates)
        has_out = {u for (u, _v) in self._rates.keys()}
        for a in self.A:
            if a not in has_out:
                raise ValueError("an element in the reactant set (A) is not in the graph")
        for b in self.B:
            if b not in has_out:
                raise ValueError("an element in the product set (B) is not in the graph")
        if len(self.A.intersection(self.B)) > 0:
            raise ValueError("A and B have at least one node in common")
        n = len(labels)
        src, dst, _k = rates_to_coo(self._rates, labels)
        # connected components by the union-find kernel (the reference sweeps the graph with networkx, :384-454)
        ncomp, comp = _capi.connected_components(n, src, dst, device=self._opt["device"])
        if ncomp != 1:
            self._check_A_B_connected(labels, comp)
            self._graph = None

    def _check_node_arrays(self, P_rows):
        for u, (tau, probs) in P_rows.items():
            assert tau >= 0
            total = 0.
            for p in probs:
                assert 1 >= p >= 0
                total += p
            assert np.abs(total - 1.) < 1e-6, "%s: total_prob %g" % (str(u), total)

    def check_graph(self):
        """tau >= 0, 0 <= P <= 1 and sum_v P_uv = 1 for every node (reference _kmc_rates.py:371-382, 460-466)"""
        if self._computed:
            g = self.graph
            rows = {u: (g.nodes[u]["tau"], [d["P"] for _u, _v, d in g.out_edges(u, data=True)]) for u in g.nodes()}
        else:
            sumk = defaultdict(float)
            per = defaultdict(list)
            for (u, _v), r in self._rates.items():
                sumk[u] += r
            for (u, _v), r in self._rates.items():
                per[u].append(r / sumk[u])
            rows = {u: (1. / sumk[u], per[u]) for u in sumk}
        self._check_node_arrays(rows)

// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement of the reference's New Graph Transformation (NGT) so that the
// CUDA path has something to be checked against on a box where /root/reference
// does not exist.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this
// file.  The product (kmc_rates_b200/) never links, imports or calls it.
//
// Parity status: PINNED.  tests/test_oracle.py checks this restatement against
//   (a) the reference's own golden values (cpp_tests/source/test_ngt.cpp:27-46,
//       68-93; kmc_rates/tests/test_cpp_ngt.py:14-60), and
//   (b) outputs of the unmodified reference engine compiled into oracle/_ref
//       (oracle/ref_shim.cpp), frozen as tests/golden/*.json by
//       oracle/make_golden.py.
//
// What is restated (reference file:line -> function here):
//   source/kmc_rates/ngt.hpp:105-181   NGT ctor: tau=1/sum k, P=k*tau, self loop   -> Oracle::Oracle
//   source/kmc_rates/ngt.hpp:192-200   dynamic min in+out-degree ordering         -> Net::pop_min_degree
//   source/kmc_rates/ngt.hpp:222-238   1-Pxx, by off-diagonal sum if Pxx>=0.99     -> Net::one_minus_P
//   source/kmc_rates/ngt.hpp:246-255   tau_u += Pux tau_x/(1-Pxx)                  -> Net::remove_node
//   source/kmc_rates/ngt.hpp:260-333   P_uv += Pux Pxv/(1-Pxx), fill-in, delete x  -> Net::remove_node
//   source/kmc_rates/ngt.hpp:338-354   phase one                                  -> Net::remove_all
//   source/kmc_rates/ngt.hpp:362-431   phase two (per-x copies)                   -> Oracle::reduce_group
//   source/kmc_rates/ngt.hpp:444-513   k_AB, k_BA, steady-state rates             -> Oracle::rate_final / rate_ss
//   source/kmc_rates/ngt.hpp:518-630   committors of all intermediates            -> Oracle::compute_rates_and_committors
//
// The data structure is NOT the reference's pointer graph (graph.hpp): nodes are
// dense ids, out-edges are an ordered map per node, in-neighbours a set per
// node.  Elimination order differs from the reference only in tie-breaking, so
// results agree to rounding (checked at rel 1e-11 in tests), not bitwise.
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <map>
#include <set>
#include <vector>
#include <string>
#include <stdexcept>
#include <algorithm>

namespace {

typedef int64_t id_t64;

struct Net {
    // out[u][v] = P_uv (includes v==u, the self loop); in[v] = {u : u->v exists, u != v}
    std::vector<std::map<id_t64, double> > out;
    std::vector<std::set<id_t64> > in;
    std::vector<double> tau;
    std::vector<char> alive;

    explicit Net(size_t n = 0) : out(n), in(n), tau(n, 0.0), alive(n, 0) {}

    size_t degree(id_t64 u) const {
        // ngt.hpp:29-31 / graph.hpp:104: in_degree + out_degree, the self loop counted in both
        return out[u].size() + in[u].size() + (out[u].count(u) ? 1 : 0);
    }

    // ngt.hpp:222-238
    double one_minus_P(id_t64 x) const {
        std::map<id_t64, double>::const_iterator it = out[x].find(x);
        double Pxx = (it == out[x].end()) ? 0.0 : it->second;
        if (Pxx < 0.99) return 1.0 - Pxx;
        double s = 0.0;
        for (std::map<id_t64, double>::const_iterator e = out[x].begin(); e != out[x].end(); ++e)
            if (e->first != x) s += e->second;
        return s;
    }

    // ngt.hpp:295-333 (+246-255, 260-290); graph.hpp:245-277 for the deletion
    void remove_node(id_t64 x) {
        const double taux = tau[x];
        const double om = one_minus_P(x);
        // snapshot of row x without the self loop
        std::vector<std::pair<id_t64, double> > row;
        row.reserve(out[x].size());
        for (std::map<id_t64, double>::const_iterator e = out[x].begin(); e != out[x].end(); ++e)
            if (e->first != x) row.push_back(*e);
        std::vector<id_t64> preds(in[x].begin(), in[x].end());
        for (size_t i = 0; i < preds.size(); ++i) {
            const id_t64 u = preds[i];
            std::map<id_t64, double> &ou = out[u];
            const double Pux = ou[x];
            tau[u] += Pux * taux / om;
            for (size_t j = 0; j < row.size(); ++j) {
                const id_t64 v = row[j].first;
                std::map<id_t64, double>::iterator f = ou.find(v);
                if (f == ou.end()) {
                    f = ou.insert(std::make_pair(v, 0.0)).first;
                    if (v != u) in[v].insert(u);
                }
                f->second += Pux * row[j].second / om;   // one divide per pair, as ngt.hpp:281
            }
            ou.erase(x);
        }
        for (size_t j = 0; j < row.size(); ++j) in[row[j].first].erase(x);
        out[x].clear();
        in[x].clear();
        alive[x] = 0;
    }

    // ngt.hpp:338-346 with the ordering of ngt.hpp:192-200: always take the
    // node of smallest in+out degree (ties: smallest id).
    void remove_all(std::vector<id_t64> nodes) {
        std::set<std::pair<size_t, id_t64> > heap;
        std::vector<size_t> key(out.size(), 0);
        std::vector<char> listed(out.size(), 0);
        for (size_t i = 0; i < nodes.size(); ++i) {
            key[nodes[i]] = degree(nodes[i]);
            listed[nodes[i]] = 1;
            heap.insert(std::make_pair(key[nodes[i]], nodes[i]));
        }
        while (!heap.empty()) {
            const id_t64 x = heap.begin()->second;
            heap.erase(heap.begin());
            listed[x] = 0;
            std::set<id_t64> touched(in[x].begin(), in[x].end());
            for (std::map<id_t64, double>::const_iterator e = out[x].begin(); e != out[x].end(); ++e)
                if (e->first != x) touched.insert(e->first);
            remove_node(x);
            for (std::set<id_t64>::const_iterator t = touched.begin(); t != touched.end(); ++t) {
                if (!listed[*t]) continue;
                const size_t d = degree(*t);
                if (d != key[*t]) {
                    heap.erase(std::make_pair(key[*t], *t));
                    key[*t] = d;
                    heap.insert(std::make_pair(d, *t));
                }
            }
        }
    }
};

struct Oracle {
    Net g;
    size_t n;
    std::vector<id_t64> A, B;
    std::vector<char> inA, inB;
    std::vector<double> initial_tau;
    std::vector<double> final_omPxx, final_tau, committor, weight;
    bool have_weights;
    bool phase_one_done;
    std::string err;

    Oracle(size_t n_, size_t nnz, const id_t64 *src, const id_t64 *dst, const double *k,
           size_t nA, const id_t64 *A_, size_t nB, const id_t64 *B_)
        : g(n_), n(n_), A(A_, A_ + nA), B(B_, B_ + nB), inA(n_, 0), inB(n_, 0),
          initial_tau(n_, 0.0), final_omPxx(n_, NAN), final_tau(n_, NAN), committor(n_, NAN),
          weight(n_, 1.0), have_weights(false), phase_one_done(false) {
        // ngt.hpp:112-126: sum of outgoing rate constants (a given self rate counts)
        std::vector<double> sumk(n, 0.0);
        std::vector<char> seen(n, 0);
        for (size_t e = 0; e < nnz; ++e) {
            if (src[e] < 0 || dst[e] < 0 || (size_t)src[e] >= n || (size_t)dst[e] >= n)
                throw std::out_of_range("edge endpoint out of range");
            sumk[src[e]] += k[e];
            seen[src[e]] = seen[dst[e]] = 1;
        }
        // ngt.hpp:130-136
        for (size_t x = 0; x < n; ++x) {
            if (!seen[x]) continue;
            g.alive[x] = 1;
            g.tau[x] = 1.0 / sumk[x];
            initial_tau[x] = g.tau[x];
            g.out[x][x] = 0.0;
        }
        // ngt.hpp:139-147 (a later duplicate of the same (u,v) overwrites, like the map would)
        for (size_t e = 0; e < nnz; ++e) {
            const id_t64 u = src[e], v = dst[e];
            g.out[u][v] = k[e] * g.tau[u];
            if (u != v) g.in[v].insert(u);
        }
        for (size_t i = 0; i < nA; ++i) {
            if (A[i] < 0 || (size_t)A[i] >= n || !g.alive[A[i]]) throw std::invalid_argument("A node not in graph");
            inA[A[i]] = 1;
        }
        for (size_t i = 0; i < nB; ++i) {
            if (B[i] < 0 || (size_t)B[i] >= n || !g.alive[B[i]]) throw std::invalid_argument("B node not in graph");
            if (inA[B[i]]) throw std::invalid_argument("A and B overlap");
            inB[B[i]] = 1;
        }
    }

    std::vector<id_t64> intermediates() const {
        std::vector<id_t64> r;
        for (size_t x = 0; x < n; ++x)
            if (g.alive[x] && !inA[x] && !inB[x]) r.push_back((id_t64)x);
        return r;
    }

    void phase_one() {
        g.remove_all(intermediates());
        phase_one_done = true;
    }

    // ngt.hpp:362-423
    void reduce_group(const std::vector<id_t64> &grp) {
        if (grp.empty()) return;
        if (grp.size() == 1) {
            final_omPxx[grp[0]] = g.one_minus_P(grp[0]);
            final_tau[grp[0]] = g.tau[grp[0]];
            return;
        }
        Net work = g;
        std::vector<id_t64> rest(grp);
        while (rest.size() > 1) {
            const id_t64 x = rest.back();
            rest.pop_back();
            Net copy = work;
            copy.remove_all(rest);
            final_omPxx[x] = copy.one_minus_P(x);
            final_tau[x] = copy.tau[x];
            work.remove_node(x);
        }
        final_omPxx[rest[0]] = work.one_minus_P(rest[0]);
        final_tau[rest[0]] = work.tau[rest[0]];
    }

    void phase_two() {
        reduce_group(A);
        reduce_group(B);
    }

    void compute_rates() {
        phase_one();
        phase_two();
    }

    // ngt.hpp:444-458
    double rate_final(const std::vector<id_t64> &grp) const {
        double s = 0.0, norm = 0.0;
        for (size_t i = 0; i < grp.size(); ++i) {
            const double w = have_weights ? weight[grp[i]] : 1.0;
            s += w * final_omPxx[grp[i]] / final_tau[grp[i]];
            norm += w;
        }
        return s / norm;
    }

    // ngt.hpp:474-495
    double rate_ss(const std::vector<id_t64> &from, const std::vector<char> &into) const {
        double s = 0.0, norm = 0.0;
        for (size_t i = 0; i < from.size(); ++i) {
            const id_t64 a = from[i];
            double PaB = 0.0;
            for (std::map<id_t64, double>::const_iterator e = g.out[a].begin(); e != g.out[a].end(); ++e)
                if (into[e->first]) PaB += e->second;
            const double w = have_weights ? weight[a] : 1.0;
            s += w * PaB / initial_tau[a];
            norm += w;
        }
        return s / norm;
    }

    // ngt.hpp:518-539
    static double PxB(const Net &net, id_t64 x, const std::vector<char> &inB) {
        double pxb = 0.0, pxx = 0.0, om = 0.0;
        for (std::map<id_t64, double>::const_iterator e = net.out[x].begin(); e != net.out[x].end(); ++e) {
            if (e->first == x) pxx = e->second; else om += e->second;
            if (inB[e->first]) pxb += e->second;
        }
        if (pxx < 0.9) om = 1.0 - pxx;
        return pxb / om;
    }

    // ngt.hpp:553-630
    void compute_rates_and_committors() {
        std::vector<id_t64> todo = intermediates();
        while (!todo.empty()) {
            const id_t64 x = todo.back();
            todo.pop_back();
            Net copy = g;
            copy.remove_all(todo);
            final_omPxx[x] = copy.one_minus_P(x);
            final_tau[x] = copy.tau[x];
            committor[x] = PxB(copy, x, inB);
            g.remove_node(x);
        }
        phase_one_done = true;
        phase_two();
        for (size_t i = 0; i < A.size(); ++i) committor[A[i]] = 0.0;
        for (size_t i = 0; i < B.size(); ++i) committor[B[i]] = 1.0;
    }
};

}  // namespace

extern "C" {

void *ngto_create(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, const double *k,
                  int64_t nA, const int64_t *A, int64_t nB, const int64_t *B) {
    try {
        return new Oracle((size_t)n, (size_t)nnz, src, dst, k, (size_t)nA, A, (size_t)nB, B);
    } catch (...) {
        return 0;
    }
}

void ngto_destroy(void *h) { delete static_cast<Oracle *>(h); }

void ngto_set_weights(void *h, int64_t nw, const int64_t *ids, const double *w) {
    Oracle *o = static_cast<Oracle *>(h);
    o->have_weights = true;
    for (int64_t i = 0; i < nw; ++i)
        if (ids[i] >= 0 && (size_t)ids[i] < o->n) o->weight[ids[i]] = w[i];
}

int ngto_compute_rates(void *h) {
    try { static_cast<Oracle *>(h)->compute_rates(); } catch (...) { return 1; }
    return 0;
}

int ngto_phase_one(void *h) {
    try { static_cast<Oracle *>(h)->phase_one(); } catch (...) { return 1; }
    return 0;
}

int ngto_compute_rates_and_committors(void *h) {
    try { static_cast<Oracle *>(h)->compute_rates_and_committors(); } catch (...) { return 1; }
    return 0;
}

// out = {k_AB, k_BA, k_AB^SS, k_BA^SS}
void ngto_get_rates(void *h, double *out4) {
    Oracle *o = static_cast<Oracle *>(h);
    out4[0] = o->rate_final(o->A);
    out4[1] = o->rate_final(o->B);
    out4[2] = o->rate_ss(o->A, o->inB);
    out4[3] = o->rate_ss(o->B, o->inA);
}

// final 1-P'_xx and tau'_x for the nodes of A then B, in the order they were passed
void ngto_get_final(void *h, double *omPxx, double *tau) {
    Oracle *o = static_cast<Oracle *>(h);
    size_t j = 0;
    for (size_t i = 0; i < o->A.size(); ++i, ++j) { omPxx[j] = o->final_omPxx[o->A[i]]; tau[j] = o->final_tau[o->A[i]]; }
    for (size_t i = 0; i < o->B.size(); ++i, ++j) { omPxx[j] = o->final_omPxx[o->B[i]]; tau[j] = o->final_tau[o->B[i]]; }
}

void ngto_get_committors(void *h, double *q) {
    Oracle *o = static_cast<Oracle *>(h);
    for (size_t i = 0; i < o->n; ++i) q[i] = o->committor[i];
}

// dense row-major P over (A then B) and tau of the phase-1-reduced graph
void ngto_get_reduced(void *h, double *P, double *tau) {
    Oracle *o = static_cast<Oracle *>(h);
    std::vector<id_t64> keep(o->A);
    keep.insert(keep.end(), o->B.begin(), o->B.end());
    const size_t m = keep.size();
    for (size_t i = 0; i < m; ++i) {
        tau[i] = o->g.tau[keep[i]];
        for (size_t j = 0; j < m; ++j) {
            std::map<id_t64, double>::const_iterator f = o->g.out[keep[i]].find(keep[j]);
            P[i * m + j] = (f == o->g.out[keep[i]].end()) ? 0.0 : f->second;
        }
    }
}

}  // extern "C"

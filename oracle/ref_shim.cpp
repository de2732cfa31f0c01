// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// C-ABI shim around the UNMODIFIED reference engine.  It is compiled by
// oracle/Makefile against the headers where they lie
// (-I/root/reference/source -> kmc_rates/ngt.hpp, kmc_rates/graph.hpp) into
// oracle/_ref/libngt_ref.so.  No reference source is copied into this repo;
// this file only calls the reference's public interface:
//   graph_ns::NGT::NGT(rate_map_t&, A, B)        ngt.hpp:105
//   set_node_occupation_probabilities             ngt.hpp:183
//   compute_rates / compute_rates_and_committors  ngt.hpp:436 / 616
//   phase_one                                     ngt.hpp:352
//   get_rate_AB/BA, get_rate_AB_SS/BA_SS          ngt.hpp:463-513
//   get_committors                                ngt.hpp:98
//   public members final_omPxx, final_tau, _graph ngt.hpp:37-57
// The entry points mirror oracle/ngt_oracle.cpp (prefix ngtr_ instead of ngto_)
// so the Python loader can drive either with the same code.
#include <cstdint>
#include <cmath>
#include <list>
#include <map>
#include <vector>

#include "kmc_rates/graph.hpp"
#include "kmc_rates/ngt.hpp"

namespace {
struct Ref {
    graph_ns::NGT *ngt;
    std::vector<graph_ns::node_id> A, B;
    size_t n;
};
}  // namespace

extern "C" {

void *ngtr_create(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, const double *k,
                  int64_t nA, const int64_t *A, int64_t nB, const int64_t *B) {
    graph_ns::NGT::rate_map_t rates;
    for (int64_t e = 0; e < nnz; ++e)
        rates[std::pair<graph_ns::node_id, graph_ns::node_id>(src[e], dst[e])] = k[e];
    std::list<graph_ns::node_id> la, lb;
    Ref *r = new Ref;
    r->n = (size_t)n;
    for (int64_t i = 0; i < nA; ++i) { la.push_back(A[i]); r->A.push_back(A[i]); }
    for (int64_t i = 0; i < nB; ++i) { lb.push_back(B[i]); r->B.push_back(B[i]); }
    try {
        r->ngt = new graph_ns::NGT(rates, la, lb);
    } catch (...) {
        delete r;
        return 0;
    }
    return r;
}

void ngtr_destroy(void *h) {
    Ref *r = static_cast<Ref *>(h);
    if (!r) return;
    delete r->ngt;
    delete r;
}

void ngtr_set_weights(void *h, int64_t nw, const int64_t *ids, const double *w) {
    std::map<graph_ns::node_id, double> peq;
    for (int64_t i = 0; i < nw; ++i) peq[ids[i]] = w[i];
    static_cast<Ref *>(h)->ngt->set_node_occupation_probabilities(peq);
}

int ngtr_compute_rates(void *h) {
    try { static_cast<Ref *>(h)->ngt->compute_rates(); } catch (...) { return 1; }
    return 0;
}

int ngtr_phase_one(void *h) {
    try { static_cast<Ref *>(h)->ngt->phase_one(); } catch (...) { return 1; }
    return 0;
}

int ngtr_compute_rates_and_committors(void *h) {
    try { static_cast<Ref *>(h)->ngt->compute_rates_and_committors(); } catch (...) { return 1; }
    return 0;
}

void ngtr_get_rates(void *h, double *out4) {
    graph_ns::NGT *g = static_cast<Ref *>(h)->ngt;
    out4[0] = g->get_rate_AB();
    out4[1] = g->get_rate_BA();
    out4[2] = g->get_rate_AB_SS();
    out4[3] = g->get_rate_BA_SS();
}

void ngtr_get_final(void *h, double *omPxx, double *tau) {
    Ref *r = static_cast<Ref *>(h);
    size_t j = 0;
    for (size_t i = 0; i < r->A.size(); ++i, ++j) { omPxx[j] = r->ngt->final_omPxx.at(r->A[i]); tau[j] = r->ngt->final_tau.at(r->A[i]); }
    for (size_t i = 0; i < r->B.size(); ++i, ++j) { omPxx[j] = r->ngt->final_omPxx.at(r->B[i]); tau[j] = r->ngt->final_tau.at(r->B[i]); }
}

void ngtr_get_committors(void *h, double *q) {
    Ref *r = static_cast<Ref *>(h);
    std::map<graph_ns::node_id, double> const &m = r->ngt->get_committors();
    for (size_t i = 0; i < r->n; ++i) q[i] = NAN;
    for (std::map<graph_ns::node_id, double>::const_iterator it = m.begin(); it != m.end(); ++it)
        if (it->first < r->n) q[it->first] = it->second;
}

void ngtr_get_reduced(void *h, double *P, double *tau) {
    Ref *r = static_cast<Ref *>(h);
    std::vector<graph_ns::node_id> keep(r->A);
    keep.insert(keep.end(), r->B.begin(), r->B.end());
    const size_t m = keep.size();
    for (size_t i = 0; i < m; ++i) {
        graph_ns::node_ptr u = r->ngt->_graph->get_node(keep[i]);
        tau[i] = u->tau;
        for (size_t j = 0; j < m; ++j) {
            graph_ns::edge_ptr e = u->get_successor_edge(r->ngt->_graph->get_node(keep[j]));
            P[i * m + j] = e ? e->P : 0.0;
        }
    }
}

}  // extern "C"

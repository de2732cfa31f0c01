// Host-side symbolic analysis: ordering, supernodes, fronts, assembly tree.  See symbolic.h.
#include "symbolic.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace ngt {

void build_adjacency(int64_t n, int64_t nnz, const int64_t *src, const int64_t *dst, Adjacency &g) {
    g.n = n;
    g.ptr.assign(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) {
        if (src[e] == dst[e]) continue;
        g.ptr[src[e] + 1]++;
        g.ptr[dst[e] + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
    std::vector<int32_t> tmp(g.ptr[n]);
    std::vector<int64_t> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (int64_t e = 0; e < nnz; ++e) {
        if (src[e] == dst[e]) continue;
        tmp[fill[src[e]]++] = (int32_t)dst[e];
        tmp[fill[dst[e]]++] = (int32_t)src[e];
    }
    // sort + dedupe each row, compact
    std::vector<int64_t> nptr(n + 1, 0);
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t *b = tmp.data() + g.ptr[i], *e = tmp.data() + g.ptr[i + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        nptr[i] = w;
        for (int32_t *q = b; q != e; ++q) tmp[w++] = *q;   // w <= read cursor, safe in place
    }
    nptr[n] = w;
    tmp.resize(w);
    g.ptr.swap(nptr);
    g.adj.swap(tmp);
}

namespace {

// ---------------------------------------------------------------------------------------------
// Nested dissection by BFS level structures (George's automatic nested dissection) on the graph
// induced on the intermediate nodes.  Emits nodes in elimination order and the supernode boundaries.
// ---------------------------------------------------------------------------------------------
struct Dissector {
    const Adjacency &g;
    int leaf;
    std::vector<int32_t> tag;      // tag[v] == cur  <=> v belongs to the subset being processed
    std::vector<int32_t> lvl;      // BFS level scratch
    std::vector<int32_t> comp;     // component label scratch
    int32_t next_tag = 1;
    std::vector<int32_t> order;    // output: node ids in elimination order
    std::vector<int32_t> sn_start; // output: start position of each supernode

    Dissector(const Adjacency &g_, int leaf_) : g(g_), leaf(leaf_), tag(g_.n, 0), lvl(g_.n, -1), comp(g_.n, -1) {}

    void emit(const std::vector<int32_t> &nodes) {
        if (nodes.empty()) return;
        sn_start.push_back((int32_t)order.size());
        order.insert(order.end(), nodes.begin(), nodes.end());
    }

    // BFS inside the tagged subset from `start`; fills lvl[] for reached nodes, returns them in visit order.
    void bfs(int32_t start, int32_t t, std::vector<int32_t> &visit, int32_t &nlev) {
        visit.clear();
        visit.push_back(start);
        lvl[start] = 0;
        size_t head = 0;
        nlev = 1;
        while (head < visit.size()) {
            int32_t u = visit[head++];
            for (int64_t e = g.ptr[u]; e < g.ptr[u + 1]; ++e) {
                int32_t v = g.adj[e];
                if (tag[v] != t || lvl[v] >= 0) continue;
                lvl[v] = lvl[u] + 1;
                nlev = lvl[v] + 1;
                visit.push_back(v);
            }
        }
    }

    void dissect(std::vector<int32_t> &S) {
        if (S.empty()) return;
        if ((int)S.size() <= leaf) { emit(S); return; }
        const int32_t t = next_tag++;
        for (int32_t v : S) { tag[v] = t; lvl[v] = -1; }

        // ---- connected components of the subset
        std::vector<std::vector<int32_t> > comps;
        {
            std::vector<int32_t> visit;
            int32_t nlev;
            for (int32_t v : S) {
                if (lvl[v] >= 0) continue;
                bfs(v, t, visit, nlev);
                comps.push_back(visit);
            }
        }
        if (comps.size() > 1) {
            // large components are dissected on their own; small ones are packed into leaf-sized bins
            std::vector<int32_t> bin;
            for (auto &c : comps) {
                if ((int)c.size() > leaf) continue;
                if ((int)(bin.size() + c.size()) > leaf) { emit(bin); bin.clear(); }
                bin.insert(bin.end(), c.begin(), c.end());
            }
            emit(bin);
            for (auto &c : comps)
                if ((int)c.size() > leaf) dissect(c);
            return;
        }
        S.swap(comps[0]);   // same set, BFS order
        comps.clear();

        // ---- pseudo-peripheral start node: repeat BFS from a min-degree node of the last level
        std::vector<int32_t> visit;
        int32_t nlev = 0, start = S[0];
        for (int it = 0; it < 3; ++it) {
            for (int32_t v : S) lvl[v] = -1;
            bfs(start, t, visit, nlev);
            int32_t best = visit.back();
            int64_t bestdeg = g.ptr[best + 1] - g.ptr[best];
            for (size_t i = visit.size(); i-- > 0;) {
                int32_t v = visit[i];
                if (lvl[v] != nlev - 1) break;
                int64_t d = g.ptr[v + 1] - g.ptr[v];
                if (d < bestdeg) { bestdeg = d; best = v; }
            }
            if (it == 2 || best == start) break;
            start = best;
        }
        if (nlev < 3) { emit(S); return; }   // no separator exists (e.g. a clique): one dense supernode

        // ---- choose the separator level: smallest level with both sides >= 25 % if possible
        std::vector<int64_t> cnt(nlev, 0);
        for (int32_t v : S) cnt[lvl[v]]++;
        std::vector<int64_t> before(nlev + 1, 0);
        for (int i = 0; i < nlev; ++i) before[i + 1] = before[i] + cnt[i];
        const int64_t tot = (int64_t)S.size();
        int bestl = -1;
        double bestcost = 1e300;
        for (int l = 1; l <= nlev - 2; ++l) {
            int64_t a = before[l], b = tot - before[l + 1];
            int64_t mn = std::min(a, b);
            if (mn <= 0) continue;
            // separator size, penalised by imbalance
            double bal = (double)mn / (double)tot;
            double cost = (double)cnt[l] * (bal >= 0.25 ? 1.0 : (0.25 / bal) * (0.25 / bal));
            if (cost < bestcost) { bestcost = cost; bestl = l; }
        }
        if (bestl < 0 || cnt[bestl] * 10 > tot * 6) { emit(S); return; }

        // ---- split; trim separator nodes that do not touch the far side
        std::vector<int32_t> P1, P2, sep;
        for (int32_t v : S) {
            if (lvl[v] < bestl) P1.push_back(v);
            else if (lvl[v] > bestl) P2.push_back(v);
            else {
                bool far = false;
                for (int64_t e = g.ptr[v]; e < g.ptr[v + 1] && !far; ++e) {
                    int32_t w = g.adj[e];
                    if (tag[w] == t && lvl[w] == bestl + 1) far = true;
                }
                if (far) sep.push_back(v); else P1.push_back(v);
            }
        }
        S.clear();
        S.shrink_to_fit();
        dissect(P1);
        dissect(P2);
        emit(sep);
    }
};

void finish_layout(Symbolic &S) {
    // pool layout, levels, batches, statistics
    const int32_t nf = (int32_t)S.fronts.size();
    int64_t off = 0;
    S.max_front = 0;
    for (auto &F : S.fronts) {
        F.f = F.k + F.m;
        F.ld = ((F.f + 1 + 7) / 8) * 8;
        F.off = off;
        off += (int64_t)F.f * F.ld;
        off = ((off + 15) / 16) * 16;
        S.max_front = std::max(S.max_front, F.f);
    }
    S.pool_doubles = off;

    // children CSR
    S.child_ptr.assign(nf + 1, 0);
    for (int32_t i = 0; i < nf; ++i)
        if (S.fronts[i].parent >= 0) S.child_ptr[S.fronts[i].parent + 1]++;
    for (int32_t i = 0; i < nf; ++i) S.child_ptr[i + 1] += S.child_ptr[i];
    S.child_ids.assign(S.child_ptr[nf], 0);
    {
        std::vector<int64_t> fill(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int32_t i = 0; i < nf; ++i)
            if (S.fronts[i].parent >= 0) S.child_ids[fill[S.fronts[i].parent]++] = i;
    }
    // levels (fronts are in elimination order, children before parents; root last)
    for (auto &F : S.fronts) F.level = 0;
    for (int32_t i = 0; i < nf; ++i) {
        int32_t p = S.fronts[i].parent;
        if (p >= 0) S.fronts[p].level = std::max(S.fronts[p].level, S.fronts[i].level + 1);
    }
    S.n_levels = 0;
    for (auto &F : S.fronts) S.n_levels = std::max(S.n_levels, F.level + 1);

    // relative index maps child -> parent
    S.rel_off.assign(nf + 1, 0);
    for (int32_t i = 0; i < nf; ++i) S.rel_off[i + 1] = S.rel_off[i] + (S.fronts[i].parent >= 0 ? S.fronts[i].m : 0);
    S.rel.assign(S.rel_off[nf], -1);
    for (int32_t i = 0; i < nf; ++i) {
        const Front &F = S.fronts[i];
        if (F.parent < 0) continue;
        for (int32_t j = 0; j < F.m; ++j) {
            int32_t li = local_index(S, F.parent, S.fidx[F.idx_off + F.k + j]);
            if (li < 0) throw std::logic_error("symbolic: child update index missing from parent front");
            S.rel[S.rel_off[i] + j] = li;
        }
    }

    // batches: per level, fronts with pivots, grouped by size class (within 2x) and capped for gridDim.y
    S.batches.clear();
    std::vector<std::vector<int32_t> > by_level(S.n_levels);
    for (int32_t i = 0; i < nf; ++i)
        if (S.fronts[i].k > 0) by_level[S.fronts[i].level].push_back(i);
    for (int32_t l = 0; l < S.n_levels; ++l) {
        auto &v = by_level[l];
        std::sort(v.begin(), v.end(), [&](int32_t a, int32_t b) {
            if (S.fronts[a].f != S.fronts[b].f) return S.fronts[a].f > S.fronts[b].f;
            return a < b;
        });
        size_t i = 0;
        while (i < v.size()) {
            Batch b;
            b.level = l;
            b.max_f = S.fronts[v[i]].f;
            b.max_k = 0;
            size_t j = i;
            while (j < v.size() && j - i < 32768 && S.fronts[v[j]].f * 2 >= b.max_f) {
                b.max_k = std::max(b.max_k, S.fronts[v[j]].k);
                b.fronts.push_back(v[j]);
                ++j;
            }
            S.batches.push_back(b);
            i = j;
        }
    }

    // algorithmic work: degrees at elimination time under this ordering
    S.alg_flops = S.alg_bytes = 0;
    for (auto &F : S.fronts) {
        for (int32_t i = 0; i < F.k; ++i) {
            double d = (double)(F.k - 1 - i) + F.m;
            S.alg_flops += 2.0 * d * d + 4.0 * d;
            S.alg_bytes += 16.0 * d * d + 24.0 * d + 16.0 * d + 8.0;
        }
    }
}

}  // namespace

int32_t local_index(const Symbolic &S, int32_t fr, int32_t p) {
    const Front &F = S.fronts[fr];
    if (F.k > 0 && p >= F.first && p < F.first + F.k) return p - F.first;
    const int32_t *b = S.fidx.data() + F.idx_off + F.k, *e = b + F.m;
    const int32_t *it = std::lower_bound(b, e, p);
    if (it == e || *it != p) return -1;
    return F.k + (int32_t)(it - b);
}

static void assign_kept_positions(const std::vector<int64_t> &A, const std::vector<int64_t> &B, Symbolic &S) {
    int32_t p = (int32_t)S.nI;
    for (int64_t a : A) { S.pos[a] = p; S.node_at[p] = (int32_t)a; ++p; }
    for (int64_t b : B) { S.pos[b] = p; S.node_at[p] = (int32_t)b; ++p; }
}

void analyse_dense(int64_t n, const std::vector<int64_t> &A, const std::vector<int64_t> &B, Symbolic &S) {
    S = Symbolic();
    S.n = n;
    S.nA = (int64_t)A.size();
    S.nB = (int64_t)B.size();
    S.nI = n - S.nA - S.nB;
    S.pos.assign(n, -1);
    S.node_at.assign(n, -1);
    std::vector<char> kept(n, 0);
    for (int64_t a : A) kept[a] = 1;
    for (int64_t b : B) kept[b] = 1;
    int32_t p = 0;
    for (int64_t u = 0; u < n; ++u)
        if (!kept[u]) { S.pos[u] = p; S.node_at[p] = (int32_t)u; ++p; }
    assign_kept_positions(A, B, S);
    const int32_t m = (int32_t)(S.nA + S.nB);
    S.pos_front.assign(n, 0);
    if (S.nI > 0) {
        Front F{};
        F.k = (int32_t)S.nI; F.m = m; F.first = 0; F.parent = 1; F.idx_off = 0;
        S.fronts.push_back(F);
        for (int32_t i = 0; i < n; ++i) S.fidx.push_back(i);
    }
    Front R{};
    R.k = 0; R.m = m; R.first = (int32_t)S.nI; R.parent = -1; R.idx_off = (int64_t)S.fidx.size();
    S.fronts.push_back(R);
    for (int32_t i = 0; i < m; ++i) S.fidx.push_back((int32_t)S.nI + i);
    S.root_id = (int32_t)S.fronts.size() - 1;
    for (int64_t q = 0; q < n; ++q) S.pos_front[q] = q < S.nI ? 0 : S.root_id;
    finish_layout(S);
}

void analyse_sparse(const Adjacency &g, const std::vector<char> &present,
                    const std::vector<int64_t> &A, const std::vector<int64_t> &B,
                    int leaf_size, Symbolic &S) {
    S = Symbolic();
    const int64_t n = g.n;
    S.n = n;
    S.nA = (int64_t)A.size();
    S.nB = (int64_t)B.size();
    std::vector<char> kept(n, 0);
    for (int64_t a : A) kept[a] = 1;
    for (int64_t b : B) kept[b] = 1;

    // ---- ordering of the intermediates
    std::vector<int32_t> inter;
    for (int64_t u = 0; u < n; ++u)
        if (present[u] && !kept[u]) inter.push_back((int32_t)u);
    S.nI = (int64_t)inter.size();
    Dissector D(g, leaf_size > 0 ? leaf_size : 64);
    // the dissector must not walk through A ∪ B or absent nodes: they never carry the current tag
    D.order.reserve(inter.size());
    D.dissect(inter);
    if ((int64_t)D.order.size() != S.nI) throw std::logic_error("symbolic: ordering lost nodes");

    const int64_t npos = S.nI + S.nA + S.nB;
    S.pos.assign(n, -1);
    S.node_at.assign(npos, -1);
    for (int64_t p = 0; p < S.nI; ++p) { S.pos[D.order[p]] = (int32_t)p; S.node_at[p] = D.order[p]; }
    assign_kept_positions(A, B, S);

    // ---- supernodal symbolic factorisation
    const int32_t ns = (int32_t)D.sn_start.size();
    D.sn_start.push_back((int32_t)S.nI);
    S.pos_front.assign(npos, -1);
    S.fronts.resize(ns + 1);
    std::vector<int32_t> mark(npos, -1);
    std::vector<std::vector<int32_t> > kids(ns + 1);
    std::vector<std::vector<int32_t> > upd(ns);   // update sets (freed into fidx at the end)
    for (int32_t s = 0; s < ns; ++s)
        for (int32_t p = D.sn_start[s]; p < D.sn_start[s + 1]; ++p) S.pos_front[p] = s;
    for (int32_t s = 0; s < ns; ++s) {
        const int32_t first = D.sn_start[s], last = D.sn_start[s + 1];
        std::vector<int32_t> &U = upd[s];
        for (int32_t p = first; p < last; ++p) {
            const int32_t u = S.node_at[p];
            for (int64_t e = g.ptr[u]; e < g.ptr[u + 1]; ++e) {
                const int32_t q = S.pos[g.adj[e]];
                if (q >= last && mark[q] != s) { mark[q] = s; U.push_back(q); }
            }
        }
        for (int32_t c : kids[s])
            for (int32_t q : upd[c])
                if (q >= last && mark[q] != s) { mark[q] = s; U.push_back(q); }
        std::sort(U.begin(), U.end());
        Front &F = S.fronts[s];
        F.k = last - first;
        F.m = (int32_t)U.size();
        F.first = first;
        if (U.empty()) F.parent = -1;
        else if (U[0] >= S.nI) F.parent = ns;
        else F.parent = S.pos_front[U[0]];
        if (F.parent >= 0) kids[F.parent].push_back(s);
    }
    // root
    {
        Front &R = S.fronts[ns];
        R.k = 0;
        R.m = (int32_t)(S.nA + S.nB);
        R.first = (int32_t)S.nI;
        R.parent = -1;
        for (int64_t p = S.nI; p < npos; ++p) S.pos_front[p] = ns;
    }
    S.root_id = ns;
    // index lists
    int64_t tot = 0;
    for (int32_t s = 0; s < ns; ++s) tot += S.fronts[s].k + S.fronts[s].m;
    tot += S.fronts[ns].m;
    S.fidx.reserve(tot);
    for (int32_t s = 0; s < ns; ++s) {
        Front &F = S.fronts[s];
        F.idx_off = (int64_t)S.fidx.size();
        for (int32_t p = F.first; p < F.first + F.k; ++p) S.fidx.push_back(p);
        S.fidx.insert(S.fidx.end(), upd[s].begin(), upd[s].end());
        std::vector<int32_t>().swap(upd[s]);
    }
    S.fronts[ns].idx_off = (int64_t)S.fidx.size();
    for (int64_t p = S.nI; p < npos; ++p) S.fidx.push_back((int32_t)p);
    finish_layout(S);
}

}  // namespace ngt

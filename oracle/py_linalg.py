"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

Python-3 restatement of the reference's linear-algebra cross-check
(reference kmc_rates/rates_linalg.py: reduce_rates :13-41, compute_sum_out_rates :43-59,
CommittorLinalg :148-207, MfptLinalgSparse :210-285, TwoStateRates :389-477).  It is the
INDEPENDENT oracle: it never eliminates nodes; it solves

    sum_v k_uv (t_v - t_u) = -1      (t = 0 on B)      mean first-passage times
    sum_v k_uv (q_v - q_u) =  0      (q = 0 on A, 1 on B)  committors

with scipy's sparse LU, exactly what the reference does (`spsolve`, :197, :279).  Used at the sizes
the elimination oracles cannot reach (10^5-node landscape, 16 384-node complete graph).

Third-party arithmetic: scipy.sparse.linalg.spsolve (SuperLU).  The reference does not pin a scipy
version (no requirements file); this container has scipy 1.18.1.  Pinned by the reference's
tests: 3-state k=1.0, k_SS=1.5 (reference kmc_rates/tests/test_linalg.py:12-21).
"""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg


class TwoStateRatesCOO(object):
    """TwoStateRates over dense ids and COO arrays (src, dst, k); A, B lists of ids."""

    def __init__(self, n, src, dst, k, A, B, weights=None):
        self.n = int(n)
        self.src = np.asarray(src, np.int64)
        self.dst = np.asarray(dst, np.int64)
        self.k = np.asarray(k, np.float64)
        self.A = [int(a) for a in A]
        self.B = [int(b) for b in B]
        self.weights = {a: 1.0 for a in self.A} if weights is None else weights
        # reference :13-41 reduce_rates: drop every rate whose tail is not connected to B
        import scipy.sparse.csgraph
        adj = scipy.sparse.coo_matrix((np.ones(self.src.size), (self.src, self.dst)), shape=(self.n, self.n))
        _nc, lab = scipy.sparse.csgraph.connected_components(adj, directed=False)
        good = lab == lab[self.B[0]]
        if not all(good[b] for b in self.B):
            raise Exception("the nodes in B are not all connected")
        if not all(good[a] for a in self.A):
            raise Exception("the A nodes are not all connected to the B nodes")
        keep = good[self.src]
        self.src, self.dst, self.k = self.src[keep], self.dst[keep], self.k[keep]
        # reference :43-59 sums each node's rates in sorted order; np.add.at on rate-sorted edges does the same
        order = np.argsort(self.k, kind="stable")
        self.sum_out = np.zeros(self.n)
        np.add.at(self.sum_out, self.src[order], self.k[order])
        self.mfpt = None
        self.q = None

    def _solve(self, absorbing, rhs_fn):
        free = np.ones(self.n, bool)
        free[list(absorbing)] = False
        free &= self.sum_out > 0
        idx = -np.ones(self.n, np.int64)
        idx[free] = np.arange(int(free.sum()))
        m = int(free.sum())
        inner = free[self.src] & free[self.dst] & (self.src != self.dst)
        M = scipy.sparse.coo_matrix((self.k[inner], (idx[self.src[inner]], idx[self.dst[inner]])), shape=(m, m)).tocsr()
        self_rate = np.zeros(self.n)
        loops = self.src == self.dst
        np.add.at(self_rate, self.src[loops], self.k[loops])
        M = M - scipy.sparse.diags((self.sum_out - self_rate)[free])
        rhs = rhs_fn(free, idx, m)
        if m == 1:
            x = np.array([rhs[0] / M[0, 0]])
        else:
            x = scipy.sparse.linalg.spsolve(M.tocsc(), rhs)
        out = np.full(self.n, np.nan)
        out[free] = x
        return out

    def compute_rates(self):
        """reference :246-285, :458-471: K_II t = -1 with B absorbing."""
        t = self._solve(self.B, lambda free, idx, m: -np.ones(m))
        t[self.B] = 0.0
        self.mfpt = t
        return t

    def compute_committors(self):
        """reference :161-207: K_II q = -k_IB with A, B absorbing."""
        inB = np.zeros(self.n, bool)
        inB[self.B] = True

        def rhs(free, idx, m):
            r = np.zeros(m)
            sel = free[self.src] & inB[self.dst]
            np.add.at(r, idx[self.src[sel]], -self.k[sel])
            return r

        q = self._solve(self.A + self.B, rhs)
        q[self.A] = 0.0
        q[self.B] = 1.0
        self.q = q
        return q

    def get_rate_AB(self):
        """reference :409-417"""
        num = sum(self.weights[a] / self.mfpt[a] for a in self.A)
        return num / sum(self.weights[a] for a in self.A)

    def get_rate_AB_SS(self):
        """reference :419-447: sum_a w_a sum_{v not in A} k_av q_v / sum w."""
        inA = np.zeros(self.n, bool)
        inA[self.A] = True
        sel = inA[self.src] & ~inA[self.dst]
        flux = np.zeros(self.n)
        np.add.at(flux, self.src[sel], self.k[sel] * self.q[self.dst[sel]])
        num = sum(self.weights[a] * flux[a] for a in self.A)
        return num / sum(self.weights[a] for a in self.A)

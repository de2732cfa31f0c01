// PATHSAMPLE ingest on the device: rate constants of a min.data / ts.data database.
//
// What the reference does on the host, one transition state at a time (reference kmc_rates/utils.py:29-69,73-95):
//   log k(u->v) = log(pg_u / (2 pi pg_ts)) + (fvib_u - fvib_ts) / 2 - (E_ts - E_u) / T        for both directions,
//   shifted by the largest log rate, transition states between the same ordered pair of minima summed in log space
//   (log_sum2, utils.py:6-11), self-connecting transition states dropped, then exponentiated.
// Here: one kernel evaluates the 2 nts directed log rates and their (u, v) keys, a stable sort groups equal pairs and a
// segmented log-sum-exp merges them (thrust / CUB primitives: sorting is plumbing, the arithmetic is the functor below).
#include "../../include/ngt_b200.h"

#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/extrema.h>
#include <thrust/reduce.h>
#include <thrust/sort.h>

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace ngt {

constexpr long long PS_DROPPED = 0x7fffffffffffffffLL;     // key of a self-connecting transition state (sorts last)

__global__ void ps_log_rates_kernel(long long nmin, long long nts, const double *mind, const double *tsd, const long long *m1,
                                    const long long *m2, double T, double *logk, long long *key, int *errflag) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nts) return;
    const long long u = m1[t], v = m2[t];
    if (u < 0 || u >= nmin || v < 0 || v >= nmin) {
        atomicOr(errflag, 1);
        logk[t] = logk[nts + t] = -INFINITY;
        key[t] = key[nts + t] = PS_DROPPED;
        return;
    }
    const double ets = tsd[3 * t], fts = tsd[3 * t + 1], pts = tsd[3 * t + 2];
    const double two_pi = 6.283185307179586476925286766559;
    logk[t] = log(mind[3 * u + 2] / (two_pi * pts)) + (mind[3 * u + 1] - fts) / 2. - (ets - mind[3 * u]) / T;
    logk[nts + t] = log(mind[3 * v + 2] / (two_pi * pts)) + (mind[3 * v + 1] - fts) / 2. - (ets - mind[3 * v]) / T;
    key[t] = u == v ? PS_DROPPED : u * nmin + v;
    key[nts + t] = u == v ? PS_DROPPED : v * nmin + u;
}

struct LogSum2 {       // log(exp(a) + exp(b)), the larger argument factored out (reference utils.py:6-11)
    __host__ __device__ double operator()(double a, double b) const {
        return a > b ? a + log(1.0 + exp(b - a)) : b + log(1.0 + exp(a - b));
    }
};

__global__ void ps_finish_kernel(long long n, long long nmin, const long long *key, const double *logk, double shift,
                                 long long *src, long long *dst, double *k) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    src[i] = key[i] / nmin;
    dst[i] = key[i] % nmin;
    k[i] = exp(logk[i] - shift);
}

}  // namespace ngt

namespace {
thread_local std::string g_ps_err;
template <class T>
struct Dev {
    T *p = nullptr;
    explicit Dev(size_t n) { if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) { p = nullptr; throw std::runtime_error("cudaMalloc failed"); } }
    ~Dev() { if (p) cudaFree(p); }
    Dev(const Dev &) = delete;
    Dev &operator=(const Dev &) = delete;
};
void ck(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace

extern "C" const char *ngt_pathsample_last_error(void) { return g_ps_err.c_str(); }

extern "C" int ngt_pathsample_rates(int64_t nmin, int64_t nts, const double *mindata, const double *tsdata, const int64_t *m1,
                                    const int64_t *m2, double T, int32_t device, int64_t *n_out, int64_t *src, int64_t *dst,
                                    double *k, double *log_rate_subtracted) {
    try {
        if (nmin <= 0 || nts <= 0 || !mindata || !tsdata || !m1 || !m2 || !n_out || !src || !dst || !k || !(T > 0))
            { g_ps_err = "bad argument"; return NGT_E_INVALID; }
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_ps_err = "no CUDA device — kmc_rates_b200 has no CPU fallback"; return NGT_E_CUDA; }
        if (device < 0 || device >= ndev) { g_ps_err = "device ordinal out of range"; return NGT_E_INVALID; }
        ck(cudaSetDevice(device), "cudaSetDevice");
        static_assert(sizeof(long long) == sizeof(int64_t), "ids are uploaded as they are");
        const long long n2 = 2 * nts;
        Dev<double> dmin(3 * nmin), dts(3 * nts), dlog(n2), dmerged(n2), dk(n2);
        Dev<long long> dm1(nts), dm2(nts), dkey(n2), dukey(n2), dsrc(n2), ddst(n2);
        Dev<int> derr(1);
        ck(cudaMemset(derr.p, 0, sizeof(int)), "memset");
        ck(cudaMemcpy(dmin.p, mindata, 3 * nmin * sizeof(double), cudaMemcpyHostToDevice), "H2D min.data");
        ck(cudaMemcpy(dts.p, tsdata, 3 * nts * sizeof(double), cudaMemcpyHostToDevice), "H2D ts.data");
        ck(cudaMemcpy(dm1.p, m1, nts * sizeof(int64_t), cudaMemcpyHostToDevice), "H2D min1");
        ck(cudaMemcpy(dm2.p, m2, nts * sizeof(int64_t), cudaMemcpyHostToDevice), "H2D min2");
        ngt::ps_log_rates_kernel<<<(unsigned)((nts + 255) / 256), 256>>>(nmin, nts, dmin.p, dts.p, dm1.p, dm2.p, T, dlog.p, dkey.p, derr.p);
        ck(cudaGetLastError(), "ps_log_rates_kernel");
        int flag = 0;
        ck(cudaMemcpy(&flag, derr.p, sizeof(int), cudaMemcpyDeviceToHost), "D2H flag");
        if (flag) { g_ps_err = "ts.data names a minimum that is not in min.data"; return NGT_E_INVALID; }
        thrust::device_ptr<double> lg(dlog.p), mg(dmerged.p);
        thrust::device_ptr<long long> ky(dkey.p), uk(dukey.p);
        // the shift is the largest log rate over ALL transition states, self-connecting ones included (utils.py:84)
        const double shift = *thrust::max_element(thrust::device, lg, lg + n2);
        thrust::stable_sort_by_key(thrust::device, ky, ky + n2, lg);
        auto ends = thrust::reduce_by_key(thrust::device, ky, ky + n2, lg, uk, mg, thrust::equal_to<long long>(), ngt::LogSum2());
        long long nuniq = ends.first - uk;
        if (nuniq > 0) {
            long long last = 0;
            ck(cudaMemcpy(&last, dukey.p + nuniq - 1, sizeof(long long), cudaMemcpyDeviceToHost), "D2H key");
            if (last == ngt::PS_DROPPED) --nuniq;            // the group of self-connecting transition states
        }
        if (nuniq > 0) {
            ngt::ps_finish_kernel<<<(unsigned)((nuniq + 255) / 256), 256>>>(nuniq, nmin, dukey.p, dmerged.p, shift, dsrc.p, ddst.p, dk.p);
            ck(cudaGetLastError(), "ps_finish_kernel");
            ck(cudaMemcpy(src, dsrc.p, nuniq * sizeof(int64_t), cudaMemcpyDeviceToHost), "D2H src");
            ck(cudaMemcpy(dst, ddst.p, nuniq * sizeof(int64_t), cudaMemcpyDeviceToHost), "D2H dst");
            ck(cudaMemcpy(k, dk.p, nuniq * sizeof(double), cudaMemcpyDeviceToHost), "D2H k");
        }
        *n_out = nuniq;
        if (log_rate_subtracted) *log_rate_subtracted = shift;
        return NGT_OK;
    } catch (const std::exception &e) {
        g_ps_err = e.what();
        return NGT_E_CUDA;
    }
}

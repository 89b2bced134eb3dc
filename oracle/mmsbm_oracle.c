/* CPU ORACLE - test infrastructure only, NOT a product path.
 *
 * Plain-C restatement of the MMSBM EM step and log-likelihood of
 * AleixMT/TrigenicInteractionPredictor, src/TrigenicInteractionPredictor.py (TIP.py):
 *   oracle_em_step  <- Model.make_iteration      TIP.py:984-1043
 *   oracle_loglik   <- Model.compute_likelihood  TIP.py:952-974
 * Arithmetic is IEEE double in the reference's literal operation order (compile with
 * -ffp-contract=off so no multiply-add is fused); with the same inputs the results are
 * expected to be bit-identical to CPython's.  Parity status: pinned by
 * tests/test_oracle_golden.py against vectors produced by the reference itself.
 *
 * Layout: theta[P][K], pr[K][K][K][2] row-major, ids[L][3] (key slot order), cnt[L][2].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_EPS 1e-10
#define ORACLE_R 2

/* E-step over links [lo,hi): accumulates into ntheta / npr (TIP.py:987-1012). */
static void estep_range(int K, int64_t lo, int64_t hi, const int64_t *ids, const int64_t *cnt,
                        const double *theta, const double *pr, double *ntheta, double *npr)
{
    for (int64_t l = lo; l < hi; ++l) {
        const int64_t a = ids[3 * l], b = ids[3 * l + 1], c = ids[3 * l + 2];
        const double *ta = theta + a * K, *tb = theta + b * K, *tc = theta + c * K;
        double d[ORACLE_R] = {ORACLE_EPS, ORACLE_EPS};
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j)
                for (int k = 0; k < K; ++k)
                    for (int r = 0; r < ORACLE_R; ++r)
                        d[r] += ta[i] * tb[j] * tc[k] * pr[((i * K + j) * K + k) * ORACLE_R + r];
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j)
                for (int k = 0; k < K; ++k)
                    for (int r = 0; r < ORACLE_R; ++r) {
                        const size_t cell = ((size_t)(i * K + j) * K + k) * ORACLE_R + r;
                        const double w = (ta[i] * tb[j] * tc[k] * pr[cell]) / d[r];
                        const double n = (double)cnt[2 * l + r];
                        ntheta[a * K + i] += w * n;
                        ntheta[b * K + j] += w * n;
                        ntheta[c * K + k] += w * n;
                        npr[cell] += w * n;
                    }
    }
}

/* returns 0, or 1 when some gene has no training link (the reference raises ZeroDivisionError) */
int oracle_em_step(int P, int K, int64_t L, const int64_t *ids, const int64_t *cnt,
                   const double *theta, const double *pr, double *ntheta, double *npr)
{
    int64_t *deg = (int64_t *)calloc((size_t)P, sizeof(int64_t));
    memset(ntheta, 0, sizeof(double) * (size_t)P * K);
    memset(npr, 0, sizeof(double) * (size_t)K * K * K * ORACLE_R);
    for (int64_t l = 0; l < L; ++l) {
        deg[ids[3 * l]]++;
        deg[ids[3 * l + 1]]++;
        deg[ids[3 * l + 2]]++;
    }
    estep_range(K, 0, L, ids, cnt, theta, pr, ntheta, npr);
    for (int g = 0; g < P; ++g) {
        if (deg[g] == 0) {
            free(deg);
            return 1;
        }
        for (int k = 0; k < K; ++k)
            ntheta[(size_t)g * K + k] /= (double)deg[g];
    }
    free(deg);
    for (int c = 0; c < K * K * K; ++c) {
        double d = ORACLE_EPS;
        for (int r = 0; r < ORACLE_R; ++r)
            d += npr[c * ORACLE_R + r];
        for (int r = 0; r < ORACLE_R; ++r)
            npr[c * ORACLE_R + r] /= d;
    }
    return 0;
}

double oracle_loglik(int P, int K, int64_t L, const int64_t *ids, const int64_t *cnt,
                     const double *theta, const double *pr)
{
    (void)P;
    double total = 0.0;
    for (int64_t l = 0; l < L; ++l) {
        const double *ta = theta + ids[3 * l] * K, *tb = theta + ids[3 * l + 1] * K,
                     *tc = theta + ids[3 * l + 2] * K;
        double d[ORACLE_R] = {ORACLE_EPS, ORACLE_EPS};
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j)
                for (int k = 0; k < K; ++k)
                    for (int r = 0; r < ORACLE_R; ++r)
                        d[r] += ta[i] * tb[j] * tc[k] * pr[((i * K + j) * K + k) * ORACLE_R + r];
        for (int r = 0; r < ORACLE_R; ++r)
            total += (double)cnt[2 * l + r] * log(d[r]);
    }
    return total;
}

/* Multi-threaded E-step statistics (unnormalised): one contiguous link block per thread, private
 * accumulators, summed in thread order.  Used as the "all host cores" CPU baseline leg. */
int oracle_em_stats_mt(int P, int K, int64_t L, const int64_t *ids, const int64_t *cnt,
                       const double *theta, const double *pr, double *ntheta, double *npr, int threads)
{
    const size_t nt = (size_t)P * K, np_ = (size_t)K * K * K * ORACLE_R;
    if (threads < 1) threads = 1;
    double *priv = (double *)calloc((nt + np_) * (size_t)threads, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
#ifdef _OPENMP
        const int t = omp_get_thread_num(), T = omp_get_num_threads();
#else
        const int t = 0, T = 1;
#endif
        const int64_t lo = L * t / T, hi = L * (t + 1) / T;
        estep_range(K, lo, hi, ids, cnt, theta, pr, priv + (nt + np_) * t, priv + (nt + np_) * t + nt);
    }
    memset(ntheta, 0, sizeof(double) * nt);
    memset(npr, 0, sizeof(double) * np_);
    for (int t = 0; t < threads; ++t) {
        const double *p = priv + (nt + np_) * t;
        for (size_t i = 0; i < nt; ++i) ntheta[i] += p[i];
        for (size_t i = 0; i < np_; ++i) npr[i] += p[nt + i];
    }
    free(priv);
    return 0;
}

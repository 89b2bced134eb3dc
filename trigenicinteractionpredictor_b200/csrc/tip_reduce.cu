// Cross-sample reduction of held-out scores (testResultsReducer.py:160-184 of the reference): for every test
// triplet, over the samples (random restarts) that scored it, the mean, the reference's median and the population
// standard deviation - in the reference's own operation order, so mean and median are bit-identical to CPython
// and the deviation differs at most by the rounding of its `** 2` (libm pow) against a plain product.
//
// Layout: scores[S][T] (sample-major, so a warp reads 32 consecutive triplets of one sample: coalesced);
// thread = triplet.  HBM-bound: 8*S bytes read + 8*S written (sorted copy) + 24 bytes of results per triplet.
#include "tip_common.cuh"

namespace tip {

__global__ void __launch_bounds__(256)
    reduce_samples_kernel(int S, int64_t T, const double *__restrict__ scores, const int32_t *__restrict__ n_valid,
                          double *__restrict__ sorted, double *__restrict__ mean_out, double *__restrict__ median_out,
                          double *__restrict__ std_out)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
        int n = n_valid ? n_valid[t] : S;
        if (n > S) n = S;
        if (n <= 0) {
            mean_out[t] = median_out[t] = std_out[t] = 0.0;
            continue;
        }
        // mean: `sum_total += x` in sample order, then one division (testResultsReducer.py:166-171)
        double sum = 0.0;
        for (int j = 0; j < n; ++j) {
            const double x = scores[(int64_t)j * T + t];
            sum = __dadd_rn(sum, x);
            // insertion into the ascending copy (list.sort(), :174)
            int i = j;
            while (i > 0) {
                const double y = sorted[(int64_t)(i - 1) * T + t];
                if (!(y > x)) break;
                sorted[(int64_t)i * T + t] = y;
                --i;
            }
            sorted[(int64_t)i * T + t] = x;
        }
        const double mean = __ddiv_rn(sum, (double)n);
        // median (:175-180): odd n takes element round(n / 2) - Python 3 rounds halves to even, so the index is
        // n//2 when that is even and n//2 + 1 when it is odd (n = 3 -> 2, 5 -> 2, 7 -> 4); even n averages the centre
        double median;
        if (n & 1) {
            const int h = n / 2;
            const int idx = (h & 1) ? h + 1 : h;
            median = sorted[(int64_t)(idx < n ? idx : n - 1) * T + t];
        } else {
            const int h = n / 2;
            median = __ddiv_rn(__dadd_rn(sorted[(int64_t)(h - 1) * T + t], sorted[(int64_t)h * T + t]), 2.0);
        }
        // population standard deviation over the SORTED values (the list was sorted in place, :183-186)
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
            const double d = __dsub_rn(sorted[(int64_t)j * T + t], mean);
            acc = __dadd_rn(acc, __dmul_rn(d, d));
        }
        mean_out[t] = mean;
        median_out[t] = median;
        std_out[t] = sqrt(__ddiv_rn(acc, (double)n));
    }
}

}  // namespace tip

using namespace tip;

extern "C" int tip_reduce_samples(int S, int64_t T, const double *d_scores, const int32_t *d_n, double *d_sorted,
                                  double *d_mean, double *d_median, double *d_std, void *stream)
{
    TIP_REQUIRE(S >= 1 && T >= 0, "tip_reduce_samples: need S >= 1 and T >= 0 (got S=%d T=%lld)", S, (long long)T);
    if (T == 0) return 0;
    TIP_REQUIRE(d_scores && d_sorted && d_mean && d_median && d_std, "tip_reduce_samples: null pointer");
    const int64_t want = (T + 255) / 256;
    const int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
    reduce_samples_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(S, T, d_scores, d_n, d_sorted, d_mean,
                                                                                  d_median, d_std);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Cross-sample reduction of held-out scores (testResultsReducer.py:160-184 of the reference): for every test
// triplet, over the samples (random restarts) that scored it, the mean, the reference's median and the population
// standard deviation - in the reference's own operation order, so mean and median are bit-identical to CPython
// and the deviation differs at most by the rounding of its `** 2` (libm pow) against a plain product.
//
// Layout: scores[S][T] (sample-major, so a warp reads 32 consecutive triplets of one sample: coalesced);
// thread = triplet.  HBM-bound: 8*S bytes read + 8*S written (sorted copy) + 24 bytes of results per triplet.
#include "tip_common.cuh"

namespace tip {

// SMEM = true: the column of a triplet is sorted in shared memory (col[j][thread], conflict-free) and written back once,
// coalesced; SMEM = false (S too large for shared memory): the insertion sort works in the global scratch copy.
template <bool SMEM>
__global__ void __launch_bounds__(256)
    reduce_samples_kernel(int S, int64_t T, const double *__restrict__ scores, const int32_t *__restrict__ n_valid,
                          double *__restrict__ sorted, double *__restrict__ mean_out, double *__restrict__ median_out,
                          double *__restrict__ std_out)
{
    extern __shared__ double col_sm[];   // [S][blockDim.x] when SMEM
    const int64_t n_iter = (T + blockDim.x - 1) / blockDim.x;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {
        const int64_t t = it * blockDim.x + threadIdx.x;
        if (t >= T) continue;
        // element j of this triplet's ascending copy
        double *base = SMEM ? col_sm + threadIdx.x : sorted + t;
        const int64_t stride = SMEM ? (int64_t)blockDim.x : T;
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
                const double y = base[(int64_t)(i - 1) * stride];
                if (!(y > x)) break;
                base[(int64_t)i * stride] = y;
                --i;
            }
            base[(int64_t)i * stride] = x;
        }
        const double mean = __ddiv_rn(sum, (double)n);
        // median (:175-180): odd n takes element round(n / 2) - Python 3 rounds halves to even, so the index is
        // n//2 when that is even and n//2 + 1 when it is odd (n = 3 -> 2, 5 -> 2, 7 -> 4); even n averages the centre
        double median;
        if (n & 1) {
            const int h = n / 2;
            const int idx = (h & 1) ? h + 1 : h;
            median = base[(int64_t)(idx < n ? idx : n - 1) * stride];
        } else {
            const int h = n / 2;
            median = __ddiv_rn(__dadd_rn(base[(int64_t)(h - 1) * stride], base[(int64_t)h * stride]), 2.0);
        }
        // population standard deviation over the SORTED values (the list was sorted in place, :183-186)
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
            const double v = base[(int64_t)j * stride];
            if (SMEM) sorted[(int64_t)j * T + t] = v;   // the ascending values are an output too
            const double d = __dsub_rn(v, mean);
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
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // threads per CTA so that the columns of a CTA fit ~48 KB of shared memory (4 CTAs per SM); 32 at least
    int threads = (int)(48 * 1024 / (8 * (size_t)S)) / 32 * 32;
    threads = threads > 256 ? 256 : threads;
    if (threads >= 32) {
        const size_t smem = (size_t)S * threads * sizeof(double);
        const int64_t want = (T + threads - 1) / threads;
        const int grid = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
        reduce_samples_kernel<true><<<grid, threads, smem, st>>>(S, T, d_scores, d_n, d_sorted, d_mean, d_median, d_std);
    } else {
        const int64_t want = (T + 255) / 256;
        const int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
        reduce_samples_kernel<false><<<grid, 256, 0, st>>>(S, T, d_scores, d_n, d_sorted, d_mean, d_median, d_std);
    }
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

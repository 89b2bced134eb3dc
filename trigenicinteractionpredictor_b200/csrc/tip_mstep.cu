// M-step of Model.make_iteration (TIP.py:1016-1043): normalise the summed statistics in place.
#include "tip_common.cuh"

namespace tip {

// skip_flag (optional): a device word that, when non-zero, makes the kernel leave theta / p untouched - set by the
// verification of a streamed E-step whose statistics cannot be trusted (tip_em.cu, StreamArrive)
__global__ void normalise_kernel(int P, int K, const double *__restrict__ stats, const int32_t *__restrict__ deg,
                                 double *__restrict__ theta, double *__restrict__ p, const unsigned long long *skip_flag)
{
    if (skip_flag != nullptr && *skip_flag != 0ull) return;
    const int64_t nth = (int64_t)P * K;
    const int K3 = K * K * K;
    const double *S = stats + stats_off_S(P, K);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nth + K3; e += (int64_t)gridDim.x * blockDim.x) {
        if (e < nth) {
            // TIP.py:1016-1018: ntheta[i][k] /= float(counter[i])
            theta[e] = stats[e] / (double)deg[e / K];
        } else {
            // TIP.py:1021-1028 with npr = p * S: d = eps; d += npr[0]; d += npr[1]; npr[r] /= d
            const int cell = (int)(e - nth);
            const double n0 = p[2 * cell] * S[cell];
            const double n1 = p[2 * cell + 1] * S[K3 + cell];
            double d = TIP_EPS;
            d += n0;
            d += n1;
            p[2 * cell] = n0 / d;
            p[2 * cell + 1] = n1 / d;
        }
    }
}

int launch_normalise(int P, int K, const double *stats, const int32_t *deg, double *theta, double *p, cudaStream_t st,
                     const unsigned long long *skip_flag)
{
    const int64_t n = (int64_t)P * K + (int64_t)K * K * K;
    const int threads = 256;
    int64_t want = (n + threads - 1) / threads;
    int grid = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
    if (grid < 1) grid = 1;
    normalise_kernel<<<grid, threads, 0, st>>>(P, K, stats, deg, theta, p, skip_flag);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace tip

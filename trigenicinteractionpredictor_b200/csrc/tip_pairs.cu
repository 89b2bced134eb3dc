// Pair links of the digenic extension (SURVEY f-4): the reference's src/TrigenicInteractionPredictor_23.py adds links
// between TWO genes, `dlinks`, that share theta with the triplets and have their own rating tensor qr[K][K][R]
// (make_iteration _23.py:1607-1635, 1653-1659; compute_likelihood _23.py:1551-1560).  For a pair (a, b) with counts n:
//     d_r = eps + sum_ij th_a[i] th_b[j] q_ij,r ;   s_r = n_r / d_r
//     Ntheta[a][i] += s_r th_a[i] sum_j th_b[j] q_ij,r ;   Ntheta[b][j] += s_r th_b[j] sum_i th_a[i] q_ij,r
//     Sq[r][i][j]  += s_r th_a[i] th_b[j]                      (nqr = q * Sq, like npr = p * S for the triplets)
// K^2 cells per link instead of K^3: one warp per pair, lane = group index (K <= 32), the partner's theta by shuffles,
// q through L1; the q statistic is privatised per CTA in shared memory and flushed once.  The statistics land in the SAME
// Ntheta buffer as the triplet E-step (tip_pairs_step runs between tip_em_step and tip_normalise), so that the M-step
// divides by the degree over triplets AND pairs (_23.py:1640-1643).
#include "tip_common.cuh"

namespace tip {

constexpr int kPairThreads = 256;

// mode 0: statistics (ntheta, Sq);  mode 1: log-likelihood only (ll)
__global__ void __launch_bounds__(kPairThreads) pairs_kernel(int K, const int4 *__restrict__ pairs, int64_t n_pairs,
                                                             const double *__restrict__ theta, const double *__restrict__ q,
                                                             double *__restrict__ ntheta, double *__restrict__ Sq,
                                                             double *__restrict__ ll, int mode)
{
    extern __shared__ double sq_sm[];   // [2][K*K]
    const int KK = K * K, lane = threadIdx.x & 31;
    if (mode == 0) {
        for (int e = threadIdx.x; e < 2 * KK; e += blockDim.x) sq_sm[e] = 0.0;
        __syncthreads();
    }
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double ll_acc = 0.0;
    for (int64_t w = warp0; w < n_pairs; w += n_warps) {
        const int4 v = pairs[w];
        const double ta = lane < K ? __ldg(theta + (int64_t)v.x * K + lane) : 0.0;
        const double tb = lane < K ? __ldg(theta + (int64_t)v.y * K + lane) : 0.0;
        for (int r = 0; r < 2; ++r) {
            const int n = r ? v.w : v.z;
            if (n == 0) continue;                       // a term with n_r = 0 contributes exactly 0 (SURVEY parity trap 3)
            double u = 0.0, vv = 0.0;                   // u_i = sum_j tb_j q_ij,r (lane = i);  v_j = sum_i ta_i q_ij,r (lane = j)
            for (int j = 0; j < K; ++j) {
                const double tbj = __shfl_sync(0xffffffffu, tb, j), taj = __shfl_sync(0xffffffffu, ta, j);
                if (lane < K) {
                    u = fma(tbj, __ldg(q + ((int64_t)lane * K + j) * 2 + r), u);
                    vv = fma(taj, __ldg(q + ((int64_t)j * K + lane) * 2 + r), vv);
                }
            }
            const double d = TIP_EPS + warp_sum(ta * u);
            if (mode == 1) {
                if (lane == 0) ll_acc += (double)n * log(d);
                continue;
            }
            const double s = (double)n / d;
            if (lane < K) {
                red_add_f64(ntheta + (int64_t)v.x * K + lane, s * ta * u);
                red_add_f64(ntheta + (int64_t)v.y * K + lane, s * tb * vv);
            }
            for (int j = 0; j < K; ++j) {
                const double tbj = __shfl_sync(0xffffffffu, tb, j);
                if (lane < K) atomicAdd(&sq_sm[r * KK + lane * K + j], s * ta * tbj);
            }
        }
    }
    if (mode == 1) {
        if (lane == 0 && ll_acc != 0.0) red_add_f64(ll, ll_acc);
        return;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * KK; e += blockDim.x)
        if (sq_sm[e] != 0.0) red_add_f64(Sq + e, sq_sm[e]);
}

// q[i][j][r] <- nqr / (eps + nqr_0 + nqr_1) with nqr = q * Sq   (_23.py:1653-1659, in the reference's operation order)
__global__ void pairs_normalise_kernel(int K, const double *__restrict__ Sq, double *__restrict__ q)
{
    const int KK = K * K;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < KK; c += gridDim.x * blockDim.x) {
        const double n0 = q[2 * c] * Sq[c], n1 = q[2 * c + 1] * Sq[KK + c];
        double d = TIP_EPS;
        d += n0;
        d += n1;
        q[2 * c] = n0 / d;
        q[2 * c + 1] = n1 / d;
    }
}

static int launch_pairs(int K, const int4 *pairs, int64_t n_pairs, const double *theta, const double *q, double *ntheta, double *Sq,
                        double *ll, int mode, cudaStream_t st)
{
    if (n_pairs <= 0) return 0;
    const int64_t want = (n_pairs * 32 + kPairThreads - 1) / kPairThreads;
    const int grid = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
    const size_t smem = mode == 0 ? sizeof(double) * 2 * (size_t)K * K : 0;
    pairs_kernel<<<grid, kPairThreads, smem, st>>>(K, pairs, n_pairs, theta, q, ntheta, Sq, ll, mode);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace tip

using namespace tip;

extern "C" int tip_pairs_step(int P, int K, const void *d_pairs, int64_t n_pairs, const double *d_theta, const double *d_q,
                              double *d_stats, double *d_sq, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TIP_REQUIRE(P > 0 && K >= 1 && K <= TIP_MAX_K && n_pairs >= 0 && d_theta && d_q && d_stats && d_sq && (n_pairs == 0 || d_pairs),
                "tip_pairs_step: bad arguments (1 <= K <= %d)", TIP_MAX_K);
    TIP_CHECK_CUDA(cudaMemsetAsync(d_sq, 0, sizeof(double) * 2 * (size_t)K * K, st));
    return launch_pairs(K, reinterpret_cast<const int4 *>(d_pairs), n_pairs, d_theta, d_q, d_stats, d_sq, nullptr, 0, st);
}

extern "C" int tip_pairs_normalise(int K, const double *d_sq, double *d_q, void *stream)
{
    TIP_REQUIRE(K >= 1 && K <= TIP_MAX_K && d_sq && d_q, "tip_pairs_normalise: bad arguments");
    pairs_normalise_kernel<<<(K * K + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(K, d_sq, d_q);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tip_pairs_loglik(int P, int K, const void *d_pairs, int64_t n_pairs, const double *d_theta, const double *d_q,
                                double *d_out, void *stream)
{
    TIP_REQUIRE(P > 0 && K >= 1 && K <= TIP_MAX_K && n_pairs >= 0 && d_theta && d_q && d_out && (n_pairs == 0 || d_pairs),
                "tip_pairs_loglik: bad arguments");
    return launch_pairs(K, reinterpret_cast<const int4 *>(d_pairs), n_pairs, d_theta, d_q, nullptr, nullptr, d_out, 1,
                        reinterpret_cast<cudaStream_t>(stream));
}

// Any-K (1..32) kernels: E-step statistics, log-likelihood, held-out scoring.
//
// These are the general path (run-time K, p read through L1/L2).  The K-specialised fused kernel
// in tip_em.cu replaces gen_estep_kernel + gen_pstat_kernel where it exists; log-likelihood and
// scoring always run here (they are <1/3 of an EM step and run every `fcheck` iterations).
//
// Per packed row (a, b, c, count, r), with q[ab] = sum_c p[a][b][c][r] * th_c[c]:
//   x[ab]  = th_a[a] th_b[b] q[ab]           d = eps + sum x            (TIP.py:990-1000)
//   s      = count / d
//   Ntheta[a][i] += s * sum_j x[i][j]        Ntheta[b][j] += s * sum_i x[i][j]
//   Ntheta[c][k] += s * th_c[k] * sum_ij th_a[i] th_b[j] p[i][j][k][r]       (TIP.py:1009-1011)
//   S[r][ijk]    += s * th_a[i] th_b[j] th_c[k]                              (TIP.py:1012, npr = p*S)
#include "tip_common.cuh"

namespace tip {

constexpr int kGenWarps = 8;  // warps per CTA for the warp-per-row kernels

// warp computes x[pair] (and optionally ab[pair]) into shared memory; returns sum_pairs x (no eps)
__device__ __forceinline__ double gen_row_products(int K, int r, const double *__restrict__ p, const double *ta,
                                                   const double *tb, const double *tc, double *x, double *ab_out,
                                                   int lane)
{
    const int KK = K * K;
    double part = 0.0;
    for (int pair = lane; pair < KK; pair += kWarp) {
        const int i = pair / K, j = pair - i * K;
        const double *pp = p + ((int64_t)pair * K) * 2 + r;
        double q = 0.0;
        for (int k = 0; k < K; ++k) q = fma(__ldg(pp + 2 * k), tc[k], q);
        const double ab = ta[i] * tb[j];
        const double xv = ab * q;
        if (x) x[pair] = xv;
        if (ab_out) ab_out[pair] = ab;
        part += xv;
    }
    return warp_sum(part);
}

__device__ __forceinline__ void gen_load_thetas(int K, const double *__restrict__ theta, int a, int b, int c,
                                                double *ta, double *tb, double *tc, int lane)
{
    if (lane < K) {
        ta[lane] = __ldg(theta + (int64_t)a * K + lane);
        tb[lane] = __ldg(theta + (int64_t)b * K + lane);
        tc[lane] = __ldg(theta + (int64_t)c * K + lane);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// G1: one warp per row: theta statistics, log-likelihood by-product, s[row] for the p statistics
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenWarps *kWarp)
    gen_estep_kernel(int P, int K, const int4 *__restrict__ rows, int64_t n_rows, const double *__restrict__ theta,
                     const double *__restrict__ p, double *__restrict__ stats, double *__restrict__ s_out)
{
    extern __shared__ double sm[];
    const int KK = K * K;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    double *base = sm + (size_t)warp * (3 * kWarp + 2 * KK);
    double *ta = base, *tb = base + kWarp, *tc = base + 2 * kWarp, *x = base + 3 * kWarp, *ab = x + KK;
    double ll = 0.0;
    const int64_t wstride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t row = (int64_t)blockIdx.x * kGenWarps + warp; row < n_rows; row += wstride) {
        const int4 lk = rows[row];
        const int cnt = row_count(lk.w), r = row_rating(lk.w);
        if (cnt == 0) {
            if (lane == 0) s_out[row] = 0.0;
            continue;
        }
        gen_load_thetas(K, theta, lk.x, lk.y, lk.z, ta, tb, tc, lane);
        const double d = TIP_EPS + gen_row_products(K, r, p, ta, tb, tc, x, ab, lane);
        const double s = (double)cnt / d;
        __syncwarp();
        if (lane == 0) {
            s_out[row] = s;
            ll += (double)cnt * log(d);
        }
        if (lane < K) {
            double ua = 0.0, ub = 0.0, wc = 0.0;
            for (int j = 0; j < K; ++j) ua += x[lane * K + j];
            for (int i = 0; i < K; ++i) ub += x[i * K + lane];
            const double *pc = p + 2 * lane + r;
            for (int pair = 0; pair < KK; ++pair) wc = fma(ab[pair], __ldg(pc + (int64_t)pair * K * 2), wc);
            red_add_f64(stats + (int64_t)lk.x * K + lane, s * ua);
            red_add_f64(stats + (int64_t)lk.y * K + lane, s * ub);
            red_add_f64(stats + (int64_t)lk.z * K + lane, s * tc[lane] * wc);
        }
        __syncwarp();
    }
    if (lane == 0 && ll != 0.0) red_add_f64(stats + stats_off_ll(P, K), ll);
}

// ---------------------------------------------------------------------------------------------
// G2: S[r][cell] += sum_rows s * th_a th_b th_c.   grid = (cell tiles, row chunks)
// Each thread owns kCellsPerThread cells; rows are staged through shared memory in batches.
// ---------------------------------------------------------------------------------------------
constexpr int kPsThreads = 256;
constexpr int kCellsPerThread = 4;
constexpr int kPsBatch = 32;  // rows per shared-memory batch

__global__ void __launch_bounds__(kPsThreads)
    gen_pstat_kernel(int P, int K, const int4 *__restrict__ rows, int64_t n_rows, int64_t split,
                     const double *__restrict__ theta, const double *__restrict__ s_in, double *__restrict__ stats)
{
    extern __shared__ double sm[];  // [kPsBatch][3][K]: s*th_a, th_b, th_c
    const int r = blockIdx.z;       // rating block: rows [0,split) have rating 0, [split,n_rows) rating 1
    const int64_t row_begin = r == 0 ? 0 : split, row_end = r == 0 ? split : n_rows;
    const int K3 = K * K * K;
    int ci[kCellsPerThread], cj[kCellsPerThread], ck[kCellsPerThread];
    double acc[kCellsPerThread];
    const int cell0 = (blockIdx.x * kPsThreads + threadIdx.x) * kCellsPerThread;
#pragma unroll
    for (int t = 0; t < kCellsPerThread; ++t) {
        int cell = cell0 + t;
        if (cell >= K3) cell = K3 - 1;  // clamped duplicates are discarded at the end
        ci[t] = cell / (K * K);
        cj[t] = (cell / K) % K;
        ck[t] = cell % K;
        acc[t] = 0.0;
    }
    const int64_t n = row_end - row_begin;
    const int64_t per0 = (n + gridDim.y - 1) / gridDim.y;
    const int64_t per = ((per0 + kPsBatch - 1) / kPsBatch) * kPsBatch;
    const int64_t lo = row_begin + per * blockIdx.y;
    const int64_t hi = (lo + per < row_end) ? lo + per : row_end;
    for (int64_t b0 = lo; b0 < hi; b0 += kPsBatch) {
        const int nb = (int)((hi - b0 < kPsBatch) ? hi - b0 : kPsBatch);
        __syncthreads();
        for (int e = threadIdx.x; e < nb * 3 * K; e += kPsThreads) {
            const int l = e / (3 * K), rem = e - l * 3 * K, slot = rem / K, k = rem - slot * K;
            const int4 lk = rows[b0 + l];
            const int g = slot == 0 ? lk.x : (slot == 1 ? lk.y : lk.z);
            double v = __ldg(theta + (int64_t)g * K + k);
            if (slot == 0) v *= s_in[b0 + l];
            sm[e] = v;
        }
        __syncthreads();
        for (int l = 0; l < nb; ++l) {
            const double *ra = sm + l * 3 * K, *rb = ra + K, *rc = rb + K;
#pragma unroll
            for (int t = 0; t < kCellsPerThread; ++t) acc[t] = fma(ra[ci[t]] * rb[cj[t]], rc[ck[t]], acc[t]);
        }
    }
    double *S = stats + stats_off_S(P, K) + (int64_t)r * K3;
#pragma unroll
    for (int t = 0; t < kCellsPerThread; ++t)
        if (cell0 + t < K3 && acc[t] != 0.0) red_add_f64(S + cell0 + t, acc[t]);
}

int launch_em_generic(int P, int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta,
                      const double *p, double *stats, double *s_ws, cudaStream_t st)
{
    const int KK = K * K;
    const size_t smem1 = (size_t)kGenWarps * (3 * kWarp + 2 * KK) * sizeof(double);
    static bool attr_done = false;
    if (!attr_done) {
        TIP_CHECK_CUDA(cudaFuncSetAttribute(gen_estep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    int64_t want = (n_rows + kGenWarps - 1) / kGenWarps;
    int grid = (int)((want < (int64_t)sm_count() * 4) ? want : (int64_t)sm_count() * 4);
    if (grid < 1) grid = 1;
    gen_estep_kernel<<<grid, kGenWarps * kWarp, smem1, st>>>(P, K, rows, n_rows, theta, p, stats, s_ws);
    TIP_CHECK_CUDA(cudaGetLastError());
    const int K3 = K * K * K;
    const int cell_tiles = (K3 + kPsThreads * kCellsPerThread - 1) / (kPsThreads * kCellsPerThread);
    int chunks = (sm_count() * 4 + cell_tiles - 1) / cell_tiles;
    int64_t max_chunks = (n_rows + kPsBatch - 1) / kPsBatch;
    if (chunks > max_chunks) chunks = (int)max_chunks;
    if (chunks < 1) chunks = 1;
    const size_t smem2 = (size_t)kPsBatch * 3 * K * sizeof(double);
    gen_pstat_kernel<<<dim3(cell_tiles, chunks, 2), kPsThreads, smem2, st>>>(P, K, rows, n_rows, n_rows_r0, theta,
                                                                              s_ws, stats);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// log-likelihood: warp per row, per-CTA partials, last CTA sums them in index order
// ---------------------------------------------------------------------------------------------
constexpr int kLlMaxBlocks = 4096;

__global__ void __launch_bounds__(kGenWarps *kWarp)
    gen_loglik_kernel(int K, const int4 *__restrict__ rows, int64_t n_rows, const double *__restrict__ theta,
                      const double *__restrict__ p, double *__restrict__ partials, unsigned *__restrict__ counter,
                      double *__restrict__ out)
{
    __shared__ double sth[kGenWarps][3 * kWarp];
    __shared__ double wsum[kGenWarps];
    __shared__ bool last;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    double *ta = sth[warp], *tb = ta + kWarp, *tc = tb + kWarp;
    double ll = 0.0;
    // contiguous row range per warp so the summation order is a function of (n_rows, grid) only
    const int64_t nw = (int64_t)gridDim.x * kGenWarps;
    const int64_t per = (n_rows + nw - 1) / nw;
    const int64_t w = (int64_t)blockIdx.x * kGenWarps + warp;
    const int64_t lo = w * per, hi = (lo + per < n_rows) ? lo + per : n_rows;
    for (int64_t row = lo; row < hi; ++row) {
        const int4 lk = rows[row];
        const int cnt = row_count(lk.w);
        if (cnt == 0) continue;
        gen_load_thetas(K, theta, lk.x, lk.y, lk.z, ta, tb, tc, lane);
        const double d = TIP_EPS + gen_row_products(K, row_rating(lk.w), p, ta, tb, tc, nullptr, nullptr, lane);
        ll += (double)cnt * log(d);
        __syncwarp();
    }
    if (lane == 0) wsum[warp] = ll;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < kGenWarps; ++i) t += wsum[i];
        partials[blockIdx.x] = t;
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned i = 0; i < gridDim.x; ++i) t += reinterpret_cast<volatile double *>(partials)[i];
        *out = t;
        *counter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// scoring: warp per test triplet, rating 1, no eps
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenWarps *kWarp)
    gen_score_kernel(int K, const int32_t *__restrict__ g1, const int32_t *__restrict__ g2,
                     const int32_t *__restrict__ g3, int64_t T, const double *__restrict__ theta,
                     const double *__restrict__ p, double *__restrict__ scores)
{
    __shared__ double sth[kGenWarps][3 * kWarp];
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    double *ta = sth[warp], *tb = ta + kWarp, *tc = tb + kWarp;
    const int64_t wstride = (int64_t)gridDim.x * kGenWarps;
    for (int64_t t = (int64_t)blockIdx.x * kGenWarps + warp; t < T; t += wstride) {
        gen_load_thetas(K, theta, g1[t], g2[t], g3[t], ta, tb, tc, lane);
        const double v = gen_row_products(K, 1, p, ta, tb, tc, nullptr, nullptr, lane);
        if (lane == 0) scores[t] = v;
        __syncwarp();
    }
}

int launch_loglik(int P, int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta, const double *p,
                  double *out, void *ws, bool force_generic, cudaStream_t st)
{
    double *partials = reinterpret_cast<double *>(ws);
    unsigned *counter = reinterpret_cast<unsigned *>(partials + kLlMaxBlocks);
    if (!force_generic && n_rows_r0 >= 0) {
        TIP_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
        bool handled = false;
        const int rc = launch_loglik_tuned(K, rows, n_rows, n_rows_r0, theta, p, out, partials, counter, kLlMaxBlocks, st,
                                           &handled);
        if (rc != 0 || handled) return rc;
        if (K > 10 && P > 0) {
            // Z workspace follows the partials/counter header (see loglik_ws_bytes)
            double *Zws = reinterpret_cast<double *>(reinterpret_cast<char *>(ws) + loglik_ws_bytes(0, 0));
            return launch_loglik_seg(P, K, rows, n_rows, theta, p, out, partials, counter, kLlMaxBlocks, Zws, st);
        }
    }
    int64_t want = (n_rows + kGenWarps - 1) / kGenWarps;
    int grid = (int)((want < (int64_t)sm_count() * 4) ? want : (int64_t)sm_count() * 4);
    if (grid > kLlMaxBlocks) grid = kLlMaxBlocks;
    if (grid < 1) grid = 1;
    TIP_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
    gen_loglik_kernel<<<grid, kGenWarps * kWarp, 0, st>>>(K, rows, n_rows, theta, p, partials, counter, out);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_score(int K, const int32_t *g1, const int32_t *g2, const int32_t *g3, int64_t T, const double *theta,
                 const double *p, double *scores, cudaStream_t st)
{
    if (T == 0) return 0;
    int64_t want = (T + kGenWarps - 1) / kGenWarps;
    int grid = (int)((want < (int64_t)sm_count() * 8) ? want : (int64_t)sm_count() * 8);
    gen_score_kernel<<<grid, kGenWarps * kWarp, 0, st>>>(K, g1, g2, g3, T, theta, p, scores);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

size_t loglik_ws_bytes(int P, int K)
{
    const size_t header = (kLlMaxBlocks * sizeof(double) + 64 + 255) / 256 * 256;
    return header + loglik_seg_workspace_bytes(P, K);
}

}  // namespace tip

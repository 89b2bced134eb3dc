// Shared helpers for libtip.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tip.h"

namespace tip {

constexpr int kWarp = 32;
constexpr int kNumSM_B200 = 148;

void set_error(const char *fmt, ...);

#define TIP_CHECK_CUDA(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::tip::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                        \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

#define TIP_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            ::tip::set_error(__VA_ARGS__); \
            return -1;                     \
        }                                  \
    } while (0)

inline int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kNumSM_B200;
    }
    return n;
}

// packed row helpers: w = (count << 1) | rating
__device__ __forceinline__ int row_rating(int w) { return w & 1; }
__device__ __forceinline__ int row_count(int w) { return w >> 1; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fire-and-forget fp64 reduction (RED.E.ADD.F64 in SASS)
__device__ __forceinline__ void red_add_f64(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// stats buffer offsets
__host__ __device__ inline int64_t stats_off_S(int P, int K) { return (int64_t)P * K; }
__host__ __device__ inline int64_t stats_off_ll(int P, int K) { return (int64_t)P * K + 2ll * K * K * K; }

// launchers implemented across the .cu files
int launch_em_generic(int P, int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta,
                      const double *p, double *stats, double *s_ws, cudaStream_t st);
int launch_em_tuned(int P, int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta,
                    const double *p, double *stats, double *ws, bool with_ll, bool f32, bool seg, cudaStream_t st,
                    bool *handled, int phases = 7);
size_t em_tuned_workspace_bytes(int P, int K, bool seg);
int launch_em_seg3(int P, int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta, const double *p,
                   double *stats, double *ws, bool gather_l1, cudaStream_t st);
size_t em_seg3_workspace_bytes(int P, int K, int64_t n_rows);
int order_rows_parts(const void *d_rows, int64_t n_rows, int64_t n_rows_r0, void *d_ws, size_t ws_bytes, void *d_rows_bc,
                     cudaStream_t st, int parts, int P);
void seg3_wait_before_bc(cudaEvent_t ev);
cudaEvent_t seg3_take_bc_wait();
bool em_streamed_available(int K, bool with_ll, bool f32, bool seg);
int launch_em_streamed(int P, int K, const void *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta, double *stats,
                       double *ws, unsigned *err, unsigned long long *chk, bool compact, cudaStream_t st);
int launch_stream_verify(const void *rows, int64_t n_rows, bool compact, unsigned long long *chk, unsigned *err, cudaStream_t st);
int launch_loglik(int P, int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta, const double *p,
                  double *out, void *ws, bool force_generic, cudaStream_t st);
int launch_loglik_seg(int P, int K, const int4 *rows, int64_t n_rows, const double *theta, const double *p, double *out,
                      double *partials, unsigned *counter, int max_blocks, double *Zws, cudaStream_t st);
size_t loglik_seg_workspace_bytes(int P, int K);
int launch_loglik_tuned(int K, const int4 *rows, int64_t n_rows, int64_t n_rows_r0, const double *theta, const double *p,
                        double *out, double *partials, unsigned *counter, int max_blocks, cudaStream_t st, bool *handled);
int launch_score(int K, const int32_t *g1, const int32_t *g2, const int32_t *g3, int64_t T, const double *theta,
                 const double *p, double *scores, cudaStream_t st);
size_t loglik_ws_bytes(int P, int K);
int launch_normalise(int P, int K, const double *stats, const int32_t *deg, double *theta, double *p,
                     cudaStream_t st, const unsigned long long *skip_flag = nullptr);

}  // namespace tip

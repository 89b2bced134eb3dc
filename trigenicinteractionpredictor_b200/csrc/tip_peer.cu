// Link-shard exchange over NVLink peer memory: the statistics of every rank are summed by the M-step kernel
// itself through peer (P2P-mapped) pointers, replacing the NCCL allreduce between E-step and M-step.
//
//   tip_ipc_export / tip_ipc_import   share a device buffer between the per-GPU processes (CUDA IPC)
//   tip_peer_barrier                  all ranks have finished writing this iteration's statistics
//   tip_normalise_peers               theta, p <- M-step( sum over ranks, in rank order, of peer statistics )
//
// Every rank adds the N buffers in the same order, so the replicas of theta and p stay bit-identical.
// Statistics are double-buffered by the caller (iteration i uses buffer i & 1), which makes ONE barrier per
// iteration sufficient: a rank can run at most one iteration ahead of the slowest reader.
#include <cuda.h>

#include "tip_common.cuh"

namespace tip {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
    const double *p[kMaxPeers];
};
struct FlagPtrs {
    unsigned long long *p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *addr, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *addr)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}

// one CTA, one thread per peer.  flags[r] points at rank r's flag array (nranks slots); slot `rank` of every
// array belongs to this rank.  *epoch is a local counter bumped once per call.
__global__ void peer_barrier_kernel(FlagPtrs flags, unsigned long long *epoch, int rank, int nranks)
{
    __shared__ unsigned long long e;
    if (threadIdx.x == 0) e = *epoch + 1;
    __syncthreads();
    // a barrier that timed out poisons the epoch (~0) FOR GOOD: later barriers neither signal nor pass, so the peers time
    // out as well and every rank's host-side check (PeerExchange.check) raises - never a silent wrap-around to epoch 0
    if (e == 0ull) return;
    __threadfence_system();  // this rank's statistics (written by earlier kernels of the stream) are visible to peers
    const int t = threadIdx.x;
    __shared__ int timed_out;
    if (t == 0) timed_out = 0;
    __syncthreads();
    if (t < nranks) {
        st_release_sys(flags.p[t] + rank, e);                    // tell peer t
        const long long t0 = clock64();
        while (ld_acquire_sys(flags.p[rank] + t) < e) {          // wait for peer t
            if (clock64() - t0 > 20000000000ll) {                // ~10 s: a peer died; do not hang the GPU
                timed_out = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch = timed_out ? ~0ull : e;        // ~0 = poisoned, checked by the host
}

__global__ void normalise_peers_kernel(int P, int K, PeerPtrs stats, int nranks, const int32_t *__restrict__ deg,
                                       double *__restrict__ theta, double *__restrict__ p)
{
    const int64_t nth = (int64_t)P * K;
    const int K3 = K * K * K;
    const int64_t offS = stats_off_S(P, K);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nth + K3; e += (int64_t)gridDim.x * blockDim.x) {
        if (e < nth) {
            double s = 0.0;
            for (int r = 0; r < nranks; ++r) s += stats.p[r][e];
            theta[e] = s / (double)deg[e / K];
        } else {
            const int cell = (int)(e - nth);
            double s0 = 0.0, s1 = 0.0;
            for (int r = 0; r < nranks; ++r) {
                s0 += stats.p[r][offS + cell];
                s1 += stats.p[r][offS + K3 + cell];
            }
            const double n0 = p[2 * cell] * s0, n1 = p[2 * cell + 1] * s1;
            double d = TIP_EPS;
            d += n0;
            d += n1;
            p[2 * cell] = n0 / d;
            p[2 * cell + 1] = n1 / d;
        }
    }
}

}  // namespace tip

using namespace tip;

extern "C" int tip_ipc_export(const void *d_ptr, void *h_handle64, int64_t *h_offset)
{
    TIP_REQUIRE(d_ptr && h_handle64 && h_offset, "tip_ipc_export: null argument");
    // the allocation base is a driver-API query; resolve it at run time so that libtip.so does not link libcuda
    // (the library must also load on GPU-less build hosts)
    typedef CUresult (*GetRangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static GetRangeFn get_range = nullptr;
    if (!get_range) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TIP_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
        TIP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "tip_ipc_export: cuMemGetAddressRange not available");
        get_range = reinterpret_cast<GetRangeFn>(fn);
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    CUresult cr = get_range(&base, &size, (CUdeviceptr)d_ptr);
    TIP_REQUIRE(cr == CUDA_SUCCESS, "tip_ipc_export: cuMemGetAddressRange failed (%d)", (int)cr);
    cudaIpcMemHandle_t h;
    TIP_CHECK_CUDA(cudaIpcGetMemHandle(&h, (void *)base));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(h_handle64, &h, 64);
    *h_offset = (int64_t)((CUdeviceptr)d_ptr - base);
    return 0;
}

extern "C" int tip_ipc_import(const void *h_handle64, int64_t offset, void **d_ptr_out)
{
    TIP_REQUIRE(h_handle64 && d_ptr_out && offset >= 0, "tip_ipc_import: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle64, 64);
    void *base = nullptr;
    TIP_CHECK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr_out = (char *)base + offset;
    return 0;
}

extern "C" int tip_peer_barrier(void *const *h_flag_ptrs, void *d_epoch, int rank, int nranks, void *stream)
{
    TIP_REQUIRE(h_flag_ptrs && d_epoch && nranks >= 1 && nranks <= kMaxPeers && rank >= 0 && rank < nranks,
                "tip_peer_barrier: bad arguments (nranks <= %d)", kMaxPeers);
    FlagPtrs f;
    for (int r = 0; r < nranks; ++r) f.p[r] = reinterpret_cast<unsigned long long *>(h_flag_ptrs[r]);
    peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(f, reinterpret_cast<unsigned long long *>(d_epoch),
                                                                              rank, nranks);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tip_normalise_peers(int P, int K, void *const *h_stats_ptrs, int nranks, const int32_t *d_deg,
                                   double *d_theta, double *d_p, void *stream)
{
    TIP_REQUIRE(P > 0 && K >= 1 && K <= TIP_MAX_K && h_stats_ptrs && d_deg && d_theta && d_p && nranks >= 1 &&
                    nranks <= kMaxPeers,
                "tip_normalise_peers: bad arguments (nranks <= %d)", kMaxPeers);
    PeerPtrs s;
    for (int r = 0; r < nranks; ++r) s.p[r] = reinterpret_cast<const double *>(h_stats_ptrs[r]);
    const int64_t n = (int64_t)P * K + (int64_t)K * K * K;
    const int threads = 256;
    int64_t want = (n + threads - 1) / threads;
    int grid = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
    normalise_peers_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, K, s, nranks, d_deg, d_theta, d_p);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

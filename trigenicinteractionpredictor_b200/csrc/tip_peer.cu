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

// ---------------------------------------------------------------------------------------------
// tip_peer_mstep: the whole exchange + M-step of a link-sharded iteration in ONE kernel.
// Reduce-scatter + all-gather over peer memory instead of every rank reading every buffer:
//   1. signal "my statistics are complete" to every peer, wait for theirs            (flags [0, n) of each rank)
//   2. rank r owns slice r of the element space (theta entries, then p cells): it adds the n statistics buffers of that
//      slice IN RANK ORDER, normalises (theta = Ntheta / deg;  npr = p S, p = npr / (eps + npr0 + npr1)) and STORES the
//      new values into the theta / p arrays of every rank (remote stores, fire and forget)
//   3. the last CTA to finish signals "my slice has been delivered" to every peer and waits for theirs (flags [n, 2n)):
//      when the kernel ends, theta and p of this rank are complete and nobody reads its statistics any more.
// Per rank and iteration 1/n of the statistics is read from each peer and 1/n of the parameters is written to each,
// 2 x 496 KB at cfg2 whatever n (the all-gather formulation reads n x 496 KB).  Each value is computed by one rank
// only, so the replicas of theta and p are bit-identical by construction.  Safe without double buffering: a rank leaves
// the kernel only after every peer has finished reading its statistics (their phase-2 signal), and theta / p are only
// overwritten after every rank's E-step - their only reader - is over (phase 1).
// d_sync: three local words {epoch, CTAs done, timed out}.  A wait of ~10 s poisons the epoch (~0) for good.
struct MstepPtrs {
    const double *stats[kMaxPeers];
    double *theta[kMaxPeers];
    double *p[kMaxPeers];
    unsigned long long *flags[kMaxPeers];
};

__global__ void __launch_bounds__(256) peer_mstep_kernel(int P, int K, MstepPtrs a, int rank, int n,
                                                         const int32_t *__restrict__ deg, unsigned long long *sync)
{
    __shared__ unsigned long long e_sh;
    __shared__ int bad, last;
    if (threadIdx.x == 0) {
        e_sh = sync[0] + 1;
        bad = 0;
        last = 0;
    }
    __syncthreads();
    const unsigned long long e = e_sh;
    if (e == 0ull) return;                                   // poisoned by an earlier timeout
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t < n) {
        __threadfence_system();                              // this rank's statistics (earlier kernels) before the flag
        st_release_sys(a.flags[t] + rank, e);
    }
    if (t < n) {
        const long long t0 = clock64();
        while (ld_acquire_sys(a.flags[rank] + t) < e) {
            if (clock64() - t0 > 20000000000ll) {
                bad = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (bad) {
        if (t == 0) atomicExch(sync + 2, 1ull);
    } else {
        const int64_t nth = (int64_t)P * K;
        const int K3 = K * K * K;
        const int64_t offS = stats_off_S(P, K), total = nth + K3;
        const int64_t per = (total + n - 1) / n, lo = rank * per, hi = (lo + per < total) ? lo + per : total;
        for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + t; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
            if (i < nth) {
                double s = 0.0;
                for (int r = 0; r < n; ++r) s += a.stats[r][i];
                const double v = s / (double)deg[i / K];
                for (int r = 0; r < n; ++r) a.theta[r][i] = v;
            } else {
                const int cell = (int)(i - nth);
                double s0 = 0.0, s1 = 0.0;
                for (int r = 0; r < n; ++r) {
                    s0 += a.stats[r][offS + cell];
                    s1 += a.stats[r][offS + K3 + cell];
                }
                const double n0 = a.p[rank][2 * cell] * s0, n1 = a.p[rank][2 * cell + 1] * s1;
                double d = TIP_EPS;
                d += n0;
                d += n1;
                const double v0 = n0 / d, v1 = n1 / d;
                for (int r = 0; r < n; ++r) {
                    a.p[r][2 * cell] = v0;
                    a.p[r][2 * cell + 1] = v1;
                }
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        __threadfence_system();                              // this CTA's stores to the peers before anybody is told
        last = (atomicAdd(sync + 1, 1ull) == (unsigned long long)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!last) return;
    if (t < n) {
        st_release_sys(a.flags[t] + n + rank, e);            // my slice has been delivered everywhere
        const long long t0 = clock64();
        while (ld_acquire_sys(a.flags[rank] + n + t) < e) {
            if (clock64() - t0 > 20000000000ll) {
                atomicExch(sync + 2, 1ull);
                break;
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        const bool timed_out = atomicAdd(sync + 2, 0ull) != 0ull;
        sync[1] = 0ull;
        sync[0] = timed_out ? ~0ull : e;
    }
}

// ---------------------------------------------------------------------------------------------
// tip_peer_push_mstep: push-based all-reduce + M-step in one kernel, ONE handshake.
//   1. every CTA copies its part of this rank's statistics into this rank's slot of every peer's INBOX (remote stores);
//      the last CTA to finish tells every peer "my statistics are in your inbox" (release, system scope)
//   2. all CTAs wait until every peer's statistics have arrived in the local inbox
//   3. sum of the n buffers in rank order (own statistics in place of the own slot), M-step of ALL of theta and p, locally
// The data travels BEFORE the barrier - while the slower ranks are still in their E-step - so what is left after the last
// rank arrives is one flag latency and a local sum (the pull formulation starts its remote reads only then).  Inboxes are
// double-buffered by the caller (iteration i uses inbox i & 1): a peer can push iteration i + 1 while this rank still sums
// iteration i, but not iteration i + 2 (that needs this rank's signal of iteration i + 1).  Every rank adds the same
// numbers in the same order: replicas of theta and p stay bit-identical.
struct PushPtrs {
    double *inbox[kMaxPeers];              // inbox of rank q for this parity: [nranks][n_pad]
    unsigned long long *flags[kMaxPeers];  // flag array of rank q: slot r = "rank r's statistics have arrived"
};

__global__ void __launch_bounds__(256) peer_push_mstep_kernel(int P, int K, const double *__restrict__ own, PushPtrs a, int rank,
                                                              int n, int64_t n_pad, int64_t push_from,
                                                              const int32_t *__restrict__ deg, double *__restrict__ theta,
                                                              double *__restrict__ p, unsigned long long *sync)
{
    __shared__ unsigned long long e_sh;
    __shared__ int bad, last;
    const int t = threadIdx.x;
    if (t == 0) {
        e_sh = sync[0] + 1;
        bad = 0;
        last = 0;
    }
    __syncthreads();
    const unsigned long long e = e_sh;
    if (e == 0ull) return;                                   // poisoned by an earlier timeout
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + t, gsz = (int64_t)gridDim.x * blockDim.x;
    // ---- 1: push
    if (push_from > 0) {
        // Ntheta is already in the peers' inboxes (stored by the E-step's finish kernel, complete at the kernel boundary):
        // only the 2 K^3 + 1 doubles behind it are left - CTA 0 sends them and signals, nobody counts CTAs
        if (blockIdx.x == 0) {
            const double2 *src = reinterpret_cast<const double2 *>(own);
            const int64_t n2 = n_pad / 2;
            for (int q = 0; q < n; ++q) {
                if (q == rank) continue;
                double2 *dst = reinterpret_cast<double2 *>(a.inbox[q] + (int64_t)rank * n_pad);
                for (int64_t i = push_from / 2 + t; i < n2; i += blockDim.x) dst[i] = src[i];
            }
            __syncthreads();
            if (t < n) st_release_sys(a.flags[t] + rank, e);   // (release = system-scope fence + store, cumulative over the CTA barrier)
        }
    } else {
        const double2 *src = reinterpret_cast<const double2 *>(own);
        const int64_t n2 = n_pad / 2;
        for (int q = 0; q < n; ++q) {
            if (q == rank) continue;
            double2 *dst = reinterpret_cast<double2 *>(a.inbox[q] + (int64_t)rank * n_pad);
            for (int64_t i = gtid; i < n2; i += gsz) dst[i] = src[i];
        }
        // one system-scope fence per CTA, by the thread that then counts the CTA as done: the CTA barrier makes every
        // thread's stores "observed" by thread 0, whose fence is cumulative (a fence in each of the 16k threads costs
        // microseconds: measured, the first version of this kernel was 8 us slower than the pull formulation at n = 2)
        __syncthreads();
        if (t == 0) {
            __threadfence_system();
            last = (atomicAdd(sync + 1, 1ull) == (unsigned long long)gridDim.x - 1) ? 1 : 0;
        }
        __syncthreads();
        if (last && t < n) st_release_sys(a.flags[t] + rank, e);   // (release = fence + store; the counter ordered the CTAs)
    }
    // ---- 2: wait for everybody's statistics
    if (t < n) {
        const long long t0 = clock64();
        while (ld_acquire_sys(a.flags[rank] + t) < e) {
            if (clock64() - t0 > 20000000000ll) {
                bad = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (bad) {
        if (t == 0) atomicExch(sync + 2, 1ull);
    } else {
        // ---- 3: sum in rank order, M-step
        const int64_t nth = (int64_t)P * K;
        const int K3 = K * K * K;
        const int64_t offS = stats_off_S(P, K);
        const double *in = a.inbox[rank];
        auto sum_at = [&](int64_t i) {
            double s = 0.0;
            for (int r = 0; r < n; ++r) s += (r == rank) ? own[i] : in[(int64_t)r * n_pad + i];
            return s;
        };
        for (int64_t i = gtid; i < nth + K3; i += gsz) {
            if (i < nth) {
                theta[i] = sum_at(i) / (double)deg[i / K];
            } else {
                const int cell = (int)(i - nth);
                const double n0 = p[2 * cell] * sum_at(offS + cell), n1 = p[2 * cell + 1] * sum_at(offS + K3 + cell);
                double d = TIP_EPS;
                d += n0;
                d += n1;
                p[2 * cell] = n0 / d;
                p[2 * cell + 1] = n1 / d;
            }
        }
    }
    __syncthreads();
    if (t == 0 && atomicAdd(sync + 3, 1ull) == (unsigned long long)gridDim.x - 1) {
        const bool timed_out = atomicAdd(sync + 2, 0ull) != 0ull;
        sync[1] = 0ull;
        sync[3] = 0ull;
        sync[0] = timed_out ? ~0ull : e;
    }
}

}  // namespace tip

using namespace tip;

extern "C" int tip_peer_push_mstep(int P, int K, const double *d_own_stats, void *const *h_inbox_ptrs, void *const *h_flag_ptrs,
                                   void *d_sync, int rank, int nranks, int64_t n_pad, int theta_pushed, const int32_t *d_deg,
                                   double *d_theta, double *d_p, void *stream)
{
    TIP_REQUIRE(P > 0 && K >= 1 && K <= TIP_MAX_K && d_own_stats && h_inbox_ptrs && h_flag_ptrs && d_sync && d_deg && d_theta &&
                    d_p && nranks >= 1 && nranks <= kMaxPeers && rank >= 0 && rank < nranks && n_pad >= tip_stats_len(P, K) &&
                    n_pad % 2 == 0,
                "tip_peer_push_mstep: bad arguments (nranks <= %d, n_pad even and >= tip_stats_len)", kMaxPeers);
    PushPtrs a;
    for (int r = 0; r < nranks; ++r) {
        a.inbox[r] = reinterpret_cast<double *>(h_inbox_ptrs[r]);
        a.flags[r] = reinterpret_cast<unsigned long long *>(h_flag_ptrs[r]);
    }
    // every CTA spins on flags: the grid must be resident at once (at most one CTA per SM).  One element of the M-step per
    // thread where the SM count allows: the sum is a chain of L2 round trips per thread (32 CTAs x 8 elements per thread
    // measured 7 us slower than the pull formulation's 592-CTA M-step at n = 2)
    const int64_t total = (int64_t)P * K + (int64_t)K * K * K;
    int64_t want = (total + 255) / 256;
    const int cap = sm_count();
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    peer_push_mstep_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        P, K, d_own_stats, a, rank, nranks, n_pad, theta_pushed ? (int64_t)P * K : 0, d_deg, d_theta, d_p,
        reinterpret_cast<unsigned long long *>(d_sync));
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tip_peer_mstep(int P, int K, void *const *h_stats_ptrs, void *const *h_theta_ptrs, void *const *h_p_ptrs,
                              void *const *h_flag_ptrs, void *d_sync, int rank, int nranks, const int32_t *d_deg, void *stream)
{
    TIP_REQUIRE(P > 0 && K >= 1 && K <= TIP_MAX_K && h_stats_ptrs && h_theta_ptrs && h_p_ptrs && h_flag_ptrs && d_sync && d_deg &&
                    nranks >= 1 && nranks <= kMaxPeers && rank >= 0 && rank < nranks,
                "tip_peer_mstep: bad arguments (nranks <= %d)", kMaxPeers);
    MstepPtrs a;
    for (int r = 0; r < nranks; ++r) {
        a.stats[r] = reinterpret_cast<const double *>(h_stats_ptrs[r]);
        a.theta[r] = reinterpret_cast<double *>(h_theta_ptrs[r]);
        a.p[r] = reinterpret_cast<double *>(h_p_ptrs[r]);
        a.flags[r] = reinterpret_cast<unsigned long long *>(h_flag_ptrs[r]);
    }
    const int64_t total = (int64_t)P * K + (int64_t)K * K * K, per = (total + nranks - 1) / nranks;
    int64_t want = (per + 255) / 256;
    // every CTA spins on flags: the grid must be resident at once (one CTA per SM at most)
    const int cap = sm_count();
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    peer_mstep_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, K, a, rank, nranks, d_deg,
                                                                               reinterpret_cast<unsigned long long *>(d_sync));
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

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

// Link digestion on the device: (g1,g2,g3,n0,n1)[L] -> packed rows ordered (rating, a, b, c), each
// rating block padded to a multiple of 32 rows; plus the distinct-link degree of every gene
// (the `counter` of TIP.py:986-994).
#include <cub/device/device_radix_sort.cuh>

#include "tip_common.cuh"

namespace tip {

constexpr int kGeneBits = 20;  // ids < 2^20; key = r<<60 | a<<40 | b<<20 | c ; invalid = ~0
constexpr unsigned long long kInvalidKey = ~0ull;

__global__ void rows_make_keys_kernel(const int32_t *__restrict__ g1, const int32_t *__restrict__ g2,
                                      const int32_t *__restrict__ g3, const int32_t *__restrict__ n0,
                                      const int32_t *__restrict__ n1, int64_t L, unsigned long long *__restrict__ keys,
                                      int32_t *__restrict__ vals, unsigned long long *__restrict__ counters,
                                      int32_t *__restrict__ deg)
{
    unsigned long long c0 = 0, c1 = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long a = (unsigned)g1[i], b = (unsigned)g2[i], c = (unsigned)g3[i];
        const unsigned long long base = (a << (2 * kGeneBits)) | (b << kGeneBits) | c;
        const int m0 = n0[i], m1 = n1[i];
        keys[2 * i] = m0 > 0 ? base : kInvalidKey;
        vals[2 * i] = m0;
        keys[2 * i + 1] = m1 > 0 ? (base | (1ull << (3 * kGeneBits))) : kInvalidKey;
        vals[2 * i + 1] = m1;
        c0 += m0 > 0;
        c1 += m1 > 0;
        if (deg != nullptr && (m0 > 0 || m1 > 0)) {
            atomicAdd(deg + g1[i], 1);
            atomicAdd(deg + g2[i], 1);
            atomicAdd(deg + g3[i], 1);
        }
    }
    if (c0) atomicAdd(counters + 0, c0);
    if (c1) atomicAdd(counters + 1, c1);
}

__global__ void rows_emit_kernel(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals,
                                 const unsigned long long *__restrict__ counters, int4 *__restrict__ rows)
{
    const int64_t c0 = (int64_t)counters[0], c1 = (int64_t)counters[1];
    const int64_t p0 = (c0 + 31) / 32 * 32, p1 = (c1 + 31) / 32 * 32;
    const unsigned long long mask = (1ull << kGeneBits) - 1;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < p0 + p1; j += (int64_t)gridDim.x * blockDim.x) {
        const int r = j >= p0;
        const int64_t in_block = r ? j - p0 : j;
        const int64_t n_block = r ? c1 : c0;
        int4 out = make_int4(0, 0, 0, r);
        if (in_block < n_block) {
            const int64_t src = r ? c0 + in_block : in_block;
            const unsigned long long key = keys[src];
            out.x = (int)((key >> (2 * kGeneBits)) & mask);
            out.y = (int)((key >> kGeneBits) & mask);
            out.z = (int)(key & mask);
            out.w = (vals[src] << 1) | r;
        }
        rows[j] = out;
    }
}

struct PackWs {
    unsigned long long *keys_in, *keys_out, *counters;
    int32_t *vals_in, *vals_out;
    void *cub_tmp;
    size_t cub_bytes;
};

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

static size_t pack_layout(int64_t L, void *base, PackWs *ws)
{
    const size_t n = (size_t)(2 * L);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int64_t)n, 0, 64, (cudaStream_t)0);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align256(bytes);
        return o;
    };
    const size_t o_ki = take(n * 8), o_ko = take(n * 8), o_vi = take(n * 4), o_vo = take(n * 4), o_cnt = take(64),
                 o_cub = take(cub_bytes);
    if (ws) {
        char *b = reinterpret_cast<char *>(base);
        ws->keys_in = reinterpret_cast<unsigned long long *>(b + o_ki);
        ws->keys_out = reinterpret_cast<unsigned long long *>(b + o_ko);
        ws->vals_in = reinterpret_cast<int32_t *>(b + o_vi);
        ws->vals_out = reinterpret_cast<int32_t *>(b + o_vo);
        ws->counters = reinterpret_cast<unsigned long long *>(b + o_cnt);
        ws->cub_tmp = b + o_cub;
        ws->cub_bytes = cub_bytes;
    }
    return off + 256;
}

}  // namespace tip

using namespace tip;

extern "C" int64_t tip_rows_capacity(int64_t L) { return 2 * L + 64; }

extern "C" int tip_pack_rows_workspace_bytes(int64_t L, size_t *bytes)
{
    TIP_REQUIRE(L >= 0 && bytes != nullptr, "tip_pack_rows_workspace_bytes: bad arguments");
    *bytes = pack_layout(L < 1 ? 1 : L, nullptr, nullptr);
    return 0;
}

extern "C" int tip_pack_rows(const int32_t *d_g1, const int32_t *d_g2, const int32_t *d_g3, const int32_t *d_n0,
                             const int32_t *d_n1, int64_t L, int P, void *d_ws, size_t ws_bytes, void *d_rows,
                             int64_t *h_n_rows, int64_t *h_part, int32_t *d_deg, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TIP_REQUIRE(L >= 0 && P > 0 && P <= (1 << kGeneBits), "tip_pack_rows: need 0 < P <= 2^20 (got P=%d)", P);
    TIP_REQUIRE(h_n_rows != nullptr && h_part != nullptr, "tip_pack_rows: null output pointer");
    if (d_deg) TIP_CHECK_CUDA(cudaMemsetAsync(d_deg, 0, sizeof(int32_t) * (size_t)P, st));
    if (L == 0) {
        *h_n_rows = 0;
        h_part[0] = h_part[1] = h_part[2] = 0;
        return 0;
    }
    size_t need = pack_layout(L, nullptr, nullptr);
    TIP_REQUIRE(d_ws != nullptr && ws_bytes >= need, "tip_pack_rows: workspace too small (%zu < %zu)", ws_bytes, need);
    uintptr_t basep = (reinterpret_cast<uintptr_t>(d_ws) + 255) / 256 * 256;
    PackWs ws;
    pack_layout(L, reinterpret_cast<void *>(basep), &ws);
    TIP_CHECK_CUDA(cudaMemsetAsync(ws.counters, 0, 64, st));
    const int threads = 256;
    int64_t want = (L + threads - 1) / threads;
    int grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
    rows_make_keys_kernel<<<grid, threads, 0, st>>>(d_g1, d_g2, d_g3, d_n0, d_n1, L, ws.keys_in, ws.vals_in,
                                                    ws.counters, d_deg);
    TIP_CHECK_CUDA(cudaGetLastError());
    size_t cub_bytes = ws.cub_bytes;
    TIP_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, cub_bytes, ws.keys_in, ws.keys_out, ws.vals_in, ws.vals_out,
                                                   (int64_t)(2 * L), 0, 64, st));
    unsigned long long h_cnt[2] = {0, 0};
    TIP_CHECK_CUDA(cudaMemcpyAsync(h_cnt, ws.counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    TIP_CHECK_CUDA(cudaStreamSynchronize(st));
    const int64_t p0 = ((int64_t)h_cnt[0] + 31) / 32 * 32, p1 = ((int64_t)h_cnt[1] + 31) / 32 * 32;
    if (p0 + p1 > 0) {
        want = (p0 + p1 + threads - 1) / threads;
        grid = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
        rows_emit_kernel<<<grid, threads, 0, st>>>(ws.keys_out, ws.vals_out, ws.counters, reinterpret_cast<int4 *>(d_rows));
        TIP_CHECK_CUDA(cudaGetLastError());
        TIP_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    *h_n_rows = p0 + p1;
    h_part[0] = p0;
    h_part[1] = p1;
    h_part[2] = (int64_t)(h_cnt[0] + h_cnt[1]);
    return 0;
}

// K-specialised fused E-step for sm_100a (K = 1..10): the hot kernel of Model.make_iteration
// (TIP.py:987-1012).  fp64 CUDA-core FMA bound: 3 DFMA per (a,b,c) cell per link-update.
//
// One persistent CTA per SM, 8 warps, every warp independent (no CTA barrier in the main loop).
// A warp walks tiles of 32 packed rows (all one rating, see tip.h):
//
//   gather   cp.async the three theta rows of each of the 32 links into a [link][th_a|th_b|th_c]
//            shared-memory stage, row-contiguous so every request touches whole 8*K-byte rows;
//            double buffered: the rows of tile t+1 arrive while tile t is computed.
//   phase A  lane = link.  th_b, th_c and the accumulators v[K], w[K] live in registers;
//            p[.][.][.][r] is read from shared memory as warp-uniform (broadcast) 128-bit loads.
//              q[ab]  = sum_c p[abc] th_c[c]              K^3 DFMA
//              w[c]  += th_a[a] th_b[b] p[abc]            K^3 DFMA
//              u[a]  += th_b[b] q[ab],  v[b] += th_a[a] q[ab]
//            d = eps + sum_a th_a[a] u[a];  s = count / d.
//            Contributions s*th_a*u, s*th_b*v, s*th_c*w go to a [link][3K] shared buffer, s*th_c
//            overwrites th_c in the stage.
//   scatter  slot-a contributions are summed over runs of equal gene (rows are sorted by slot-a
//            gene) before one red.global.add.f64 per run; slot-b/c contributions are issued
//            row-contiguously (one warp instruction covers 3.2 theta rows).
//   phase B  lanes = cells.  S[a][b][c] += th_a[a] th_b[b] * (s th_c[c]) for the 32 links of the
//            tile, each lane owning a (1 x BG x K) block of S in registers for the whole kernel.
//            (K <= 4: S is small enough to be thread-private and is updated in phase A.)
//
// At the end S is reduced across the CTA in shared memory and added to the global statistics.
#include "tip_common.cuh"

namespace tip {

constexpr int kEmWarps = 8;
constexpr int kEmThreads = kEmWarps * kWarp;

template <int K>
struct EmCfg {
    static constexpr int KP = K + (K & 1);                                // sub-row length (even)
    static constexpr int S3 = 3 * KP;                                     //
    static constexpr int RS = S3 + ((S3 % 4 == 2) ? 0 : 2);               // row stride == 2 (mod 4): conflict-free 128-bit rows
    static constexpr int K3 = K * K * K;
    static constexpr bool kPrivateS = (K <= 4);                           // S thread-private
    // phase B lane mapping: lane -> (alpha = lane % K, group = lane / K), group covers BG betas
    static constexpr int NG = (32 / K) < K ? (32 / K) : K;
    static constexpr int BG = (K + NG - 1) / NG;
    static constexpr int NGU = (K + BG - 1) / BG;                         // groups actually needed
    static constexpr int SP = 2 * K * K * KP;                             // doubles of staged p (both ratings)
    // per-warp shared memory (doubles): 2 theta stages + contribution buffer, then ids (2 x 32 int4)
    static constexpr int WARP_DBL = 3 * 32 * RS;
    static constexpr size_t WARP_BYTES = (size_t)WARP_DBL * 8 + 2 * 32 * 16;
    static constexpr size_t SMEM = (size_t)SP * 8 + (size_t)2 * K3 * 8 + kEmWarps * WARP_BYTES;
};

template <int K>
__device__ __forceinline__ void gather_tile(const double *__restrict__ theta, const int4 *ids, double *stage, int lane)
{
    using C = EmCfg<K>;
    const int *idw = reinterpret_cast<const int *>(ids);
#pragma unroll 5
    for (int i = 0; i < 3 * K; ++i) {
        const int idx = i * 32 + lane;
        const int l = idx / (3 * K), rem = idx - l * (3 * K);
        const int slot = rem / K, k = rem - slot * K;
        const int g = idw[l * 4 + slot];
        cp_async_8(stage + l * C::RS + slot * C::KP + k, theta + (int64_t)g * K + k);
    }
}

template <int K>
__global__ void __launch_bounds__(kEmThreads, 1)
    em_fused_kernel(int P, const int4 *__restrict__ rows, int64_t n_tiles, const double *__restrict__ theta,
                    const double *__restrict__ p, double *__restrict__ stats)
{
    using C = EmCfg<K>;
    constexpr int KP = C::KP, RS = C::RS, K3 = C::K3, BG = C::BG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sp = reinterpret_cast<double *>(smem_raw);            // [2][K*K][KP]
    double *Ssm = sp + C::SP;                                     // [2][K3]
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    unsigned char *wbase = reinterpret_cast<unsigned char *>(Ssm + 2 * K3) + (size_t)warp * C::WARP_BYTES;
    double *stage0 = reinterpret_cast<double *>(wbase);           // [2][32][RS]
    double *cbuf = stage0 + 2 * 32 * RS;                          // [32][RS]
    int4 *ids_sm = reinterpret_cast<int4 *>(cbuf + 32 * RS);      // [2][32]

    // ---- stage p (transposed to [r][ab][c], c padded to KP) and clear the CTA's S ----
    for (int e = threadIdx.x; e < 2 * K * K * KP; e += kEmThreads) {
        const int r = e / (K * K * KP), rem = e - r * (K * K * KP);
        const int pair = rem / KP, c = rem - pair * KP;
        sp[e] = (c < K) ? __ldg(p + ((int64_t)pair * K + c) * 2 + r) : 0.0;
    }
    for (int e = threadIdx.x; e < 2 * K3; e += kEmThreads) Ssm[e] = 0.0;
    __syncthreads();

    const int64_t W = (int64_t)gridDim.x * kEmWarps;
    const int64_t w0 = (int64_t)blockIdx.x * kEmWarps + warp;

    // phase B lane mapping
    const int al_b = lane % K;
    int grp = lane / K;
    const bool b_active = (!C::kPrivateS) && grp < C::NGU;
    if (grp >= C::NGU) grp = 0;
    const int be0 = grp * BG;

    constexpr int NS = C::kPrivateS ? K3 : BG * K;
    double Sacc[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) Sacc[i] = 0.0;
    int cur_r = -1;
    double ll = 0.0;

    auto flush_S = [&](int r) {
        if (r < 0) return;
        double *dst = Ssm + r * K3;
        if constexpr (C::kPrivateS) {
#pragma unroll
            for (int i = 0; i < K3; ++i) {
                const double t = warp_sum(Sacc[i]);
                if (lane == 0) atomicAdd(dst + i, t);
                Sacc[i] = 0.0;
            }
        } else {
#pragma unroll
            for (int j = 0; j < BG; ++j)
#pragma unroll
                for (int c = 0; c < K; ++c) {
                    if (b_active && be0 + j < K) atomicAdd(dst + (al_b * K + be0 + j) * K + c, Sacc[j * K + c]);
                    Sacc[j * K + c] = 0.0;
                }
        }
    };

    // ---- software pipeline prologue ----
    int4 ids_next = make_int4(0, 0, 0, 0);
    if (w0 < n_tiles) {
        ids_sm[lane] = rows[w0 * 32 + lane];
        __syncwarp();
        gather_tile<K>(theta, ids_sm, stage0, lane);
        cp_async_commit();
        if (w0 + W < n_tiles) ids_next = rows[(w0 + W) * 32 + lane];
    }

    int buf = 0;
    for (int64_t t = w0; t < n_tiles; t += W, buf ^= 1) {
        double *stage = stage0 + buf * 32 * RS;
        const int4 *ids = ids_sm + buf * 32;
        const bool has_next = (t + W) < n_tiles;
        if (has_next) {
            ids_sm[(buf ^ 1) * 32 + lane] = ids_next;
            __syncwarp();
            gather_tile<K>(theta, ids_sm + (buf ^ 1) * 32, stage0 + (buf ^ 1) * 32 * RS, lane);
            cp_async_commit();
            if (t + 2 * W < n_tiles) ids_next = rows[(t + 2 * W) * 32 + lane];
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();

        const int4 me = ids[lane];
        const int r = row_rating(me.w);  // uniform across the tile
        const double cnt = (double)row_count(me.w);
        if (r != cur_r) {
            flush_S(cur_r);
            cur_r = r;
        }

        // ================= phase A: lane = link =================
        {
            double *row = stage + lane * RS;
            double *crow = cbuf + lane * RS;
            const double *spr = sp + r * (K * K * KP);
            double tb[KP], tc[KP], v[K], w[KP];
#pragma unroll
            for (int k = 0; k < KP; k += 2) {
                const double2 b2 = *reinterpret_cast<const double2 *>(row + KP + k);
                const double2 c2 = *reinterpret_cast<const double2 *>(row + 2 * KP + k);
                tb[k] = b2.x; tb[k + 1] = b2.y;
                tc[k] = c2.x; tc[k + 1] = c2.y;
                w[k] = 0.0; w[k + 1] = 0.0;
            }
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = 0.0;
            double dsum = 0.0;
            double tprev = 0.0;
#pragma unroll
            for (int a = 0; a < K; ++a) {
                const double ta = row[a];
                double u = 0.0;
#pragma unroll
                for (int b = 0; b < K; ++b) {
                    const double2 *pp = reinterpret_cast<const double2 *>(spr + (a * K + b) * KP);
                    const double ab = ta * tb[b];
                    double q0 = 0.0, q1 = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 < KP / 2; ++c2) {
                        const double2 pv = pp[c2];
                        q0 = fma(pv.x, tc[2 * c2], q0);
                        w[2 * c2] = fma(ab, pv.x, w[2 * c2]);
                        if (2 * c2 + 1 < K) {
                            q1 = fma(pv.y, tc[2 * c2 + 1], q1);
                            w[2 * c2 + 1] = fma(ab, pv.y, w[2 * c2 + 1]);
                        }
                    }
                    const double q = q0 + q1;
                    u = fma(tb[b], q, u);
                    v[b] = fma(ta, q, v[b]);
                }
                const double tt = ta * u;
                dsum += tt;
                if (a & 1) {
                    *reinterpret_cast<double2 *>(crow + a - 1) = make_double2(tprev, tt);
                } else if (a == K - 1) {
                    crow[a] = tt;
                }
                tprev = tt;
            }
            const double d = TIP_EPS + dsum;
            const double s = cnt / d;
            ll += cnt * log(d);
            // contributions: slot a (scale what is already there), slot b, slot c; s*th_c into the stage
#pragma unroll
            for (int k = 0; k < KP; k += 2) {
                double2 ca = *reinterpret_cast<double2 *>(crow + k);
                ca.x *= s; ca.y *= s;
                *reinterpret_cast<double2 *>(crow + k) = ca;
                const double vb1 = (k + 1 < K) ? v[k + 1] : 0.0;
                *reinterpret_cast<double2 *>(crow + KP + k) = make_double2(s * tb[k] * v[k], s * tb[k + 1] * vb1);
                const double sc0 = s * tc[k], sc1 = s * tc[k + 1];
                *reinterpret_cast<double2 *>(crow + 2 * KP + k) = make_double2(sc0 * w[k], sc1 * w[k + 1]);
                *reinterpret_cast<double2 *>(row + 2 * KP + k) = make_double2(sc0, sc1);
            }
            if constexpr (C::kPrivateS) {
                // thread-private S += th_a th_b (s th_c)
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const double ta = row[a];
#pragma unroll
                    for (int b = 0; b < K; ++b) {
                        const double sab = s * ta * tb[b];
#pragma unroll
                        for (int c = 0; c < K; ++c) Sacc[(a * K + b) * K + c] = fma(sab, tc[c], Sacc[(a * K + b) * K + c]);
                    }
                }
            }
        }
        __syncwarp();

        // ================= scatter theta statistics =================
        {
            const int *idw = reinterpret_cast<const int *>(ids);
            // slot a: run-length pre-reduction.  lane -> (k = lane % K, part = lane / K); each part
            // walks a contiguous range of the 32 links and emits one reduction per run of equal gene.
            constexpr int NPART = 32 / K > 4 ? 4 : (32 / K);  // K<=8 -> 4 parts, 9,10 -> 3 parts
            constexpr int LPP = (32 + NPART - 1) / NPART;
            const int k = lane % K, part = lane / K;
            if (part < NPART) {
                const int l0 = part * LPP;
                const int l1 = (l0 + LPP < 32) ? l0 + LPP : 32;
                double acc = 0.0;
                int g = idw[l0 * 4];
                for (int l = l0; l < l1; ++l) {
                    const int gl = idw[l * 4];
                    if (gl != g) {
                        if (acc != 0.0) red_add_f64(stats + (int64_t)g * K + k, acc);
                        acc = 0.0;
                        g = gl;
                    }
                    acc += cbuf[l * RS + k];
                }
                if (acc != 0.0) red_add_f64(stats + (int64_t)g * K + k, acc);
            }
            // slots b, c: row-contiguous reductions
#pragma unroll 4
            for (int i = 0; i < 2 * K; ++i) {
                const int idx = i * 32 + lane;
                const int l = idx / (2 * K), rem = idx - l * (2 * K);
                const int slot = 1 + rem / K, kk = rem % K;
                const double val = cbuf[l * RS + slot * KP + kk];
                if (val != 0.0) red_add_f64(stats + (int64_t)idw[l * 4 + slot] * K + kk, val);
            }
        }

        // ================= phase B: lanes = cells of S =================
        if constexpr (!C::kPrivateS) {
#pragma unroll 4
            for (int l = 0; l < 32; ++l) {
                const double *rw = stage + l * RS;
                const double ta = rw[al_b];
                double tbv[BG];
                if constexpr (BG % 2 == 0) {
#pragma unroll
                    for (int j = 0; j < BG; j += 2) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(rw + KP + be0 + j);
                        tbv[j] = t2.x; tbv[j + 1] = t2.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < BG; ++j) tbv[j] = rw[KP + be0 + j];
                }
                double sc[KP];
#pragma unroll
                for (int c = 0; c < KP; c += 2) {
                    const double2 t2 = *reinterpret_cast<const double2 *>(rw + 2 * KP + c);
                    sc[c] = t2.x; sc[c + 1] = t2.y;
                }
#pragma unroll
                for (int j = 0; j < BG; ++j) {
                    const double ab = ta * tbv[j];
#pragma unroll
                    for (int c = 0; c < K; ++c) Sacc[j * K + c] = fma(ab, sc[c], Sacc[j * K + c]);
                }
            }
        }
        __syncwarp();
    }
    flush_S(cur_r);
    ll = warp_sum(ll);
    if (lane == 0 && ll != 0.0) red_add_f64(stats + stats_off_ll(P, K), ll);
    __syncthreads();
    double *Sg = stats + stats_off_S(P, K);
    for (int e = threadIdx.x; e < 2 * K3; e += kEmThreads) {
        const double v = Ssm[e];
        if (v != 0.0) red_add_f64(Sg + e, v);
    }
}

template <int K>
static int launch_em_fused(int P, const int4 *rows, int64_t n_rows, const double *theta, const double *p,
                           double *stats, cudaStream_t st)
{
    using C = EmCfg<K>;
    static bool attr_done = false;
    if (!attr_done) {
        TIP_CHECK_CUDA(cudaFuncSetAttribute(em_fused_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_done = true;
    }
    const int64_t n_tiles = n_rows / 32;
    int64_t want = (n_tiles + kEmWarps - 1) / kEmWarps;
    int grid = (int)(want < sm_count() ? want : sm_count());
    if (grid < 1) grid = 1;
    em_fused_kernel<K><<<grid, kEmThreads, C::SMEM, st>>>(P, rows, n_tiles, theta, p, stats);
    TIP_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_em_tuned(int P, int K, const int4 *rows, int64_t n_rows, const double *theta, const double *p,
                    double *stats, cudaStream_t st, bool *handled)
{
    *handled = true;
    switch (K) {
        case 1: return launch_em_fused<1>(P, rows, n_rows, theta, p, stats, st);
        case 2: return launch_em_fused<2>(P, rows, n_rows, theta, p, stats, st);
        case 3: return launch_em_fused<3>(P, rows, n_rows, theta, p, stats, st);
        case 4: return launch_em_fused<4>(P, rows, n_rows, theta, p, stats, st);
        case 5: return launch_em_fused<5>(P, rows, n_rows, theta, p, stats, st);
        case 6: return launch_em_fused<6>(P, rows, n_rows, theta, p, stats, st);
        case 7: return launch_em_fused<7>(P, rows, n_rows, theta, p, stats, st);
        case 8: return launch_em_fused<8>(P, rows, n_rows, theta, p, stats, st);
        case 9: return launch_em_fused<9>(P, rows, n_rows, theta, p, stats, st);
        case 10: return launch_em_fused<10>(P, rows, n_rows, theta, p, stats, st);
        default: *handled = false; return 0;
    }
}

}  // namespace tip

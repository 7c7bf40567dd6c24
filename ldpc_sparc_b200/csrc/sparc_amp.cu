// SPARC AMP decoder and Ab/Az operators for the sub-sampled DCT design, sm_100a.
//
// Reference: sparc_public/sparc.py -- sparc_amp :883-999, sparc_transforms
// :703-880, msg_vector_mmse_estimator :402-465 (K = 1), msg_vector_map_estimator
// :467-512 (K = 1).  One CTA decodes one frame at a time out of a device-side
// frame queue (frames stop early, at different iterations).  Per frame the
// residual z lives in shared memory next to the FFT buffer; beta / s stream
// through a per-CTA global scratch (L2 resident).  Per AMP iteration (t >= 1
// for the first half):
//   z_r   = y_r + (gamma_r / phi_prev_r) z_r - sum_c A_rc beta_c
//           (one DCT-II per non-zero block, subtracted from z in place)
//   phi_r = awgn_var + gamma_r   |   mean(z_r^2)                    (phi_est_method)
//   tau_c = (L phi / n) / W_c    |   (L / Mr) / sum_r W_rc / phi_r
//   s_c   = beta_c + tau_c sum_r A_rc^T (z_r / phi_r)   (one DCT-III per block)
//   beta  = section-wise softmax(s / tau)  (warp shuffles, per-section max shift)
//   psi_c = 1 - sum(beta_c^2) / (L/Lc), evaluated as sum beta (1 - beta) in fp64
//   stop when |psi - psi_prev| <= atol + rtol |psi_prev| for every c      (:984)
// and finally the MAP decision argmax_section(s) (:997).
#include "common.cuh"
#include "sparc_dct.cuh"
#include "sparc_pair.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

using namespace lsb_sparc;

// resident CTAs per SM the fp32 AMP kernel's register budget is tuned for
#ifndef LSB_AMP_OCC
#define LSB_AMP_OCC 5
#endif

namespace {

constexpr int kThreads = 256;
constexpr int kMaxLr = 64, kMaxLc = 64;
constexpr int kCounters = 64;

struct DesignDev {
    int L, M, n, Lr, Lc, Mr, Mc, w, h, n_blocks, logM;
    const BlockDev *blocks;        // sorted by (c, r)
    const int *col_first;          // [Lc + 1] range of blocks of column c
    const double *W;               // [Lr * Lc]
};

struct AmpParams {
    double awgn_var, rtol, atol;
    int t_max, phi_method;
    int K;   // 1: unmodulated, 2: real 2-PSK values
};

struct AmpOut {
    int32_t *idx;        // [F, L] MAP section decisions
    int32_t *t_final;    // [F]
    double *psi;         // [F, Lc] final psi
    float *nmse;         // [F, t_max, Lc] or null (needs idx0)
    void *beta_soft;     // [F, LM] T or null: beta of the last iteration
    void *s_last;        // [F, LM] T or null: s of the last iteration
    double *tau_trace;   // [F, t_max, Lc] or null
    double *psi_trace;   // [F, t_max, Lc] or null
};

template <typename T> __device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Gather / scatter loops: the U index-table entries of a batch are loaded before any
// dependent shared-memory access, so U global loads are in flight per thread instead of
// one (ncu: the table load -> address -> shared access chain was the top stall site).
constexpr int kBatch = 4;

// byte offset of the staged-FFT tile behind the n residual entries (16-byte aligned)
__host__ __device__ inline size_t staged_tile_offset(int n, size_t elem) { return (((size_t)n * elem) + 15) & ~(size_t)15; }

template <typename T>
__device__ __forceinline__ void gather_add(T *dst, const T *bf, const uint16_t *__restrict__ tab, T sc, int n, int tid)
{
    for (int i0 = tid; i0 < n; i0 += kThreads * kBatch) {
        int id[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kThreads;
            id[u] = i < n ? (int)__ldg(tab + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kThreads;
            if (i < n) dst[i] += sc * bf[id[u]];
        }
    }
}

template <typename T>
__device__ __forceinline__ void scatter_scaled(T *bf, const uint16_t *__restrict__ tab, const T *src, T scale, int n,
                                               int tid)
{
    for (int i0 = tid; i0 < n; i0 += kThreads * kBatch) {
        int id[kBatch];
        T v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kThreads;
            id[u] = i < n ? (int)__ldg(tab + i) : -1;
            v[u] = i < n ? src[i] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u)
            if (id[u] >= 0) bf[id[u]] = v[u] * scale;
    }
}

// bf[0 .. count) = 0 with 16-byte stores (count is a multiple of 16 / sizeof(T): the buffer holds 2H >= 4 entries)
template <typename T>
__device__ __forceinline__ void zero_fill(T *bf, int count, int tid)
{
    constexpr int V = 16 / sizeof(T);
    uint4 *p = reinterpret_cast<uint4 *>(bf);
    const int n16 = count / V;
    for (int i = tid; i < n16; i += kThreads) p[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = n16 * V + tid; i < count; i += kThreads) bf[i] = (T)0;
}

// ---- A_rc x : scatter a column block into the FFT buffer, DCT-II, leave Y in buf
template <typename T, int LOGH = 0>
__device__ void block_forward(typename Cplx<T>::type *buf, const FftTables<T> &tb, const BlockDev &b,
                              const T *x /* Mc entries of the column block, or null */, int one_hot_pos,
                              int Mc, int H, int tid, typename Cplx<T>::type *tile = nullptr)
{
    T *bf = reinterpret_cast<T *>(buf);
    zero_fill(bf, 2 * H, tid);
    __syncthreads();
    if (x) {
        scatter_scaled<T>(bf, b.ab_scatter, x, (T)1, Mc, tid);
    }
    __syncthreads();
    if constexpr (LOGH > 0) fft_fixed<T, LOGH, false, kThreads>(buf, tb, tid);
    else if (tile) fft_staged_h15<T, false>(buf, tile, tb, tid, kThreads);      // buffer in global memory, H = 2^15
    else fft_inplace<T, false>(buf, tb, tid, kThreads);
    dct2_post<T>(buf, tb, tid, kThreads);
}

// ---- A_rc^T y : scatter a row block into the FFT buffer, DCT-III, leave X in buf
template <typename T, int LOGH = 0>
__device__ void block_transpose(typename Cplx<T>::type *buf, const FftTables<T> &tb, const BlockDev &b,
                                const T *y /* Mr entries of the row block */, T yscale, int Mr, int H, int tid,
                                typename Cplx<T>::type *tile = nullptr)
{
    T *bf = reinterpret_cast<T *>(buf);
    zero_fill(bf, 2 * H, tid);
    __syncthreads();
    scatter_scaled<T>(bf, b.az_scatter, y, yscale, Mr, tid);
    __syncthreads();
    dct3_pre<T>(buf, tb, tid, kThreads);
    if constexpr (LOGH > 0) fft_fixed<T, LOGH, true, kThreads>(buf, tb, tid);
    else if (tile) fft_staged_h15<T, true>(buf, tile, tb, tid, kThreads);
    else fft_inplace<T, true>(buf, tb, tid, kThreads);
}

template <typename T> struct Scale;
template <> struct Scale<float> { static __device__ __forceinline__ float get(const BlockDev &b) { return b.scale; } };
template <> struct Scale<double> { static __device__ __forceinline__ double get(const BlockDev &b) { return b.scale_d; } };

// ------------------------------------------------------------ operator kernels
// out[f, n] = A x[f, LM]   (reference Ab, sparc.py:786-789,811-819,851-862)
// GBUF: the FFT buffer does not fit shared memory (w > 32768, e.g. BASELINE config 4
// with w = 65536) and lives in a per-CTA global scratch (L2 resident) instead.
template <typename T, bool GBUF>
__global__ void __launch_bounds__(kThreads)
ab_kernel(DesignDev d, FftTables<T> tb, const T *__restrict__ x, T *__restrict__ out, long long n_frames, T *gbuf)
{
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename Cplx<T>::type T2;
    const int H = 1 << d.h;
    T2 *buf = GBUF ? reinterpret_cast<T2 *>(gbuf + (size_t)blockIdx.x * d.w) : reinterpret_cast<T2 *>(smem);
    T *acc = GBUF ? reinterpret_cast<T *>(smem) : reinterpret_cast<T *>(buf + H);   // [n]
    const int tid = threadIdx.x;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        for (int i = tid; i < d.n; i += kThreads) acc[i] = (T)0;
        for (int bi = 0; bi < d.n_blocks; ++bi) {
            const BlockDev b = d.blocks[bi];
            block_forward<T>(buf, tb, b, x + f * (long long)d.L * d.M + (long long)b.c * d.Mc, -1, d.Mc, H, tid);
            const T *bf = reinterpret_cast<const T *>(buf);
            const T sc = Scale<T>::get(b);
            gather_add<T>(acc + b.r * d.Mr, bf, b.ab_gather, sc, d.Mr, tid);
            __syncthreads();
        }
        for (int i = tid; i < d.n; i += kThreads) out[f * d.n + i] = acc[i];
        __syncthreads();
    }
}

// out[f, LM] = A^T y[f, n]   (reference Az, sparc.py:791-794,821-829,864-875)
template <typename T, bool GBUF>
__global__ void __launch_bounds__(kThreads)
az_kernel(DesignDev d, FftTables<T> tb, const T *__restrict__ y, T *__restrict__ out, long long n_frames, T *gbuf)
{
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename Cplx<T>::type T2;
    const int H = 1 << d.h;
    T2 *buf = GBUF ? reinterpret_cast<T2 *>(gbuf + (size_t)blockIdx.x * d.w) : reinterpret_cast<T2 *>(smem);
    const int tid = threadIdx.x;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        T *o = out + f * (long long)d.L * d.M;
        for (int c = 0; c < d.Lc; ++c) {
            for (int bi = d.col_first[c]; bi < d.col_first[c + 1]; ++bi) {
                const BlockDev b = d.blocks[bi];
                block_transpose<T>(buf, tb, b, y + f * d.n + b.r * d.Mr, (T)1, d.Mr, H, tid);
                const T *bf = reinterpret_cast<const T *>(buf);
                const T sc = Scale<T>::get(b);
                const bool first = bi == d.col_first[c];
                for (int j = tid; j < d.Mc; j += kThreads) {
                    const T v = sc * bf[b.az_gather[j]];
                    T *dst = o + (long long)c * d.Mc + j;
                    *dst = first ? v : *dst + v;
                }
                __syncthreads();
            }
            if (d.col_first[c] == d.col_first[c + 1])
                for (int j = tid; j < d.Mc; j += kThreads) o[(long long)c * d.Mc + j] = (T)0;
        }
    }
}

// ------------------------------------------------------------------ AMP kernel
// LOGH > 0: log2 of the FFT size fixed at compile time (fft_fixed), 0: any size.
template <typename T, bool GBUF, int LOGH = 0>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? LSB_AMP_OCC : 2)
amp_kernel(DesignDev d, FftTables<T> tb, AmpParams p, const T *__restrict__ y_all,
           const int32_t *__restrict__ idx0, AmpOut o, T *__restrict__ scratch, long long n_frames,
           unsigned long long *__restrict__ queue)
{
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename Cplx<T>::type T2;
    const int H = 1 << d.h;
    const int LM = d.L * d.M;
    // per-CTA global scratch: [beta LM][s LM] (+ [FFT buffer w] when it does not fit shared memory)
    T *cta_scratch = scratch + (size_t)blockIdx.x * (2 * (size_t)LM + (GBUF ? (size_t)d.w : 0));
    T2 *buf = GBUF ? reinterpret_cast<T2 *>(cta_scratch + 2 * (size_t)LM) : reinterpret_cast<T2 *>(smem);
    T *z = GBUF ? reinterpret_cast<T *>(smem) : reinterpret_cast<T *>(buf + H);     // [n]
    // GBUF, H = 2^15: a 4096-point tile behind z stages the FFT passes through shared memory
    T2 *tile = (GBUF && d.h == 15) ? reinterpret_cast<T2 *>(smem + staged_tile_offset(d.n, sizeof(T))) : nullptr;
    __shared__ double psi[kMaxLc], psi_prev[kMaxLc], tau[kMaxLc], nm[kMaxLc];
    __shared__ double phi[kMaxLr], phi_prev[kMaxLr], gam[kMaxLr], zsq[kMaxLr];
    // per-warp partial sums, added in fixed warp order by one thread: results are bit-reproducible run to run
    // (slots of column parity c & 1 are rewritten two columns, i.e. several barriers, after they were read)
    __shared__ double part_psi[2][kThreads / 32], part_nm[2][kThreads / 32], part_z[kThreads / 32];
    __shared__ long long s_frame;
    __shared__ int s_stop;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T *beta = cta_scratch;
    T *s = beta + LM;
    const int secs_per_col = d.Mc / d.M;                       // sections per column block
    const double sec_norm = (double)d.L / d.Lc;
    const bool fuse_denoiser = p.K == 1 && d.M <= 32;

    for (;;) {
        if (tid == 0) s_frame = (long long)atomicAdd(queue, 1ull);
        __syncthreads();
        const long long f = s_frame;
        if (f >= n_frames) break;
        const T *y = y_all + f * d.n;

        for (int i = tid; i < d.n; i += kThreads) z[i] = y[i];            // z = y   (:915)
        if (tid < d.Lc) psi[tid] = 1.0;                                   // gamma = W.1/Lc  (:918,926)
        if (tid < kMaxLr) { phi[tid] = 0.0; }
        __syncthreads();

        int t = 0;
        for (; t < p.t_max - 1; ++t) {
            // ---- gamma_r (and, for t > 0, the Onsager coefficient and the new residual)
            if (tid < d.Lr) {
                double g = 0.0;
                for (int c = 0; c < d.Lc; ++c) g += d.W[tid * d.Lc + c] * psi[c];
                gam[tid] = g / d.Lc;
                phi_prev[tid] = phi[tid];
                zsq[tid] = 0.0;
            }
            if (tid < d.Lc) psi_prev[tid] = psi[tid];
            __syncthreads();
            if (t > 0) {
                // z = y - A beta + b z (:944-946), accumulated in place: z <- y + b z, then every
                // non-zero block subtracts its A_rc beta_c (no separate A*beta vector)
                for (int i = tid; i < d.n; i += kThreads) {
                    const int r = i / d.Mr;
                    const T coef = (T)(gam[r] / phi_prev[r]);
                    z[i] = y[i] + coef * z[i];
                }
                for (int bi = 0; bi < d.n_blocks; ++bi) {
                    const BlockDev b = d.blocks[bi];
                    block_forward<T, LOGH>(buf, tb, b, beta + (size_t)b.c * d.Mc, -1, d.Mc, H, tid, tile);
                    const T *bf = reinterpret_cast<const T *>(buf);
                    gather_add<T>(z + b.r * d.Mr, bf, b.ab_gather, -Scale<T>::get(b), d.Mr, tid);
                    __syncthreads();
                }
            }
            // ---- phi (:949-955)
            if (p.phi_method == 2) {
                for (int r = 0; r < d.Lr; ++r) {
                    double part = 0.0;
                    for (int i = tid; i < d.Mr; i += kThreads) { const double v = (double)z[r * d.Mr + i]; part += v * v; }
                    part = warp_sum(part);
                    if (lane == 0) part_z[warp] = part;
                    __syncthreads();
                    if (tid == 0) {
                        double a = 0.0;
                        for (int w = 0; w < kThreads / 32; ++w) a += part_z[w];
                        zsq[r] = a;
                    }
                    __syncthreads();
                }
            }
            if (tid < d.Lr) phi[tid] = p.phi_method == 1 ? p.awgn_var + gam[tid] : zsq[tid] / d.Mr;
            __syncthreads();
            // ---- tau (:958-969)
            if (tid < d.Lc) {
                double tc;
                if (d.Lr == 1) {
                    tc = ((double)d.L * phi[0] / d.n) / d.W[tid];
                } else {
                    double acc_w = 0.0;
                    for (int r = 0; r < d.Lr; ++r) acc_w += d.W[r * d.Lc + tid] / phi[r];
                    tc = ((double)d.L / d.Mr) / acc_w;
                }
                tau[tid] = tc;
                psi[tid] = 0.0;
                nm[tid] = 0.0;
                if (o.tau_trace) o.tau_trace[(f * p.t_max + t) * d.Lc + tid] = tc;
            }
            __syncthreads();
            // ---- s = beta + tau * Az(z / phi), column block by column block (:972), then the denoiser
            for (int c = 0; c < d.Lc; ++c) {
                const T tau_c = (T)tau[c];
                T *s_c = s + (size_t)c * d.Mc;
                T *beta_c = beta + (size_t)c * d.Mc;
                const int b0 = d.col_first[c], b1 = d.col_first[c + 1];
                for (int bi = b0; bi < b1; ++bi) {
                    const BlockDev b = d.blocks[bi];
                    block_transpose<T, LOGH>(buf, tb, b, z + b.r * d.Mr, (T)(1.0 / phi[b.r]), d.Mr, H, tid, tile);
                    const T *bf = reinterpret_cast<const T *>(buf);
                    const T sc = Scale<T>::get(b) * tau_c;
                    const T *prev = bi == b0 ? (t > 0 ? beta_c : nullptr) : s_c;
                    if (fuse_denoiser && bi == b1 - 1) {
                        // last block of the column, unmodulated, M <= 32: a row of 32 consecutive j in a
                        // warp holds whole sections, so the softmax runs on the gathered values in
                        // registers (no second pass over s, one exp per entry, one barrier less)
                        const T inv_tau = (T)1 / tau_c;
                        double psi_part = 0.0, nm_part = 0.0;
                        for (int base = 0; base < d.Mc; base += kThreads * kBatch) {     // warp-uniform trip count
                            int id[kBatch];
                            T pv[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = base + tid + u * kThreads;
                                id[u] = j < d.Mc ? (int)__ldg(b.az_gather + j) : 0;
                                pv[u] = (prev && j < d.Mc) ? prev[j] : (T)0;
                            }
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = base + tid + u * kThreads;
                                const bool live = j < d.Mc;
                                const T v = pv[u] + sc * bf[id[u]];
                                if (live) s_c[j] = v;
                                T mx = live ? v : -INFINITY;
                                if (d.M == 32) mx = warp_max(mx);                 // unrolled full-warp butterfly
                                else for (int off = d.M >> 1; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                                const T e = live ? exp((v - mx) * inv_tau) : (T)0;
                                T sum = e;
                                if (d.M == 32) sum = warp_sum(sum);
                                else for (int off = d.M >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                                if (live) {
                                    const T inv = (T)1 / sum;
                                    const T bta = e * inv;
                                    beta_c[j] = bta;
                                    psi_part += (double)bta * (double)((sum - e) * inv);
                                    if (idx0) {
                                        const int sec = j / d.M;
                                        const int truth = idx0[f * d.L + c * secs_per_col + sec];
                                        const double dlt = (double)bta - ((j - sec * d.M) == truth ? 1.0 : 0.0);
                                        nm_part += dlt * dlt;
                                    } else {
                                        nm_part += (double)bta * (double)bta;
                                    }
                                }
                            }
                        }
                        psi_part = warp_sum(psi_part);
                        nm_part = warp_sum(nm_part);
                        if (lane == 0) {
                            part_psi[c & 1][warp] = psi_part;
                            part_nm[c & 1][warp] = nm_part;
                        }
                    } else {
                        for (int j0 = tid; j0 < d.Mc; j0 += kThreads * kBatch) {
                            int id[kBatch];
                            T pv[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = j0 + u * kThreads;
                                id[u] = j < d.Mc ? (int)__ldg(b.az_gather + j) : 0;
                                pv[u] = (prev && j < d.Mc) ? prev[j] : (T)0;
                            }
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = j0 + u * kThreads;
                                if (j < d.Mc) s_c[j] = pv[u] + sc * bf[id[u]];
                            }
                        }
                    }
                    __syncthreads();
                }
                if (fuse_denoiser && b1 > b0) {              // beta, psi, nmse of this column are done
                    if (tid == 0) {
                        double a = 0.0, bsum = 0.0;
                        for (int w = 0; w < kThreads / 32; ++w) { a += part_psi[c & 1][w]; bsum += part_nm[c & 1][w]; }
                        psi[c] = a;
                        nm[c] = bsum;
                    }
                    continue;
                }
                // denoiser: sections of M entries, beta = softmax(s / tau) (:429-432)
                double psi_part = 0.0, nm_part = 0.0;
                const T inv_tau = (T)1 / tau_c;
                if (d.M >= 32) {
                    const int per_lane = d.M >> 5;
                    for (int sec = warp; p.K == 2 && sec < secs_per_col; sec += kThreads / 32) {
                        // real 2-PSK (:433-441): beta = (e^x - e^-x) / sum_section (e^x + e^-x), x = s / tau,
                        // every exponent shifted by the section's max |x|
                        T *ss = s_c + (size_t)sec * d.M;
                        T mx = (T)0;
                        for (int k = 0; k < per_lane; ++k) mx = fmax(mx, fabs(ss[lane + 32 * k]));
                        mx = warp_max(mx);
                        T sum = (T)0;
                        for (int k = 0; k < per_lane; ++k) {
                            const T v = ss[lane + 32 * k];
                            sum += exp((v - mx) * inv_tau) + exp((-v - mx) * inv_tau);
                        }
                        sum = warp_sum(sum);
                        const T inv = (T)1 / sum;
                        const int truth = idx0 ? idx0[f * d.L + c * secs_per_col + sec] : -1;
                        const int t_loc = truth >= 0 ? (truth & (d.M - 1)) : -1;
                        const double t_val = truth >= d.M ? -1.0 : 1.0;
                        double q = 0.0, e2 = 0.0;
                        for (int k = 0; k < per_lane; ++k) {
                            const int j = lane + 32 * k;
                            const T v = ss[j];
                            const T ep = exp((v - mx) * inv_tau), em = exp((-v - mx) * inv_tau);
                            const T bta = (ep - em) * inv;
                            beta[(size_t)c * d.Mc + (size_t)sec * d.M + j] = bta;
                            // 1 - sum beta^2 = sum p (1 - p) + 4 e+ e- / bot^2,  p = (e+ + e-) / bot: no cancellation
                            const T pj = (ep + em) * inv;
                            q += (double)pj * (double)((sum - (ep + em)) * inv) + 4.0 * (double)(ep * inv) * (double)(em * inv);
                            const double dlt = (double)bta - (j == t_loc ? t_val : 0.0);
                            e2 += dlt * dlt;
                        }
                        psi_part += q;
                        nm_part += e2;
                    }
                    for (int sec = warp; p.K != 2 && sec < secs_per_col; sec += kThreads / 32) {
                        T *ss = s_c + (size_t)sec * d.M;
                        T mx = -INFINITY;
                        for (int k = 0; k < per_lane; ++k) mx = fmax(mx, ss[lane + 32 * k]);
                        mx = warp_max(mx);
                        T sum = (T)0;
                        for (int k = 0; k < per_lane; ++k) sum += exp((ss[lane + 32 * k] - mx) * inv_tau);
                        sum = warp_sum(sum);
                        const T inv = (T)1 / sum;
                        const int truth = idx0 ? idx0[f * d.L + c * secs_per_col + sec] : -1;
                        double q = 0.0, e2 = 0.0;
                        for (int k = 0; k < per_lane; ++k) {
                            const int j = lane + 32 * k;
                            const T e = exp((ss[j] - mx) * inv_tau);
                            const T bta = e * inv;
                            beta[(size_t)c * d.Mc + (size_t)sec * d.M + j] = bta;
                            // 1 - sum beta^2 = sum beta (1 - beta); 1 - beta = (sum - e) / sum without cancellation
                            q += (double)bta * (double)((sum - e) * inv);
                            const double dlt = (double)bta - (j == truth ? 1.0 : 0.0);
                            e2 += dlt * dlt;
                        }
                        psi_part += q;
                        nm_part += e2;
                    }
                } else {
                    // M < 32: one lane per entry, 32/M sections per warp pass
                    const int spw = 32 / d.M;                 // sections per warp
                    const int sub = lane / d.M, k = lane % d.M;
                    for (int sec0 = warp * spw; sec0 < secs_per_col; sec0 += (kThreads / 32) * spw) {
                        const int sec = sec0 + sub;
                        const bool live = sec < secs_per_col;
                        const T v = live ? s_c[(size_t)sec * d.M + k] : (T)0;
                        T mx = p.K == 2 ? fabs(v) : v;
                        for (int off = d.M >> 1; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                        const T e = exp((v - mx) * inv_tau);
                        const T em = p.K == 2 ? exp((-v - mx) * inv_tau) : (T)0;
                        T sum = e + em;
                        for (int off = d.M >> 1; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                        if (live && p.K == 2) {
                            const T inv = (T)1 / sum;
                            const T bta = (e - em) * inv;
                            beta[(size_t)c * d.Mc + (size_t)sec * d.M + k] = bta;
                            const T pj = (e + em) * inv;
                            psi_part += (double)pj * (double)((sum - (e + em)) * inv) + 4.0 * (double)(e * inv) * (double)(em * inv);
                            const int truth = idx0 ? idx0[f * d.L + c * secs_per_col + sec] : -1;
                            const double tv = truth < 0 || k != (truth & (d.M - 1)) ? 0.0 : (truth >= d.M ? -1.0 : 1.0);
                            const double dlt = (double)bta - tv;
                            nm_part += dlt * dlt;
                        } else if (live) {
                            const T inv = (T)1 / sum;
                            const T bta = e * inv;
                            beta[(size_t)c * d.Mc + (size_t)sec * d.M + k] = bta;
                            psi_part += (double)bta * (double)((sum - e) * inv);
                            const int truth = idx0 ? idx0[f * d.L + c * secs_per_col + sec] : -1;
                            const double dlt = (double)bta - (k == truth ? 1.0 : 0.0);
                            nm_part += dlt * dlt;
                        }
                    }
                }
                psi_part = warp_sum(psi_part);
                nm_part = warp_sum(nm_part);
                if (lane == 0) {
                    part_psi[c & 1][warp] = psi_part;
                    part_nm[c & 1][warp] = nm_part;
                }
                __syncthreads();
                if (tid == 0) {
                    double a = 0.0, bsum = 0.0;
                    for (int w = 0; w < kThreads / 32; ++w) { a += part_psi[c & 1][w]; bsum += part_nm[c & 1][w]; }
                    psi[c] = a;
                    nm[c] = bsum;
                }
            }
            // ---- psi, nmse, early stop (:976-986)
            if (tid == 0) {
                int stop = t > 0;
                for (int c = 0; c < d.Lc; ++c) {
                    psi[c] /= sec_norm;
                    if (!(fabs(psi[c] - psi_prev[c]) <= p.atol + p.rtol * fabs(psi_prev[c]))) stop = 0;
                    if (o.psi_trace) o.psi_trace[(f * p.t_max + t) * d.Lc + c] = psi[c];
                    if (o.nmse) o.nmse[(f * p.t_max + t + 1) * d.Lc + c] = (float)(nm[c] / sec_norm);
                }
                s_stop = stop;
                if (stop && o.nmse)                                   // nmse[t:] = nmse[t]  (:985)
                    for (int tt = t + 1; tt < p.t_max; ++tt)
                        for (int c = 0; c < d.Lc; ++c)
                            o.nmse[(f * p.t_max + tt) * d.Lc + c] = o.nmse[(f * p.t_max + t) * d.Lc + c];
            }
            __syncthreads();
            if (s_stop) break;
        }
        const int t_final = (t < p.t_max - 1) ? t + 1 : p.t_max - 1;   // t + 1 after a break, else last t + 1

        // ---- outputs: MAP decision on s (:997), final psi, optional soft state
        if (tid == 0) o.t_final[f] = t_final;
        if (tid < d.Lc) o.psi[f * d.Lc + tid] = psi[tid];
        for (int sec = warp; sec < d.L; sec += kThreads / 32) {
            const T *ss = s + (size_t)sec * d.M;
            T best = -INFINITY;
            int arg = 0x7fffffff;
            for (int j = lane; j < d.M; j += 32) {
                const T sv = ss[j];
                const T v = p.K == 2 ? fabs(sv) : sv;                        // K = 2: largest magnitude (:488-492)
                if (v > best) { best = v; arg = j + ((p.K == 2 && sv < (T)0) ? d.M : 0); }
            }
            const int loc_mask = d.M - 1;                                    // ties go to the lowest location
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
                if (ob > best || (ob == best && (oa & loc_mask) < (arg & loc_mask))) { best = ob; arg = oa; }
            }
            if (lane == 0) o.idx[f * d.L + sec] = arg;
        }
        if (o.beta_soft) {
            T *dst = reinterpret_cast<T *>(o.beta_soft) + f * LM;
            for (int i = tid; i < LM; i += kThreads) dst[i] = beta[i];
        }
        if (o.s_last) {
            T *dst = reinterpret_cast<T *>(o.s_last) + f * LM;
            for (int i = tid; i < LM; i += kThreads) dst[i] = s[i];
        }
        __syncthreads();
    }
}

#include "sparc_amp_pair.cuh"

}  // namespace

// ------------------------------------------------------------------ host side
struct lsb_sparc_design {
    int device = 0;
    DesignDev dev{};
    std::vector<void *> allocs;
    FftTables<float> tb32{};
    FftTables<double> tb64{};
    unsigned long long *d_queue = nullptr;
    std::atomic<unsigned> next_queue{0};
    std::mutex mu;
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t scratch_event = nullptr;
    int amp_grid[2] = {0, 0};
    // pair kernels (fp32, w = 2048 / 4096 / 8192): tables with their own swizzle, see sparc_pair.cuh
    bool has_pair = false, full_columns = true;
    PairDev pair{};
    int amp_grid_pair = 0;
    // host-pointer API staging (lazily sized), one host call at a time per handle
    std::mutex stage_mu;
    cudaStream_t stage_stream = nullptr;
    void *stage = nullptr;
    size_t stage_bytes = 0;
};

namespace {

template <typename T>
int dev_upload(lsb_sparc_design *g, const T **dst, const std::vector<T> &src)
{
    void *p = nullptr;
    LSB_CUDA(cudaMalloc(&p, src.size() * sizeof(T)));
    g->allocs.push_back(p);
    LSB_CUDA(cudaMemcpy(p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return LSB_OK;
}

template <typename T, typename T2>
int make_tables(lsb_sparc_design *g, FftTables<T> *tb, int h)
{
    const int H = 1 << h;
    const double N = 2.0 * H;
    const double PI = 3.14159265358979323846;
    std::vector<T2> tw(H), pt(H + 1);
    for (int k = 0; k < H; ++k) tw[k] = T2{(T)cos(2 * PI * k / H), (T)-sin(2 * PI * k / H)};
    for (int k = 0; k <= H; ++k) pt[k] = T2{(T)cos(PI * k / (2 * N)), (T)-sin(PI * k / (2 * N))};
    std::vector<T2> ptr(H / 2 + 1);
    for (int t = 0; t < H / 2; ++t) {
        int k = 0;
        for (int i = 0; i < h - 1; ++i) k |= ((t >> i) & 1) << (h - 2 - i);   // brev over h-1 bits
        ptr[t] = pt[k];
    }
    ptr[H / 2] = pt[H / 2];
    int rc;
    if ((rc = dev_upload(g, &tb->tw, tw))) return rc;
    if ((rc = dev_upload(g, &tb->pt, pt))) return rc;
    if ((rc = dev_upload(g, &tb->ptr, ptr))) return rc;
    tb->h = h;
    return LSB_OK;
}

inline int brev_host(int k, int h)
{
    int r = 0;
    for (int i = 0; i < h; ++i) r |= ((k >> i) & 1) << (h - 1 - i);
    return r;
}

int check_device(const lsb_sparc_design *g)
{
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    return LSB_OK;
}

// shared memory per CTA; gbuf = FFT buffer kept in global memory instead
size_t amp_smem(const DesignDev &d, size_t elem, bool gbuf)
{
    if (gbuf && d.h == 15) return staged_tile_offset(d.n, elem) + (size_t)kTilePts * 2 * elem;   // z + staged-FFT tile
    return ((gbuf ? 0 : (size_t)d.w) + (size_t)d.n) * elem;
}
size_t ab_smem(const DesignDev &d, size_t elem, bool gbuf) { return ((gbuf ? 0 : (size_t)d.w) + (size_t)d.n) * elem; }
size_t az_smem(const DesignDev &d, size_t elem, bool gbuf) { return (gbuf ? 16 : (size_t)d.w * elem); }
bool needs_gbuf(const DesignDev &d, size_t elem) { return amp_smem(d, elem, false) + 8192 > (size_t)lsb::max_smem_optin(); }

// (re)allocate the per-handle scratch and chain this launch behind the previous user
int ensure_scratch(lsb_sparc_design *g, size_t need, cudaStream_t st)
{
    if (g->scratch_bytes < need) {
        LSB_CUDA(cudaDeviceSynchronize());
        if (g->d_scratch) cudaFree(g->d_scratch);
        g->d_scratch = nullptr;
        g->scratch_bytes = 0;
        LSB_CUDA(cudaMalloc(&g->d_scratch, need));
        g->scratch_bytes = need;
    }
    if (!g->scratch_event) LSB_CUDA(cudaEventCreateWithFlags(&g->scratch_event, cudaEventDisableTiming));
    else LSB_CUDA(cudaStreamWaitEvent(st, g->scratch_event, 0));
    return LSB_OK;
}

template <typename T> const FftTables<T> &tables(const lsb_sparc_design *g);
template <> const FftTables<float> &tables<float>(const lsb_sparc_design *g) { return g->tb32; }
template <> const FftTables<double> &tables<double>(const lsb_sparc_design *g) { return g->tb64; }

// LSB_AMP_PAIR=0 in the environment forces the one-frame kernels (A/B runs, tests of the generic path)
bool pair_enabled()
{
    static const bool on = [] {
        const char *e = getenv("LSB_AMP_PAIR");
        return !(e && e[0] == '0');
    }();
    return on;
}

bool pair_fits(const lsb_sparc_design *g)
{
    return g->has_pair && pair_enabled() &&
           pair_amp_dyn_smem(g->dev.h, g->dev.n, g->dev.Lr, g->dev.Lc, false) + 4096 <= (size_t)lsb::max_smem_optin();
}

// Mirror pair s (0..3) of thread t in the fused first pass of the transposed transform (post = false) or the fused
// last pass of the forward transform (post = true) at H = 2^h, Q = H/4 (sparc_pair.cuh: PairTables)
inline int pair_fused_k(int h, int t, int s, bool post)
{
    const int Q = 1 << (h - 2);
    if (t == 0) return s == 0 ? Q / 2 : s == 1 ? 3 * Q / 2 : s == 2 ? Q : 2 * Q;
    const int b = post ? brev_host(2 * t, h - 2) : t;
    return s == 0 ? b : s == 1 ? b + Q : s == 2 ? Q - b : 2 * Q - b;
}

// Order of the row gather of one block (sparc_pair.cuh: PairBlock::fgat).  Row i reads the float2 slot id[i] of the FFT
// buffer and updates entry row0 + i of the residual; a 64-bit shared-memory access of a warp is served one half-warp
// at a time and is conflict free when its 16 lanes hit 16 different bank pairs.  Rows are the edges of a bipartite
// multigraph between the bank pairs of the residual ((row0 + i) mod 16) and those of the buffer (id mod 16); a proper
// edge colouring with Delta = max degree colours (Koenig: alternating-path recolouring) splits them into Delta half-warp
// groups that are conflict free on both sides.  Group c goes to half-warp c mod 16 of the CTA, slot c / 16 of its
// threads, lane = the row's residual bank pair.  Returns [chunks][256][4] words row | id << 16, 0xFFFFFFFF = empty.
std::vector<uint32_t> schedule_row_gather(const std::vector<uint16_t> &id, int row0, int *chunks_out)
{
    const int n = (int)id.size();
    std::vector<int> eu(n), ev(n), col(n, -1);
    int degu[16] = {0}, degv[16] = {0};
    for (int i = 0; i < n; ++i) {
        eu[i] = (row0 + i) & 15;
        ev[i] = id[i] & 15;
        ++degu[eu[i]];
        ++degv[ev[i]];
    }
    int delta = 1;
    for (int k = 0; k < 16; ++k) delta = std::max(delta, std::max(degu[k], degv[k]));
    std::vector<int> atu((size_t)16 * delta, -1), atv((size_t)16 * delta, -1);     // edge of colour c at a node, or -1
    std::vector<int> path;
    for (int e = 0; e < n; ++e) {
        const int u = eu[e], v = ev[e];
        int a = 0, b = 0;
        while (atu[(size_t)u * delta + a] >= 0) ++a;       // free at u
        while (atv[(size_t)v * delta + b] >= 0) ++b;       // free at v
        if (a != b) {
            // alternating a / b path from v (it cannot reach u: left nodes are entered through colour a, free at u)
            path.clear();
            int x = v, c = a;
            bool right = true;
            for (;;) {
                const int e2 = right ? atv[(size_t)x * delta + c] : atu[(size_t)x * delta + c];
                if (e2 < 0) break;
                path.push_back(e2);
                x = right ? eu[e2] : ev[e2];
                right = !right;
                c = c == a ? b : a;
            }
            for (int e2 : path) { atu[(size_t)eu[e2] * delta + col[e2]] = -1; atv[(size_t)ev[e2] * delta + col[e2]] = -1; }
            for (int e2 : path) {
                col[e2] = col[e2] == a ? b : a;
                atu[(size_t)eu[e2] * delta + col[e2]] = e2;
                atv[(size_t)ev[e2] * delta + col[e2]] = e2;
            }
        }
        col[e] = a;
        atu[(size_t)u * delta + a] = e;
        atv[(size_t)v * delta + a] = e;
    }
    const int slots = (delta + 15) / 16, chunks = (slots + 3) / 4;
    std::vector<uint32_t> out((size_t)chunks * kPairThreads * 4, 0xFFFFFFFFu);
    for (int e = 0; e < n; ++e) {
        const int c = col[e], slot = c / 16, tid = (c % 16) * 16 + eu[e];
        out[((size_t)(slot / 4) * kPairThreads + tid) * 4 + (slot % 4)] = (uint32_t)e | ((uint32_t)id[e] << 16);
    }
    *chunks_out = chunks;
    return out;
}

// Host-side tables of one non-zero block for the pair kernels (index conventions in sparc_pair.cuh).
int build_pair_block(lsb_sparc_design *g, int h, int Mr, int Mc, const uint32_t *o0, const uint32_t *o1, PairBlock *pb)
{
    const int H = 1 << h, N = 2 * H, Hh = H / 2, NB = H / 8;
    std::vector<int> inv0(N + 1, -1), inv1(N + 1, -1);
    for (int i = 0; i < Mr; ++i) inv0[o0[i]] = i;
    for (int j = 0; j < Mc; ++j) inv1[o1[j]] = j;
    auto ent = [](int v) -> uint16_t { return v < 0 ? pairfft::kNone : (uint16_t)v; };
    auto makhoul = [&](int p) -> int { return (p & 1) ? N - 1 - (p - 1) / 2 : p / 2; };
    // forward, first radix-8 pass: butterfly q holds complex elements e = q + m H/8, packed z[e] = v[2e] + i v[2e+1],
    // v the Makhoul permutation of the zero-padded column block.  Layout [half][q][4 words], word = re | im << 16.
    std::vector<uint16_t> fsrc((size_t)2 * NB * 8);
    for (int q = 0; q < NB; ++q)
        for (int m = 0; m < 8; ++m)
            for (int comp = 0; comp < 2; ++comp) {
                const int mpos = 2 * (q + m * NB) + comp;
                const int pidx = mpos < N / 2 ? 2 * mpos : 2 * (N - 1 - mpos) + 1;
                fsrc[((size_t)(m / 4) * NB + q) * 8 + 2 * (m % 4) + comp] = ent(inv1[pidx]);
            }
    // forward gather: after the post-processing slot brev(k) holds (Y[k], Y[N-k]), slot 0 holds (Y[0], Y[H])
    std::vector<uint16_t> fgat(Mr);
    for (int i = 0; i < Mr; ++i) {
        const int K = (int)o0[i];
        int slot, comp;
        if (K < H) { slot = K; comp = 0; }
        else if (K == H) { slot = 0; comp = 1; }
        else { slot = N - K; comp = 1; }
        fgat[i] = (uint16_t)(2 * pairfft::pswz_rt(h, brev_host(slot, h)) + comp);
    }
    // transposed sources: pair k needs c[k], c[N-k] (slot k) and c[H-k], c[H+k] (slot H-k); slot 0 is (c[0], c[H])
    std::vector<uint16_t> tsrc((size_t)(Hh + 1) * 4);
    for (int k = 0; k <= Hh; ++k) {
        tsrc[(size_t)k * 4 + 0] = ent(inv0[k]);
        tsrc[(size_t)k * 4 + 1] = ent(k == 0 ? inv0[H] : inv0[N - k]);
        tsrc[(size_t)k * 4 + 2] = ent(inv0[H - k]);
        tsrc[(size_t)k * 4 + 3] = ent(inv0[H + k]);
    }
    if (pairfft::fused_inv_rt(h)) {             // per-thread order of the fused first pass (sparc_pair.cuh), then k = 0
        std::vector<uint16_t> t4((size_t)(4 * kPairThreads + 1) * 4);
        for (int t = 0; t < kPairThreads; ++t)
            for (int s = 0; s < 4; ++s) {
                const int k = pair_fused_k(h, t, s, false);
                for (int e = 0; e < 4; ++e) t4[(((size_t)(s / 2) * kPairThreads + t) * 2 + (s % 2)) * 4 + e] = tsrc[(size_t)k * 4 + e];
            }
        for (int e = 0; e < 4; ++e) t4[(size_t)4 * kPairThreads * 4 + e] = tsrc[e];
        tsrc.swap(t4);
    }
    std::vector<uint16_t> tgat(Mc);
    for (int j = 0; j < Mc; ++j) {
        const int m = makhoul((int)o1[j]);
        tgat[j] = (uint16_t)(2 * pairfft::pswz_t_rt(h, brev_host(m >> 1, h)) + (m & 1));
    }
    int rc;
    const uint16_t *p16 = nullptr;
    if ((rc = dev_upload(g, &p16, fsrc))) return rc;
    pb->fsrc = reinterpret_cast<const uint4 *>(p16);
    {
        int chunks = 0;
        const std::vector<uint32_t> sched = schedule_row_gather(fgat, pb->r * Mr, &chunks);
        const uint32_t *p32 = nullptr;
        if ((rc = dev_upload(g, &p32, sched))) return rc;
        pb->fgat = reinterpret_cast<const uint4 *>(p32);
        pb->gat_chunks = chunks;
    }
    if ((rc = dev_upload(g, &p16, tsrc))) return rc;
    pb->tsrc = reinterpret_cast<const uint2 *>(p16);
    tgat.resize((size_t)pairfft::chunk_count(Mc) * kPairThreads * 4, pairfft::kNone);   // four consecutive columns per thread
    if ((rc = dev_upload(g, &p16, tgat))) return rc;
    pb->tgat = reinterpret_cast<const uint2 *>(p16);
    return LSB_OK;
}

// Per-size tables of the pair kernels: 0.5 pt[k] in natural and in bit-reversed pair order, and the
// swizzled slots of the post-processing pairs.
int build_pair_tables(lsb_sparc_design *g, int h)
{
    const int H = 1 << h, Hh = H / 2;
    const double N = 2.0 * H, PI = 3.14159265358979323846;
    auto entry = [&](int k) {
        const double a = PI * k / (2 * N);
        return float2{(float)(0.5 * cos(a)), (float)(-0.5 * sin(a))};
    };
    std::vector<float2> pre(Hh + 1), post(Hh + 1);
    std::vector<uint32_t> slot(Hh + 1);
    for (int k = 0; k <= Hh; ++k) pre[k] = entry(k);
    for (int t = 0; t <= Hh; ++t) {
        const int k = t == Hh ? Hh : brev_host(t, h - 1);
        post[t] = entry(k);
        const int pa = pairfft::pswz_rt(h, brev_host(k, h)), pb = pairfft::pswz_rt(h, brev_host((H - k) & (H - 1), h));
        slot[t] = (uint32_t)pa | ((uint32_t)pb << 16);
    }
    if (pairfft::fused_inv_rt(h)) {
        pre.assign((size_t)4 * kPairThreads, float2{0.f, 0.f});
        for (int t = 0; t < kPairThreads; ++t)
            for (int s = 0; s < 4; ++s) pre[(size_t)t * 4 + s] = entry(pair_fused_k(h, t, s, false));
    }
    if (pairfft::fused_fwd_rt(h)) {
        post.assign((size_t)4 * kPairThreads, float2{0.f, 0.f});
        for (int t = 0; t < kPairThreads; ++t)
            for (int s = 0; s < 4; ++s) post[(size_t)t * 4 + s] = entry(pair_fused_k(h, t, s, true));
    }
    g->pair.tb.tw = g->tb32.tw;
    int rc;
    if ((rc = dev_upload(g, &g->pair.tb.pre, pre))) return rc;
    if ((rc = dev_upload(g, &g->pair.tb.post, post))) return rc;
    return dev_upload(g, &g->pair.tb.post_slot, slot);
}

template <int LOGH>
int launch_op_pair_h(lsb_sparc_design *g, bool transpose, const float *in, float *out, int64_t n_frames, cudaStream_t st)
{
    const size_t smem = pair_op_dyn_smem(LOGH, g->dev.n);
    long long grid = (long long)lsb::sm_count() * 3;
    if (grid > (n_frames + 1) / 2) grid = (n_frames + 1) / 2;
    if (transpose) {
        LSB_CUDA(lsb::ensure_dynamic_smem(az_pair_kernel<LOGH>, smem));
        az_pair_kernel<LOGH><<<(unsigned)grid, kPairThreads, smem, st>>>(g->dev, g->pair, in, out, (long long)n_frames);
    } else {
        LSB_CUDA(lsb::ensure_dynamic_smem(ab_pair_kernel<LOGH>, smem));
        ab_pair_kernel<LOGH><<<(unsigned)grid, kPairThreads, smem, st>>>(g->dev, g->pair, in, out, (long long)n_frames);
    }
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

int launch_op_pair(lsb_sparc_design *g, bool transpose, const float *in, float *out, int64_t n_frames, cudaStream_t st)
{
    switch (g->dev.h) {
        case 10: return launch_op_pair_h<10>(g, transpose, in, out, n_frames, st);
        case 11: return launch_op_pair_h<11>(g, transpose, in, out, n_frames, st);
        default: return launch_op_pair_h<12>(g, transpose, in, out, n_frames, st);
    }
}

template <typename T, bool GBUF>
int launch_op(lsb_sparc_design *g, bool transpose, const T *in, T *out, int64_t n_frames, long long grid,
              cudaStream_t st)
{
    const size_t smem = transpose ? az_smem(g->dev, sizeof(T), GBUF) : ab_smem(g->dev, sizeof(T), GBUF);
    LSB_REQUIRE(smem <= (size_t)lsb::max_smem_optin(), "design needs %zu B of shared memory (max %d)", smem,
                lsb::max_smem_optin());
    T *gbuf = GBUF ? (T *)g->d_scratch : nullptr;
    if (transpose) {
        LSB_CUDA(lsb::ensure_dynamic_smem(az_kernel<T, GBUF>, smem));
        az_kernel<T, GBUF><<<(unsigned)grid, kThreads, smem, st>>>(g->dev, tables<T>(g), in, out,
                                                                  (long long)n_frames, gbuf);
    } else {
        LSB_CUDA(lsb::ensure_dynamic_smem(ab_kernel<T, GBUF>, smem));
        ab_kernel<T, GBUF><<<(unsigned)grid, kThreads, smem, st>>>(g->dev, tables<T>(g), in, out,
                                                                  (long long)n_frames, gbuf);
    }
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

template <typename T>
int run_op(const lsb_sparc_design *cg, bool transpose, const T *in, T *out, int64_t n_frames, cudaStream_t st)
{
    lsb_sparc_design *g = const_cast<lsb_sparc_design *>(cg);
    if (n_frames == 0) return LSB_OK;              // an empty batch is valid with any pointers
    LSB_REQUIRE(g && in && out, "null pointer");
    LSB_REQUIRE(n_frames >= 0, "negative frame count");
    if (n_frames == 0) return LSB_OK;
    int rc = check_device(g);
    if (rc) return rc;
    if constexpr (sizeof(T) == 4) {
        if (pair_fits(g)) return launch_op_pair(g, transpose, in, out, n_frames, st);
    }
    long long grid = (long long)lsb::sm_count() * 4;
    if (grid > n_frames) grid = n_frames;
    if (!needs_gbuf(g->dev, sizeof(T))) return launch_op<T, false>(g, transpose, in, out, n_frames, grid, st);
    std::lock_guard<std::mutex> lock(g->mu);
    if ((rc = ensure_scratch(g, (size_t)grid * g->dev.w * sizeof(T), st))) return rc;
    rc = launch_op<T, true>(g, transpose, in, out, n_frames, grid, st);
    if (!rc) LSB_CUDA(cudaEventRecord(g->scratch_event, st));
    return rc;
}

template <typename T>
using AmpKernelFn = void (*)(DesignDev, FftTables<T>, AmpParams, const T *, const int32_t *, AmpOut, T *, long long,
                             unsigned long long *);

// fp32 (the throughput mode) gets the compile-time-size transforms for the sizes the
// BASELINE configs and the golden cases use (w = 2048, 4096, 8192); anything else, and
// the fp64 parity mode, runs the size-generic passes.
template <typename T>
AmpKernelFn<T> pick_amp_kernel(bool gbuf, int h)
{
    if (gbuf) return amp_kernel<T, true, 0>;
    if constexpr (sizeof(T) == 4) {
        switch (h) {
            case 10: return amp_kernel<T, false, 10>;
            case 11: return amp_kernel<T, false, 11>;
            case 12: return amp_kernel<T, false, 12>;
            default: break;
        }
    }
    return amp_kernel<T, false, 0>;
}

template <int LOGH>
int launch_amp_pair_h(lsb_sparc_design *g, const float *y, int64_t n_frames, const AmpParams &p, const int32_t *idx0,
                      const AmpOut &o, cudaStream_t st)
{
    static_assert(sizeof(PairBlock) % sizeof(uint2) == 0, "descriptor copied as 8-byte words");
    const bool cached = LSB_PAIR_CACHE && pair_design_cached(g->dev.Lr, g->dev.Lc, g->dev.n_blocks);
    // beta column staged in shared memory (2 CTAs per SM): transforms with staged tables, column block <= 8 KB
    const bool bstage = LSB_PAIR_BSTAGE && cached && pairfft::stage_fwd<LOGH>() && g->dev.Mc <= kBetaStageBytes / 8 &&
                        g->dev.Mc % 2 == 0 &&
                        pair_amp_dyn_smem(LOGH, g->dev.n, g->dev.Lr, g->dev.Lc, true) + 4096 <= (size_t)lsb::max_smem_optin();
    const size_t smem = pair_amp_dyn_smem(LOGH, g->dev.n, g->dev.Lr, g->dev.Lc, bstage);
    auto kernel = bstage ? amp_pair_kernel<LOGH, true, true> : cached ? amp_pair_kernel<LOGH, true, false> : amp_pair_kernel<LOGH, false, false>;
    LSB_CUDA(lsb::ensure_dynamic_smem(kernel, smem));
    if (g->amp_grid_pair == 0) {
        int per_sm = 0;
        LSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPairThreads, smem));
        if (per_sm < 1) per_sm = 1;
        g->amp_grid_pair = per_sm * lsb::sm_count();
    }
    const long long pairs = (n_frames + 1) / 2;
    const long long grid = g->amp_grid_pair < pairs ? g->amp_grid_pair : pairs;
    const size_t need = (size_t)g->amp_grid_pair * ((size_t)g->dev.L * g->dev.M + (size_t)g->dev.Mc) * sizeof(float2);
    std::lock_guard<std::mutex> lock(g->mu);
    int rc;
    if ((rc = ensure_scratch(g, need, st))) return rc;
    unsigned long long *queue = g->d_queue + (g->next_queue.fetch_add(1) % kCounters);
    LSB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
    kernel<<<(unsigned)grid, kPairThreads, smem, st>>>(g->dev, g->pair, p, y, idx0, o, (float2 *)g->d_scratch,
                                                      (long long)n_frames, queue);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    LSB_CUDA(cudaEventRecord(g->scratch_event, st));
    return LSB_OK;
}

int launch_amp_pair(lsb_sparc_design *g, const float *y, int64_t n_frames, const AmpParams &p, const int32_t *idx0,
                    const AmpOut &o, cudaStream_t st)
{
    switch (g->dev.h) {
        case 10: return launch_amp_pair_h<10>(g, y, n_frames, p, idx0, o, st);
        case 11: return launch_amp_pair_h<11>(g, y, n_frames, p, idx0, o, st);
        default: return launch_amp_pair_h<12>(g, y, n_frames, p, idx0, o, st);
    }
}

template <typename T>
int run_amp(lsb_sparc_design *g, const T *y, int64_t n_frames, const lsb_sparc_amp_params *pp,
            const int32_t *idx0, const lsb_sparc_amp_out *po, cudaStream_t st)
{
    if (n_frames == 0) return LSB_OK;              // an empty batch is valid with any pointers
    LSB_REQUIRE(g && y && pp && po && po->idx && po->t_final && po->psi, "null pointer");
    LSB_REQUIRE(pp->t_max > 1, "t_max must be > 1");
    LSB_REQUIRE(pp->phi_est_method == 1 || pp->phi_est_method == 2, "phi_est_method must be 1 or 2");
    LSB_REQUIRE(n_frames >= 0, "negative frame count");
    LSB_REQUIRE(!po->nmse || idx0, "nmse output needs the true section indices idx0");
    if (n_frames == 0) return LSB_OK;
    int rc = check_device(g);
    if (rc) return rc;
    LSB_REQUIRE(pp->modulation >= 0 && pp->modulation <= 2, "modulation must be 1 (none) or 2 (real 2-PSK)");
    if constexpr (sizeof(T) == 4) {
        // throughput path: two frames per CTA on the packed-fp32 transforms (sparc_amp_pair.cuh)
        if (pair_fits(g) && g->full_columns && pp->modulation != 2 && g->dev.M >= 4 && g->dev.M <= 32 && !po->beta_soft && !po->s_last) {
            AmpParams p{pp->awgn_var, pp->rtol, pp->atol, pp->t_max, pp->phi_est_method, 1};
            AmpOut o{po->idx, po->t_final, po->psi, po->nmse, nullptr, nullptr, po->tau_trace, po->psi_trace};
            return launch_amp_pair(g, y, n_frames, p, idx0, o, st);
        }
    }
    const bool gbuf = needs_gbuf(g->dev, sizeof(T));
    const size_t smem = amp_smem(g->dev, sizeof(T), gbuf);
    LSB_REQUIRE(smem + 8192 <= (size_t)lsb::max_smem_optin(),
                "design needs %zu B of shared memory per CTA (n = %d): unsupported", smem, g->dev.n);
    const AmpKernelFn<T> kernel = pick_amp_kernel<T>(gbuf, g->dev.h);
    int &cached = g->amp_grid[sizeof(T) == 8];
    LSB_CUDA(lsb::ensure_dynamic_smem(kernel, smem));      // per (kernel, device), raise-only: other designs share it
    if (cached == 0) {
        int per_sm = 0;
        LSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem));
        if (per_sm < 1) per_sm = 1;
        if (gbuf) {
            // the FFT buffers of all resident CTAs should stay in L2 (w elements each): more CTAs than
            // that turn every radix pass into DRAM traffic
            int l2 = 0;
            LSB_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, g->device));
            int cap = (int)((0.6 * l2) / ((double)g->dev.w * sizeof(T)) / lsb::sm_count());   // measured: 2 CTAs/SM best for w = 65536 fp32
            if (cap < 1) cap = 1;
            if (cap < per_sm) per_sm = cap;
        }
        cached = per_sm * lsb::sm_count();
    }
    long long grid = cached < n_frames ? cached : n_frames;
    const size_t need = (size_t)cached * (2 * (size_t)g->dev.L * g->dev.M + (gbuf ? (size_t)g->dev.w : 0)) * sizeof(double);
    std::lock_guard<std::mutex> lock(g->mu);
    if ((rc = ensure_scratch(g, need, st))) return rc;
    unsigned long long *queue = g->d_queue + (g->next_queue.fetch_add(1) % kCounters);
    LSB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
    AmpParams p{pp->awgn_var, pp->rtol, pp->atol, pp->t_max, pp->phi_est_method, pp->modulation == 2 ? 2 : 1};
    AmpOut o{po->idx, po->t_final, po->psi, po->nmse, po->beta_soft, po->s_last, po->tau_trace, po->psi_trace};
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(g->dev, tables<T>(g), p, y, idx0, o, (T *)g->d_scratch,
                                                   (long long)n_frames, queue);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    LSB_CUDA(cudaEventRecord(g->scratch_event, st));
    return LSB_OK;
}

}  // namespace

extern "C" {

int lsb_sparc_schedule_row_gather(const uint16_t *slot_ids, int n_rows, int row0, uint32_t *words, int64_t max_words,
                                  int *chunks)
{
    LSB_REQUIRE(slot_ids && words && chunks && n_rows > 0 && n_rows <= 65535 && row0 >= 0, "bad argument");
    int c = 0;
    const std::vector<uint32_t> out = schedule_row_gather(std::vector<uint16_t>(slot_ids, slot_ids + n_rows), row0, &c);
    LSB_REQUIRE((int64_t)out.size() <= max_words, "output needs %zu words", out.size());
    std::copy(out.begin(), out.end(), words);
    *chunks = c;
    return LSB_OK;
}


int lsb_sparc_design_create(const double *W, int Lr, int Lc, int L, int M, int n, const uint32_t *order0,
                            const uint32_t *order1, lsb_sparc_design **out)
{
    LSB_REQUIRE(W && order0 && order1 && out, "null pointer");
    LSB_REQUIRE(Lr >= 1 && Lr <= kMaxLr && Lc >= 1 && Lc <= kMaxLc, "base matrix %dx%d outside [1,%d]x[1,%d]", Lr,
                Lc, kMaxLr, kMaxLc);
    LSB_REQUIRE(L > 0 && M > 0 && (M & (M - 1)) == 0 && n > 0, "bad L / M / n");
    LSB_REQUIRE(((long long)L * M) % Lc == 0 && n % Lr == 0 && L % Lc == 0, "Lc must divide L and LM, Lr must divide n");
    const int Mr = n / Lr, Mc = (int)((long long)L * M / Lc);
    LSB_REQUIRE(Mc % M == 0, "column blocks must hold whole sections");
    int w = 1;
    while (w < (Mr > Mc ? Mr : Mc) + 1) w <<= 1;
    LSB_REQUIRE(w >= 4 && w <= 65536, "transform size %d outside [4, 65536] (index tables are 16-bit)", w);
    int h = 0;
    while ((2 << h) < w) ++h;   // H = w/2 = 2^h
    const int H = w / 2, N = w;

    lsb_sparc_design *g = new (std::nothrow) lsb_sparc_design();
    if (!g) return LSB_ERR_NOMEM;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    int rc = LSB_OK;
    std::vector<BlockDev> blocks;
    std::vector<PairBlock> pair_blocks;
    const bool want_pair = h >= 10 && h <= 12;
    std::vector<int> col_first(Lc + 1, 0);
    std::vector<uint16_t> t_abs(Mc), t_abg(Mr), t_azs(Mr), t_azg(Mc);
    auto slot_of_freq = [&](int K, bool reversed) -> int {   // where (Y[K]) / (c[K]) lives
        int slot, comp;
        if (K < H) { slot = K; comp = 0; }
        else if (K == H) { slot = 0; comp = 1; }
        else { slot = N - K; comp = 1; }
        if (reversed) slot = brev_host(slot, h);
        return 2 * swz(slot) + comp;
    };
    auto makhoul = [&](int p) -> int { return (p & 1) ? N - 1 - (p - 1) / 2 : p / 2; };
    for (int c = 0; c < Lc && !rc; ++c) {
        col_first[c] = (int)blocks.size();
        for (int r = 0; r < Lr && !rc; ++r) {
            const double wrc = W[r * Lc + c];
            if (wrc == 0.0) continue;
            if (wrc < 0.0) { lsb::set_error("negative base-matrix entry"); rc = LSB_ERR_INVALID; break; }
            const uint32_t *o0 = order0 + ((size_t)r * Lc + c) * Mr;
            const uint32_t *o1 = order1 + ((size_t)r * Lc + c) * Mc;
            for (int j = 0; j < Mc; ++j) {
                const int p = (int)o1[j];
                if (p < 1 || p >= N) { lsb::set_error("order1 entry %d outside [1, %d)", p, N); rc = LSB_ERR_INVALID; break; }
                const int m = makhoul(p);
                t_abs[j] = (uint16_t)(2 * swz(m >> 1) + (m & 1));
                t_azg[j] = (uint16_t)(2 * swz(brev_host(m >> 1, h)) + (m & 1));
            }
            for (int i = 0; i < Mr && !rc; ++i) {
                const int K = (int)o0[i];
                if (K < 1 || K >= N) { lsb::set_error("order0 entry %d outside [1, %d)", K, N); rc = LSB_ERR_INVALID; break; }
                t_abg[i] = (uint16_t)slot_of_freq(K, true);
                t_azs[i] = (uint16_t)slot_of_freq(K, false);
            }
            if (rc) break;
            BlockDev b;
            b.r = r;
            b.c = c;
            b.scale_d = std::sqrt(2.0 * wrc / L);
            b.scale = (float)b.scale_d;
            if ((rc = dev_upload(g, &b.ab_scatter, t_abs))) break;
            if ((rc = dev_upload(g, &b.ab_gather, t_abg))) break;
            if ((rc = dev_upload(g, &b.az_scatter, t_azs))) break;
            if ((rc = dev_upload(g, &b.az_gather, t_azg))) break;
            blocks.push_back(b);
            if (want_pair) {
                PairBlock pb{};
                pb.r = r;
                pb.c = c;
                pb.scale = b.scale;
                if ((rc = build_pair_block(g, h, Mr, Mc, o0, o1, &pb))) break;
                pair_blocks.push_back(pb);
            }
        }
    }
    col_first[Lc] = (int)blocks.size();
    if (!rc && blocks.empty()) { lsb::set_error("base matrix has no non-zero entry"); rc = LSB_ERR_INVALID; }
    DesignDev &d = g->dev;
    d.L = L; d.M = M; d.n = n; d.Lr = Lr; d.Lc = Lc; d.Mr = Mr; d.Mc = Mc; d.w = w; d.h = h;
    d.n_blocks = (int)blocks.size();
    d.logM = 0;
    while ((1 << d.logM) < M) ++d.logM;
    if (!rc) rc = dev_upload(g, &d.blocks, blocks);
    if (!rc) rc = dev_upload(g, &d.col_first, col_first);
    if (!rc) rc = dev_upload(g, &d.W, std::vector<double>(W, W + (size_t)Lr * Lc));
    if (!rc) rc = make_tables<float, float2>(g, &g->tb32, h);
    if (!rc) rc = make_tables<double, double2>(g, &g->tb64, h);
    if (!rc && want_pair) {
        rc = dev_upload(g, &g->pair.blocks, pair_blocks);
        if (!rc) rc = build_pair_tables(g, h);
        g->has_pair = !rc;
        for (int c = 0; c < Lc; ++c)
            if (col_first[c] == col_first[c + 1]) g->full_columns = false;
    }
    if (!rc) {
        cudaError_t e = cudaMalloc((void **)&g->d_queue, kCounters * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(g->d_queue, 0, kCounters * sizeof(unsigned long long));
        if (e != cudaSuccess) { lsb::set_error("queue alloc: %s", cudaGetErrorString(e)); rc = LSB_ERR_CUDA; }
    }
    if (rc) {
        lsb_sparc_design_destroy(g);
        return rc;
    }
    *out = g;
    return LSB_OK;
}

void lsb_sparc_design_destroy(lsb_sparc_design *g)
{
    if (!g) return;
    cudaDeviceSynchronize();
    for (void *p : g->allocs) cudaFree(p);
    cudaFree(g->d_queue);
    cudaFree(g->d_scratch);
    cudaFree(g->stage);
    if (g->stage_stream) cudaStreamDestroy(g->stage_stream);
    if (g->scratch_event) cudaEventDestroy(g->scratch_event);
    delete g;
}

int lsb_sparc_design_info(const lsb_sparc_design *g, int64_t info[8])
{
    LSB_REQUIRE(g && info, "null pointer");
    const DesignDev &d = g->dev;
    info[0] = d.L; info[1] = d.M; info[2] = d.n; info[3] = d.Lr; info[4] = d.Lc; info[5] = d.w;
    info[6] = d.n_blocks; info[7] = (int64_t)amp_smem(d, 4, needs_gbuf(d, 4));
    return LSB_OK;
}

int lsb_sparc_ab_dev_f32(const lsb_sparc_design *g, const float *x, float *out, int64_t n_frames, void *stream)
{
    return run_op<float>(g, false, x, out, n_frames, (cudaStream_t)stream);
}
int lsb_sparc_ab_dev_f64(const lsb_sparc_design *g, const double *x, double *out, int64_t n_frames, void *stream)
{
    return run_op<double>(g, false, x, out, n_frames, (cudaStream_t)stream);
}
int lsb_sparc_az_dev_f32(const lsb_sparc_design *g, const float *y, float *out, int64_t n_frames, void *stream)
{
    return run_op<float>(g, true, y, out, n_frames, (cudaStream_t)stream);
}
int lsb_sparc_az_dev_f64(const lsb_sparc_design *g, const double *y, double *out, int64_t n_frames, void *stream)
{
    return run_op<double>(g, true, y, out, n_frames, (cudaStream_t)stream);
}

int lsb_sparc_amp_dev_f32(lsb_sparc_design *g, const float *y, int64_t n_frames, const lsb_sparc_amp_params *p,
                          const int32_t *idx0, const lsb_sparc_amp_out *out, void *stream)
{
    return run_amp<float>(g, y, n_frames, p, idx0, out, (cudaStream_t)stream);
}
int lsb_sparc_amp_dev_f64(lsb_sparc_design *g, const double *y, int64_t n_frames, const lsb_sparc_amp_params *p,
                          const int32_t *idx0, const lsb_sparc_amp_out *out, void *stream)
{
    return run_amp<double>(g, y, n_frames, p, idx0, out, (cudaStream_t)stream);
}

// Host-pointer decode of a batch (what a loop over the reference's sparc_decode, sparc.py:55-74, consumes and
// returns): y in, section decisions + iteration counts + final psi out.  The copies are a few per cent of the
// decode (13.6 KB in and 4 KB out per frame against ~17 us of kernel time at config 1), so a batch goes up in
// one piece per 65 536 frames rather than in small overlapped chunks, which would leave a partial last wave
// of CTAs per chunk.
int lsb_sparc_amp_host_f32(lsb_sparc_design *g, const float *y, int64_t n_frames, const lsb_sparc_amp_params *p,
                           int32_t *idx, int32_t *t_final, double *psi)
{
    if (n_frames == 0) return LSB_OK;
    LSB_REQUIRE(g && y && p && idx && t_final, "null pointer");
    LSB_REQUIRE(n_frames >= 0, "negative frame count");
    int rc = check_device(g);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g->stage_mu);
    const DesignDev &d = g->dev;
    const int64_t chunk = n_frames < 65536 ? n_frames : 65536;
    auto pad = [](size_t b) { return (b + 255) & ~(size_t)255; };       // every staging segment starts 256-byte aligned
    const size_t y_b = pad((size_t)chunk * d.n * sizeof(float)), idx_b = pad((size_t)chunk * d.L * sizeof(int32_t));
    const size_t t_b = pad((size_t)chunk * sizeof(int32_t)), psi_b = pad((size_t)chunk * d.Lc * sizeof(double));
    const size_t need = y_b + idx_b + t_b + psi_b;
    if (!g->stage_stream) LSB_CUDA(cudaStreamCreateWithFlags(&g->stage_stream, cudaStreamNonBlocking));
    if (g->stage_bytes < need) {
        LSB_CUDA(cudaStreamSynchronize(g->stage_stream));
        cudaFree(g->stage);
        g->stage = nullptr;
        g->stage_bytes = 0;
        LSB_CUDA(cudaMalloc(&g->stage, need));
        g->stage_bytes = need;
    }
    char *base = (char *)g->stage;
    float *d_y = (float *)base;
    double *d_psi = (double *)(base + y_b);
    int32_t *d_idx = (int32_t *)(base + y_b + psi_b);
    int32_t *d_t = (int32_t *)(base + y_b + psi_b + idx_b);
    const cudaStream_t st = g->stage_stream;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk) {
        const int64_t nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        LSB_CUDA(cudaMemcpyAsync(d_y, y + (size_t)f0 * d.n, (size_t)nf * d.n * sizeof(float), cudaMemcpyHostToDevice, st));
        lsb_sparc_amp_out o{};
        o.idx = d_idx;
        o.t_final = d_t;
        o.psi = d_psi;
        if ((rc = run_amp<float>(g, d_y, nf, p, nullptr, &o, st))) return rc;
        LSB_CUDA(cudaMemcpyAsync(idx + (size_t)f0 * d.L, d_idx, (size_t)nf * d.L * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        LSB_CUDA(cudaMemcpyAsync(t_final + f0, d_t, (size_t)nf * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (psi)
            LSB_CUDA(cudaMemcpyAsync(psi + (size_t)f0 * d.Lc, d_psi, (size_t)nf * d.Lc * sizeof(double), cudaMemcpyDeviceToHost, st));
        LSB_CUDA(cudaStreamSynchronize(st));
    }
    return LSB_OK;
}

}  // extern "C"

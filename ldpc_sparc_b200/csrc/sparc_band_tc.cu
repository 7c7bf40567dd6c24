// Tensor-core form of the sparc_public AMP decoder: the sub-sampled DCT design as an explicit block-banded matrix.
//
// Reference: sparc_public/sparc.py -- sparc_amp :883-999 with the operators of sparc_transforms :703-880.  Block
// (r, c) of the design is  A_rc[i, j] = sqrt(2 W_rc / L) cos(pi o0[i] (2 o1[j] + 1) / (2 w))  (sub_dct :648-701 with
// the orderings of generate_ordering :735-775).  The FFT path (sparc_amp.cu) evaluates A_rc x with one length-w
// transform per block; that is the right algorithm while a block keeps a good part of the w outputs.  A spatially
// coupled design keeps very few: BASELINE config 4 (sparc_demo_sc_decode_wave) has Mr = 332 rows per block out of
// w = 65536, 192 non-zero blocks, so one iteration is 384 transforms of 65536 points (~1 GFLOP of butterflies
// through a 256 KB buffer that does not fit shared memory) to produce what a 332 x 32768 matrix block produces
// directly.  Stored explicitly the whole design is 192 x 332 x 32768 entries = 4.2 GB per fp16 array, which a B200
// holds many times over, and A beta / A^T z over a frame batch become block-banded GEMMs on the tcgen05 pipeline
// of tc_gemm.cuh (split-fp16 operands, fp32 accumulation in TMEM): 8.4 GFLOP per frame-iteration, 25 issued.
//
// Layout.  Arow [Lr * Mr_pad, kmax] holds, for row block r, the blocks (r, clo_r .. chi_r) side by side (K pitch
// Mc_pad per block, zero padded), rows padded to whole 128-row tiles; Acol [L M, kz] holds, for column block c, the
// blocks (rlo_c .. rhi_c, c) transposed, side by side along K (= the rows of z they multiply, which are contiguous).
// A work item (TcItem) is one 128 x 128 output tile of one block (A beta: partial p = c - clo_r, summed in index
// order by the residual kernel) or of one column block (A^T z: complete).  Everything else of the iteration is two
// memory-bound kernels (band_residual_kernel, band_denoise_kernel) that also produce the fp16 operand pairs.
#include "tc_gemm.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <vector>

namespace {

constexpr int kBandThreads = 256;
constexpr int kMaxLr = 64, kMaxLc = 64;

struct BandDev {
    int L, M, n, Lr, Lc, Mr, Mc;
    int Mr_p, n_pitch;                        // fp16 pair of z: row block r starts at r * Mr_p (Mr rounded up to 8: TMA
                                              // needs 16-byte aligned box starts), row pitch n_pitch = Lr * Mr_p
    const double *W;                          // [Lr, Lc]
    const int *clo, *chi;                     // [Lr] first / last non-zero column block of row block r
};

struct BandParams {
    double awgn_var, rtol, atol;
    int t_max, phi_method;
};

// per-call state and outputs (device pointers)
struct BandState {
    float *beta, *atz, *z, *part;
    tc_t *b_hi, *b_lo, *z_hi, *z_lo;
    double *psi, *psi_prev, *phi, *tau;
    int *done, *n_done;
    long long part_stride;                    // elements between two partials of A beta
    // outputs
    int32_t *idx, *t_final;
    float *nmse, *s_last;
    double *tau_trace, *psi_trace;
    const int32_t *idx0;
};

__device__ __forceinline__ double block_sum_d(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kBandThreads / 32; ++w) tot += red[w];
    __syncthreads();
    return tot;
}

// ---------------------------------------------------------------- operand generation (once per design)
// value of design entry (r, c, i, j) in double: scale_rc cos(pi (o0 (2 o1 + 1) mod 4w) / 2w); the phase is reduced
// exactly in integers (4 w is a power of two)
__device__ __forceinline__ double design_entry(double scale, unsigned o0, unsigned o1, unsigned w)
{
    const unsigned k = (o0 * (2u * o1 + 1u)) & (4u * w - 1u);
    return scale * cospi((double)k / (double)(2u * w));
}

__device__ __forceinline__ void store_split(tc_t *hi, tc_t *lo, size_t e, double v)
{
    const float x = (float)v;
    const tc_t h = to_tc(x);
    hi[e] = h;
    lo[e] = to_tc((float)(v - (double)from_tc(h)));
}

__global__ void gen_rows_kernel(BandDev d, const uint32_t *__restrict__ order0, const uint32_t *__restrict__ order1, unsigned w,
                                int Mr_pad, int Mc_pad, long long kmax, tc_t *hi, tc_t *lo)
{
    const long long total = (long long)d.Lr * Mr_pad * kmax;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / kmax;
        const int k = (int)(e - row * kmax);
        const int r = (int)(row / Mr_pad), i = (int)(row - (long long)r * Mr_pad);
        const int cc = k / Mc_pad, j = k - cc * Mc_pad, c = d.clo[r] + cc;
        double v = 0.0;
        if (i < d.Mr && j < d.Mc && c <= d.chi[r]) {
            const double wrc = d.W[r * d.Lc + c];
            if (wrc != 0.0) {
                const size_t b = (size_t)r * d.Lc + c;
                v = design_entry(sqrt(2.0 * wrc / d.L), order0[b * d.Mr + i], order1[b * d.Mc + j], w);
            }
        }
        store_split(hi, lo, (size_t)e, v);
    }
}

__global__ void gen_cols_kernel(BandDev d, const int *__restrict__ rlo, const int *__restrict__ rhi, const uint32_t *__restrict__ order0,
                                const uint32_t *__restrict__ order1, unsigned w, long long kz, tc_t *hi, tc_t *lo)
{
    const long long total = (long long)d.L * d.M * kz;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / kz;
        const int k = (int)(e - row * kz);
        const int c = (int)(row / d.Mc), j = (int)(row - (long long)c * d.Mc);
        const int rr = k / d.Mr_p, i = k - rr * d.Mr_p, r = rlo[c] + rr;
        double v = 0.0;
        if (r <= rhi[c] && i < d.Mr) {
            const double wrc = d.W[r * d.Lc + c];
            if (wrc != 0.0) {
                const size_t b = (size_t)r * d.Lc + c;
                v = design_entry(sqrt(2.0 * wrc / d.L), order0[b * d.Mr + i], order1[b * d.Mc + j], w);
            }
        }
        store_split(hi, lo, (size_t)e, v);
    }
}

// ---------------------------------------------------------------- plain operand split (operators Ab / Az)
// x [F, len] fp32 -> fp16 pair with row pitch `pitch`; entry i goes to (i / blk) * blk_p + i % blk
__global__ void band_split_kernel(const float *__restrict__ x, tc_t *__restrict__ hi, tc_t *__restrict__ lo,
                                  long long F, int len, int pitch, int blk, int blk_p)
{
    const long long total = F * (long long)len;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long f = e / len;
        const int i = (int)(e - f * len);
        const float v = x[e];
        const tc_t h = to_tc(v);
        const int o = (i / blk) * blk_p + i % blk;
        hi[f * pitch + o] = h;
        lo[f * pitch + o] = to_tc(v - from_tc(h));
    }
}

// out[f, i] = sum over the valid partials of row block r(i), in index order
__global__ void band_sum_partials_kernel(BandDev d, const float *__restrict__ part, long long part_stride, float *__restrict__ out,
                                         long long F)
{
    const long long total = F * (long long)d.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % d.n);
        const int r = i / d.Mr;
        float a = 0.0f;
        for (int c = d.clo[r]; c <= d.chi[r]; ++c)
            if (d.W[r * d.Lc + c] != 0.0) a += part[(long long)(c - d.clo[r]) * part_stride + e];
        out[e] = a;
    }
}

// ---------------------------------------------------------------- AMP iteration: residual side
// One CTA per frame, iteration t (sparc.py:931-969):
//   [t >= 2: early-stop test of iteration t - 1 (:984-986); a stopped frame is frozen from here on]
//   gamma_r = sum_c W_rc psi_c / Lc;  t > 0: z_r <- y_r - (A beta)_r + (gamma_r / phi_prev_r) z_r   (t = 0: z = y)
//   phi_r = awgn_var + gamma_r | mean(z_r^2);   tau_c = (L / Mr) / sum_r W_rc / phi_r
//   operand of A^T: z_r / phi_r as a fp16 pair
__global__ void __launch_bounds__(kBandThreads)
band_residual_kernel(BandDev d, BandParams p, BandState s, const float *__restrict__ y_all, long long F, int t)
{
    __shared__ double gam[kMaxLr], phi[kMaxLr], coef[kMaxLr], zsq[kMaxLr], red[kBandThreads / 32];
    __shared__ int s_stop;
    const long long f = blockIdx.x;
    const int tid = threadIdx.x;
    if (s.done[f]) return;
    double *psi = s.psi + f * d.Lc, *psi_prev = s.psi_prev + f * d.Lc, *phi_g = s.phi + f * d.Lr, *tau_g = s.tau + f * d.Lc;
    if (tid == 0) {
        int stop = 0;
        if (t >= 2) {                                          // iteration t - 1 >= 1 has run: compare its psi with the one before
            stop = 1;
            for (int c = 0; c < d.Lc; ++c)
                if (!(fabs(psi[c] - psi_prev[c]) <= p.atol + p.rtol * fabs(psi_prev[c]))) stop = 0;
            if (stop) {
                s.done[f] = 1;
                s.t_final[f] = t;                              // stopped in iteration t - 1: t_final = (t - 1) + 1
                atomicAdd(s.n_done, 1);
                if (s.nmse)                                    // nmse[ts:] = nmse[ts], ts = t - 1  (:985)
                    for (int tt = t; tt < p.t_max; ++tt)
                        for (int c = 0; c < d.Lc; ++c)
                            s.nmse[(f * p.t_max + tt) * d.Lc + c] = s.nmse[(f * p.t_max + t - 1) * d.Lc + c];
            }
        }
        s_stop = stop;
    }
    __syncthreads();
    if (s_stop) return;
    if (tid < d.Lr) {
        double g = 0.0;
        for (int c = 0; c < d.Lc; ++c) g += d.W[tid * d.Lc + c] * (t == 0 ? 1.0 : psi[c]);
        gam[tid] = g / d.Lc;
        coef[tid] = t > 0 ? gam[tid] / phi_g[tid] : 0.0;
        zsq[tid] = 0.0;
    }
    __syncthreads();
    if (tid < d.Lc && t > 0) psi_prev[tid] = psi[tid];
    const float *y = y_all + f * d.n;
    float *z = s.z + f * d.n;
    if (t == 0) {
        for (int i = tid; i < d.n; i += kBandThreads) z[i] = y[i];
    } else {
        const long long e0 = f * d.n;
        for (int i = tid; i < d.n; i += kBandThreads) {
            const int r = i / d.Mr;
            float ab = 0.0f;
            for (int c = d.clo[r]; c <= d.chi[r]; ++c)
                if (d.W[r * d.Lc + c] != 0.0) ab += s.part[(long long)(c - d.clo[r]) * s.part_stride + e0 + i];
            z[i] = (y[i] - ab) + (float)coef[r] * z[i];
        }
    }
    if (p.phi_method == 2) {
        __syncthreads();
        for (int r = 0; r < d.Lr; ++r) {
            double part = 0.0;
            for (int i = tid; i < d.Mr; i += kBandThreads) { const double v = (double)z[r * d.Mr + i]; part += v * v; }
            const double tot = block_sum_d(part, red);
            if (tid == 0) zsq[r] = tot;
        }
    }
    __syncthreads();
    if (tid < d.Lr) {
        phi[tid] = p.phi_method == 1 ? p.awgn_var + gam[tid] : zsq[tid] / d.Mr;
        phi_g[tid] = phi[tid];
    }
    __syncthreads();
    if (tid < d.Lc) {
        double tc;
        if (d.Lr == 1) {
            tc = ((double)d.L * phi[0] / d.n) / d.W[tid];
        } else {
            double acc_w = 0.0;
            for (int r = 0; r < d.Lr; ++r) acc_w += d.W[r * d.Lc + tid] / phi[r];
            tc = ((double)d.L / d.Mr) / acc_w;
        }
        tau_g[tid] = tc;
        if (s.tau_trace) s.tau_trace[(f * p.t_max + t) * d.Lc + tid] = tc;
    }
    tc_t *zh = s.z_hi + f * d.n_pitch, *zl = s.z_lo + f * d.n_pitch;
    for (int i = tid; i < d.n; i += kBandThreads) {
        const int r = i / d.Mr, o = r * d.Mr_p + (i - r * d.Mr);
        const float v = z[i] * (float)(1.0 / phi[r]);
        const tc_t h = to_tc(v);
        zh[o] = h;
        zl[o] = to_tc(v - from_tc(h));
    }
}

// ---------------------------------------------------------------- AMP iteration: denoiser side
// One CTA per (column block, frame), one warp per section (M >= 32: PL = M / 32 entries per lane; M < 32: 32 / M
// sections per warp pass):  s = beta + tau_c (A^T z)_c  (:972),  beta = softmax(s / tau_c)  (:402-432, per-section max
// shift), psi_c = sum beta (1 - beta) / (L / Lc)  (:976-978), nmse (:979-981), the MAP decision argmax s (:997, kept
// from the last iteration a frame runs), beta as fp32 and as the fp16 pair of the next A beta.
template <int PL>
__global__ void __launch_bounds__(kBandThreads)
band_denoise_kernel(BandDev d, BandParams p, BandState s, int t)
{
    __shared__ double red[kBandThreads / 32];
    const long long f = blockIdx.y;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (s.done[f]) return;
    const float tau_c = (float)s.tau[f * d.Lc + c], inv_tau = 1.0f / tau_c;
    const int secs = d.Mc / d.M, sec0 = c * secs;
    const size_t base_c = (size_t)f * (size_t)d.L * d.M + (size_t)c * d.Mc;
    const bool has_beta = t > 0;
    double psi_part = 0.0, nm_part = 0.0;
    if (PL > 0) {
        // a lane owns PL CONSECUTIVE entries of the section (16-byte loads / stores for PL >= 4, fp16 pairs packed)
        constexpr int P_ = PL > 0 ? PL : 1;
        for (int sec = warp; sec < secs; sec += kBandThreads / 32) {
            const size_t a0 = base_c + (size_t)sec * d.M + (size_t)lane * P_;
            float sv[P_];
            if (P_ % 4 == 0) {
#pragma unroll
                for (int k = 0; k < P_; k += 4) {
                    const float4 az4 = *reinterpret_cast<const float4 *>(s.atz + a0 + k);
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_beta) b4 = *reinterpret_cast<const float4 *>(s.beta + a0 + k);
                    sv[k] = b4.x + tau_c * az4.x;
                    sv[k + 1] = b4.y + tau_c * az4.y;
                    sv[k + 2] = b4.z + tau_c * az4.z;
                    sv[k + 3] = b4.w + tau_c * az4.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < P_; ++k) sv[k] = (has_beta ? s.beta[a0 + k] : 0.0f) + tau_c * s.atz[a0 + k];
            }
            float mx = -INFINITY;
            int arg = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < P_; ++k) {
                if (s.s_last) s.s_last[a0 + k] = sv[k];
                if (sv[k] > mx) { mx = sv[k]; arg = lane * P_ + k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, mx, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
            }
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < P_; ++k) {
                sv[k] = expf((sv[k] - mx) * inv_tau);
                sum += sv[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.0f / sum;
            const int truth = s.idx0 ? s.idx0[f * d.L + sec0 + sec] : -1;
#pragma unroll
            for (int k = 0; k < P_; ++k) {
                const float e = sv[k];
                sv[k] = e * inv;
                psi_part += (double)sv[k] * (double)((sum - e) * inv);
                const double dlt = (double)sv[k] - ((lane * P_ + k) == truth ? 1.0 : 0.0);
                nm_part += dlt * dlt;
            }
            if (P_ % 8 == 0) {
#pragma unroll
                for (int k = 0; k < P_; k += 8) {
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float v0 = sv[k + 2 * u], v1 = sv[k + 2 * u + 1];
                        const tc_t h0 = to_tc(v0), h1 = to_tc(v1);
                        h[u] = pack_pair(h0, h1);
                        l[u] = pack_pair(to_tc(v0 - from_tc(h0)), to_tc(v1 - from_tc(h1)));
                    }
                    *reinterpret_cast<float4 *>(s.beta + a0 + k) = make_float4(sv[k], sv[k + 1], sv[k + 2], sv[k + 3]);
                    *reinterpret_cast<float4 *>(s.beta + a0 + k + 4) = make_float4(sv[k + 4], sv[k + 5], sv[k + 6], sv[k + 7]);
                    *reinterpret_cast<uint4 *>(s.b_hi + a0 + k) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4 *>(s.b_lo + a0 + k) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < P_; ++k) {
                    s.beta[a0 + k] = sv[k];
                    const tc_t h = to_tc(sv[k]);
                    s.b_hi[a0 + k] = h;
                    s.b_lo[a0 + k] = to_tc(sv[k] - from_tc(h));
                }
            }
            if (lane == 0) s.idx[f * d.L + sec0 + sec] = arg;
        }
    } else {
        const int spw = 32 / d.M;                               // sections per warp pass
        for (int g0 = warp * spw; g0 < secs; g0 += (kBandThreads / 32) * spw) {
            const int sec = g0 + lane / d.M, j = lane % d.M;
            const bool live = sec < secs;
            const size_t a = base_c + (size_t)sec * d.M + j;
            float v = -INFINITY;
            if (live) {
                v = (has_beta ? s.beta[a] : 0.0f) + tau_c * s.atz[a];
                if (s.s_last) s.s_last[a] = v;
            }
            float mx = v;
            int arg = j;
            for (int o = d.M >> 1; o > 0; o >>= 1) {
                const float om = __shfl_xor_sync(0xffffffffu, mx, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
            }
            const float e = live ? expf((v - mx) * inv_tau) : 0.0f;
            float sum = e;
            for (int o = d.M >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (live) {
                const float inv = 1.0f / sum, b = e * inv;
                s.beta[a] = b;
                const tc_t h = to_tc(b);
                s.b_hi[a] = h;
                s.b_lo[a] = to_tc(b - from_tc(h));
                psi_part += (double)b * (double)((sum - e) * inv);
                const int truth = s.idx0 ? s.idx0[f * d.L + sec0 + sec] : -1;
                const double dlt = (double)b - (j == truth ? 1.0 : 0.0);
                nm_part += dlt * dlt;
                if (j == 0) s.idx[f * d.L + sec0 + sec] = arg;
            }
        }
    }
    const double sec_norm = (double)d.L / d.Lc;
    const double psi_tot = block_sum_d(psi_part, red);
    const double nm_tot = block_sum_d(nm_part, red);
    if (tid == 0) {
        const double ps = psi_tot / sec_norm;
        s.psi[f * d.Lc + c] = ps;
        if (s.psi_trace) s.psi_trace[(f * p.t_max + t) * d.Lc + c] = ps;
        if (s.nmse) s.nmse[(f * p.t_max + t + 1) * d.Lc + c] = (float)(nm_tot / sec_norm);
    }
}

// after the last iteration: frames that never stopped (or stopped in the very last iteration) get t_final = t_max - 1
__global__ void band_finish_kernel(BandDev d, BandParams p, BandState s, long long F, int t_last /* last iteration run */)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F || s.done[f]) return;
    s.t_final[f] = t_last + 1;
    if (s.nmse && t_last >= 1) {
        const double *psi = s.psi + f * d.Lc, *psi_prev = s.psi_prev + f * d.Lc;
        int stop = 1;
        for (int c = 0; c < d.Lc; ++c)
            if (!(fabs(psi[c] - psi_prev[c]) <= p.atol + p.rtol * fabs(psi_prev[c]))) stop = 0;
        if (stop)
            for (int tt = t_last + 1; tt < p.t_max; ++tt)
                for (int c = 0; c < d.Lc; ++c)
                    s.nmse[(f * p.t_max + tt) * d.Lc + c] = s.nmse[(f * p.t_max + t_last) * d.Lc + c];
    }
}

}  // namespace

struct lsb_band_design {
    int device = 0;
    BandDev dev{};
    int Mr_pad = 0, Mc_pad = 0, max_cols = 0;
    long long kmax = 0, kz = 0;
    std::vector<double> W;
    std::vector<int> clo, chi, rlo, rhi;
    std::vector<void *> allocs;
    tc_t *row_hi = nullptr, *row_lo = nullptr, *col_hi = nullptr, *col_lo = nullptr;
    CUtensorMap map_row_hi, map_row_lo, map_col_hi, map_col_lo;
    std::mutex mu;
    // per (stream) work state, sized for the largest batch seen
    struct Work {
        long long frames = 0;
        std::vector<void *> allocs;
        BandState st{};
        TcItem *ab_items = nullptr, *az_items = nullptr;
        int *ab_first = nullptr, *az_first = nullptr;
        int ab_grid = 0, az_grid = 0;
        CUtensorMap mb_hi, mb_lo, mz_hi, mz_lo;
    };
    std::map<cudaStream_t, Work> work;
    // host-pointer entry: its own stream and staging buffer
    std::mutex host_mu;
    cudaStream_t host_stream = nullptr;
    void *host_stage = nullptr;
    size_t host_stage_bytes = 0;
};

namespace {

template <typename T>
int band_alloc(std::vector<void *> &allocs, T **p, size_t count)
{
    void *q = nullptr;
    LSB_CUDA(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
    allocs.push_back(q);
    *p = (T *)q;
    return LSB_OK;
}

template <typename T>
int band_upload(std::vector<void *> &allocs, const T **dst, const std::vector<T> &src)
{
    T *p = nullptr;
    int rc = band_alloc(allocs, &p, src.size());
    if (rc) return rc;
    LSB_CUDA(cudaMemcpy(p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = p;
    return LSB_OK;
}

// Deals the items round-robin to `grid` CTAs (CTA b gets items b, b + grid, ...) and rewrites the list so that every
// CTA's share is contiguous.  Items are generated with the frame tile varying fastest, so the CTAs that run at the
// same time work on the same W tile for neighbouring frame tiles and share its K blocks through L2 -- a CTA that
// walked the frame tiles of one W tile by itself would stream the tile's whole K panel (16 MB for a config-4 block)
// from DRAM once per frame tile.  All items of one product have (nearly) the same K length, so the deal is balanced.
void deal_items(std::vector<TcItem> *items, int grid, std::vector<int> *first)
{
    std::vector<TcItem> out;
    out.reserve(items->size());
    first->assign(grid + 1, 0);
    for (int b = 0; b < grid; ++b) {
        (*first)[b] = (int)out.size();
        for (size_t i = b; i < items->size(); i += grid) out.push_back((*items)[i]);
    }
    (*first)[grid] = (int)out.size();
    items->swap(out);
}

int prepare_work(lsb_band_design *g, lsb_band_design::Work &wk, long long F)
{
    if (wk.frames >= F && wk.frames <= F + 127) return LSB_OK;       // item lists depend on the number of frame tiles
    for (void *p : wk.allocs) cudaFree(p);
    wk.allocs.clear();
    wk.frames = 0;
    const BandDev &d = g->dev;
    const size_t LM = (size_t)d.L * d.M;
    BandState &s = wk.st;
    int rc = LSB_OK;
    if (!rc) rc = band_alloc(wk.allocs, &s.beta, (size_t)F * LM);
    if (!rc) rc = band_alloc(wk.allocs, &s.atz, (size_t)F * LM);
    if (!rc) rc = band_alloc(wk.allocs, &s.b_hi, (size_t)F * LM);
    if (!rc) rc = band_alloc(wk.allocs, &s.b_lo, (size_t)F * LM);
    if (!rc) rc = band_alloc(wk.allocs, &s.z, (size_t)F * d.n);
    if (!rc) rc = band_alloc(wk.allocs, &s.z_hi, (size_t)F * d.n_pitch);
    if (!rc) rc = band_alloc(wk.allocs, &s.z_lo, (size_t)F * d.n_pitch);
    if (!rc) rc = band_alloc(wk.allocs, &s.part, (size_t)g->max_cols * (size_t)F * d.n);
    if (!rc) rc = band_alloc(wk.allocs, &s.psi, (size_t)F * d.Lc);
    if (!rc) rc = band_alloc(wk.allocs, &s.psi_prev, (size_t)F * d.Lc);
    if (!rc) rc = band_alloc(wk.allocs, &s.phi, (size_t)F * d.Lr);
    if (!rc) rc = band_alloc(wk.allocs, &s.tau, (size_t)F * d.Lc);
    if (!rc) rc = band_alloc(wk.allocs, &s.done, (size_t)F);
    if (!rc) rc = band_alloc(wk.allocs, &s.n_done, 1);
    if (rc) return rc;
    s.part_stride = (long long)F * d.n;
    // the padding columns of the z pair are read by TMA (multiplied by zero rows of W): keep them finite
    LSB_CUDA(cudaMemset(s.z_hi, 0, (size_t)F * d.n_pitch * sizeof(tc_t)));
    LSB_CUDA(cudaMemset(s.z_lo, 0, (size_t)F * d.n_pitch * sizeof(tc_t)));

    const int n_f_tiles = (int)((F + kTileN - 1) / kTileN), sms = lsb::sm_count();
    std::vector<TcItem> ab, az;
    // column block outermost: the row blocks that use column block c run together and share its beta tiles in L2
    // (row block outermost re-read every beta panel from DRAM once per row block of the band: 6x for config 4)
    for (int c = 0; c < d.Lc; ++c)
        for (int r = g->rlo[c]; r <= g->rhi[c]; ++r) {
            if (g->W[r * d.Lc + c] == 0.0) continue;
            const int cc = c - g->clo[r];
            for (int tt = 0; tt * kTileRows < d.Mr; ++tt)
                for (int ft = 0; ft < n_f_tiles; ++ft) {
                    TcItem im;
                    im.w_row0 = r * g->Mr_pad + tt * kTileRows;
                    im.w_k0 = cc * g->Mc_pad;
                    im.x_k0 = c * d.Mc;
                    im.f0 = ft * kTileN;
                    im.kb_count = g->Mc_pad / kBlockK;
                    im.rows_valid = std::min(kTileRows, d.Mr - tt * kTileRows);
                    im.out_off = (long long)cc * s.part_stride + (long long)r * d.Mr + tt * kTileRows;
                    ab.push_back(im);
                }
        }
    for (int c = 0; c < d.Lc; ++c) {
        const int kb = (int)(((long long)(g->rhi[c] - g->rlo[c] + 1) * d.Mr_p + kBlockK - 1) / kBlockK);
        for (int tt = 0; tt * kTileRows < d.Mc; ++tt)
            for (int ft = 0; ft < n_f_tiles; ++ft) {
                TcItem im;
                im.w_row0 = c * d.Mc + tt * kTileRows;
                im.w_k0 = 0;
                im.x_k0 = g->rlo[c] * d.Mr_p;
                im.f0 = ft * kTileN;
                im.kb_count = kb;
                im.rows_valid = std::min(kTileRows, d.Mc - tt * kTileRows);
                im.out_off = (long long)c * d.Mc + tt * kTileRows;
                az.push_back(im);
            }
    }
    wk.ab_grid = std::min<int>(sms, (int)ab.size());
    wk.az_grid = std::min<int>(sms, (int)az.size());
    std::vector<int> ab_first, az_first;
    deal_items(&ab, wk.ab_grid, &ab_first);
    deal_items(&az, wk.az_grid, &az_first);
    const TcItem *p_items = nullptr;
    const int *p_first = nullptr;
    if ((rc = band_upload(wk.allocs, &p_items, ab))) return rc;
    wk.ab_items = const_cast<TcItem *>(p_items);
    if ((rc = band_upload(wk.allocs, &p_items, az))) return rc;
    wk.az_items = const_cast<TcItem *>(p_items);
    if ((rc = band_upload(wk.allocs, &p_first, ab_first))) return rc;
    wk.ab_first = const_cast<int *>(p_first);
    if ((rc = band_upload(wk.allocs, &p_first, az_first))) return rc;
    wk.az_first = const_cast<int *>(p_first);
    if ((rc = make_map(&wk.mb_hi, s.b_hi, (uint64_t)F, LM))) return rc;
    if ((rc = make_map(&wk.mb_lo, s.b_lo, (uint64_t)F, LM))) return rc;
    if ((rc = make_map(&wk.mz_hi, s.z_hi, (uint64_t)F, (uint64_t)d.n_pitch))) return rc;
    if ((rc = make_map(&wk.mz_lo, s.z_lo, (uint64_t)F, (uint64_t)d.n_pitch))) return rc;
    wk.frames = F;
    return LSB_OK;
}

int get_work(lsb_band_design *g, cudaStream_t st, long long F, lsb_band_design::Work **out)
{
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    std::unique_lock<std::mutex> lock(g->mu);
    LSB_REQUIRE(g->work.size() < 8 || g->work.count(st), "more than 8 streams use this design");
    lsb_band_design::Work &wk = g->work[st];
    lock.unlock();
    if (!(wk.frames >= F && wk.frames <= F + 127)) LSB_CUDA(cudaStreamSynchronize(st));
    int rc = prepare_work(g, wk, F);
    if (rc) return rc;
    *out = &wk;
    return LSB_OK;
}

void launch_ab(lsb_band_design *g, lsb_band_design::Work &wk, long long F, cudaStream_t st)
{
    tc_items_kernel<<<wk.ab_grid, kThreadsTc, kSmemBytes, st>>>(g->map_row_hi, g->map_row_lo, wk.mb_hi, wk.mb_lo, wk.st.part,
                                                               (long long)g->dev.n, (int)F, wk.ab_items, wk.ab_first);
    lsb::count_launch();
}

void launch_az(lsb_band_design *g, lsb_band_design::Work &wk, long long F, cudaStream_t st)
{
    tc_items_kernel<<<wk.az_grid, kThreadsTc, kSmemBytes, st>>>(g->map_col_hi, g->map_col_lo, wk.mz_hi, wk.mz_lo, wk.st.atz,
                                                               (long long)g->dev.L * g->dev.M, (int)F, wk.az_items, wk.az_first);
    lsb::count_launch();
}

template <int PL>
void launch_denoise(const BandDev &d, const BandParams &p, const BandState &s, long long F, int t, cudaStream_t st)
{
    band_denoise_kernel<PL><<<dim3((unsigned)d.Lc, (unsigned)F), kBandThreads, 0, st>>>(d, p, s, t);
}

}  // namespace

extern "C" {

int lsb_band_design_create(const double *W, int Lr, int Lc, int L, int M, int n, const uint32_t *order0, const uint32_t *order1,
                           lsb_band_design **out)
{
    LSB_REQUIRE(W && order0 && order1 && out, "null pointer");
    LSB_REQUIRE(Lr >= 1 && Lr <= kMaxLr && Lc >= 1 && Lc <= kMaxLc, "base matrix %dx%d outside [1,%d]x[1,%d]", Lr, Lc, kMaxLr,
                kMaxLc);
    LSB_REQUIRE(L > 0 && M >= 2 && M <= 1024 && (M & (M - 1)) == 0 && n > 0, "bad L / M / n (M: power of two in [2, 1024])");
    LSB_REQUIRE(((long long)L * M) % Lc == 0 && n % Lr == 0 && L % Lc == 0, "Lc must divide L and LM, Lr must divide n");
    const int Mr = n / Lr, Mc = (int)((long long)L * M / Lc);
    LSB_REQUIRE(Mc % M == 0, "column blocks must hold whole sections");
    LSB_REQUIRE(Mc % 8 == 0, "column blocks (L M / Lc = %d entries) must be a multiple of 8 long: TMA boxes start on 16 bytes", Mc);
    unsigned w = 1;
    while ((int)w < (Mr > Mc ? Mr : Mc) + 1) w <<= 1;
    LSB_REQUIRE(w <= 65536, "transform size %u above 65536", w);

    lsb_band_design *g = new (std::nothrow) lsb_band_design();
    if (!g) return LSB_ERR_NOMEM;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device");
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    g->W.assign(W, W + (size_t)Lr * Lc);
    g->clo.assign(Lr, Lc);
    g->chi.assign(Lr, -1);
    g->rlo.assign(Lc, Lr);
    g->rhi.assign(Lc, -1);
    int rc = LSB_OK;
    for (int r = 0; r < Lr; ++r)
        for (int c = 0; c < Lc; ++c) {
            if (W[r * Lc + c] < 0.0) { lsb::set_error("negative base-matrix entry"); rc = LSB_ERR_INVALID; }
            if (W[r * Lc + c] == 0.0) continue;
            g->clo[r] = std::min(g->clo[r], c);
            g->chi[r] = std::max(g->chi[r], c);
            g->rlo[c] = std::min(g->rlo[c], r);
            g->rhi[c] = std::max(g->rhi[c], r);
        }
    int max_cols = 0, max_rows = 0;
    for (int r = 0; r < Lr && !rc; ++r) {
        if (g->chi[r] < 0) { lsb::set_error("row block %d of the base matrix is empty", r); rc = LSB_ERR_INVALID; }
        max_cols = std::max(max_cols, g->chi[r] - g->clo[r] + 1);
    }
    for (int c = 0; c < Lc && !rc; ++c) {
        if (g->rhi[c] < 0) { lsb::set_error("column block %d of the base matrix is empty", c); rc = LSB_ERR_INVALID; }
        max_rows = std::max(max_rows, g->rhi[c] - g->rlo[c] + 1);
    }
    BandDev &d = g->dev;
    d.L = L; d.M = M; d.n = n; d.Lr = Lr; d.Lc = Lc; d.Mr = Mr; d.Mc = Mc;
    d.Mr_p = (Mr + 7) / 8 * 8;
    d.n_pitch = Lr * d.Mr_p;
    g->Mr_pad = (Mr + kTileRows - 1) / kTileRows * kTileRows;
    g->Mc_pad = (Mc + kBlockK - 1) / kBlockK * kBlockK;
    g->max_cols = max_cols;
    g->kmax = (long long)max_cols * g->Mc_pad;
    g->kz = (long long)max_rows * d.Mr_p;
    const uint32_t *d_o0 = nullptr, *d_o1 = nullptr;
    const int *d_rlo = nullptr, *d_rhi = nullptr;
    std::vector<void *> tmp;
    if (!rc) rc = band_upload(g->allocs, &d.W, g->W);
    if (!rc) rc = band_upload(g->allocs, &d.clo, g->clo);
    if (!rc) rc = band_upload(g->allocs, &d.chi, g->chi);
    if (!rc) rc = band_upload(tmp, &d_rlo, g->rlo);
    if (!rc) rc = band_upload(tmp, &d_rhi, g->rhi);
    if (!rc) rc = band_upload(tmp, &d_o0, std::vector<uint32_t>(order0, order0 + (size_t)Lr * Lc * Mr));
    if (!rc) rc = band_upload(tmp, &d_o1, std::vector<uint32_t>(order1, order1 + (size_t)Lr * Lc * Mc));
    const size_t n_row = (size_t)Lr * g->Mr_pad * (size_t)g->kmax, n_col = (size_t)L * M * (size_t)g->kz;
    if (!rc) rc = band_alloc(g->allocs, &g->row_hi, n_row);
    if (!rc) rc = band_alloc(g->allocs, &g->row_lo, n_row);
    if (!rc) rc = band_alloc(g->allocs, &g->col_hi, n_col);
    if (!rc) rc = band_alloc(g->allocs, &g->col_lo, n_col);
    if (!rc) {
        const int grid = lsb::sm_count() * 16;
        gen_rows_kernel<<<grid, 256>>>(d, d_o0, d_o1, w, g->Mr_pad, g->Mc_pad, g->kmax, g->row_hi, g->row_lo);
        gen_cols_kernel<<<grid, 256>>>(d, d_rlo, d_rhi, d_o0, d_o1, w, g->kz, g->col_hi, g->col_lo);
        lsb::count_launch(2);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            lsb::set_error("generating the band operands: %s", cudaGetErrorString(cudaGetLastError()));
            rc = LSB_ERR_CUDA;
        }
    }
    for (void *p : tmp) cudaFree(p);
    if (!rc) rc = make_map(&g->map_row_hi, g->row_hi, (uint64_t)Lr * g->Mr_pad, (uint64_t)g->kmax);
    if (!rc) rc = make_map(&g->map_row_lo, g->row_lo, (uint64_t)Lr * g->Mr_pad, (uint64_t)g->kmax);
    if (!rc) rc = make_map(&g->map_col_hi, g->col_hi, (uint64_t)L * M, (uint64_t)g->kz);
    if (!rc) rc = make_map(&g->map_col_lo, g->col_lo, (uint64_t)L * M, (uint64_t)g->kz);
    if (rc) {
        lsb_band_design_destroy(g);
        return rc;
    }
    *out = g;
    return LSB_OK;
}

void lsb_band_design_destroy(lsb_band_design *g)
{
    if (!g) return;
    cudaDeviceSynchronize();
    for (void *p : g->allocs) cudaFree(p);
    for (auto &kv : g->work)
        for (void *p : kv.second.allocs) cudaFree(p);
    cudaFree(g->host_stage);
    if (g->host_stream) cudaStreamDestroy(g->host_stream);
    delete g;
}

/* bytes of device memory held by the explicit design (both orientations, fp16 pairs) */
int64_t lsb_band_design_bytes(const lsb_band_design *g)
{
    if (!g) return 0;
    return (int64_t)(2 * sizeof(tc_t)) *
           ((int64_t)g->dev.Lr * g->Mr_pad * g->kmax + (int64_t)g->dev.L * g->dev.M * g->kz);
}

int lsb_band_ab_dev_f32(lsb_band_design *g, const float *x, float *out, int64_t F, void *stream)
{
    if (F == 0) return LSB_OK;
    LSB_REQUIRE(g && x && out && F > 0, "null pointer or negative frame count");
    cudaStream_t st = (cudaStream_t)stream;
    lsb_band_design::Work *wk = nullptr;
    int rc = get_work(g, st, F, &wk);
    if (rc) return rc;
    const BandDev &d = g->dev;
    LSB_CUDA(lsb::ensure_dynamic_smem(tc_items_kernel, kSmemBytes));
    band_split_kernel<<<lsb::sm_count() * 8, 256, 0, st>>>(x, wk->st.b_hi, wk->st.b_lo, F, d.L * d.M, d.L * d.M, d.L * d.M,
                                                           d.L * d.M);
    launch_ab(g, *wk, F, st);
    band_sum_partials_kernel<<<lsb::sm_count() * 8, 256, 0, st>>>(d, wk->st.part, wk->st.part_stride, out, F);
    lsb::count_launch(2);
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

int lsb_band_az_dev_f32(lsb_band_design *g, const float *y, float *out, int64_t F, void *stream)
{
    if (F == 0) return LSB_OK;
    LSB_REQUIRE(g && y && out && F > 0, "null pointer or negative frame count");
    cudaStream_t st = (cudaStream_t)stream;
    lsb_band_design::Work *wk = nullptr;
    int rc = get_work(g, st, F, &wk);
    if (rc) return rc;
    const BandDev &d = g->dev;
    LSB_CUDA(lsb::ensure_dynamic_smem(tc_items_kernel, kSmemBytes));
    band_split_kernel<<<lsb::sm_count() * 8, 256, 0, st>>>(y, wk->st.z_hi, wk->st.z_lo, F, d.n, d.n_pitch, d.Mr, d.Mr_p);
    launch_az(g, *wk, F, st);
    lsb::count_launch(1);
    LSB_CUDA(cudaMemcpyAsync(out, wk->st.atz, (size_t)F * d.L * d.M * sizeof(float), cudaMemcpyDeviceToDevice, st));
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

int lsb_band_amp_dev_f32(lsb_band_design *g, const float *y, int64_t F, const lsb_sparc_amp_params *pp, const int32_t *idx0,
                         const lsb_sparc_amp_out *po, void *stream)
{
    if (F == 0) return LSB_OK;
    LSB_REQUIRE(g && y && pp && po && po->idx && po->t_final && po->psi, "null pointer");
    LSB_REQUIRE(F > 0 && F <= 65535, "frame count outside [1, 65535]");
    LSB_REQUIRE(pp->t_max >= 2 && (pp->phi_est_method == 1 || pp->phi_est_method == 2), "bad t_max / phi_est_method");
    LSB_REQUIRE(pp->modulation <= 1, "the tensor-core AMP path handles unmodulated SPARCs (K = 1)");
    LSB_REQUIRE(!po->nmse || idx0, "nmse output needs the true section indices idx0");
    cudaStream_t st = (cudaStream_t)stream;
    lsb_band_design::Work *wk = nullptr;
    int rc = get_work(g, st, F, &wk);
    if (rc) return rc;
    const BandDev &d = g->dev;
    BandParams p{pp->awgn_var, pp->rtol, pp->atol, pp->t_max, pp->phi_est_method};
    BandState s = wk->st;
    s.idx = po->idx;
    s.t_final = po->t_final;
    s.nmse = po->nmse;
    s.s_last = (float *)po->s_last;
    s.tau_trace = po->tau_trace;
    s.psi_trace = po->psi_trace;
    s.idx0 = idx0;
    LSB_CUDA(lsb::ensure_dynamic_smem(tc_items_kernel, kSmemBytes));
    LSB_CUDA(cudaMemsetAsync(s.done, 0, (size_t)F * sizeof(int), st));
    LSB_CUDA(cudaMemsetAsync(s.n_done, 0, sizeof(int), st));
    int t = 0, t_last = 0;
    for (; t < pp->t_max - 1; ++t) {
        if (t > 0) launch_ab(g, *wk, F, st);
        band_residual_kernel<<<(unsigned)F, kBandThreads, 0, st>>>(d, p, s, y, F, t);
        launch_az(g, *wk, F, st);
        switch (d.M >= 32 ? d.M / 32 : 0) {
            case 0: launch_denoise<0>(d, p, s, F, t, st); break;
            case 1: launch_denoise<1>(d, p, s, F, t, st); break;
            case 2: launch_denoise<2>(d, p, s, F, t, st); break;
            case 4: launch_denoise<4>(d, p, s, F, t, st); break;
            case 8: launch_denoise<8>(d, p, s, F, t, st); break;
            case 16: launch_denoise<16>(d, p, s, F, t, st); break;
            default: launch_denoise<32>(d, p, s, F, t, st); break;
        }
        lsb::count_launch(2);
        LSB_CUDA(cudaGetLastError());
        t_last = t;
        // every fourth iteration from the eighth on: have all frames stopped?  (the stop of iteration t - 1 is seen
        // by the residual kernel of iteration t, so this reads the state after iteration t - 1)
        if (t >= 8 && (t & 3) == 0) {
            int n_done = 0;
            LSB_CUDA(cudaMemcpyAsync(&n_done, s.n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
            LSB_CUDA(cudaStreamSynchronize(st));
            if (n_done == F) break;
        }
    }
    band_finish_kernel<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(d, p, s, F, t_last);
    lsb::count_launch();
    LSB_CUDA(cudaMemcpyAsync(po->psi, s.psi, (size_t)F * d.Lc * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (po->beta_soft)
        LSB_CUDA(cudaMemcpyAsync(po->beta_soft, s.beta, (size_t)F * d.L * d.M * sizeof(float), cudaMemcpyDeviceToDevice, st));
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

/* Host-pointer form (what a loop over the reference's sparc_decode, sparc.py:55-74, takes and returns): y [F, n] in
 * host memory (pinned memory makes the copies asynchronous), idx [F, L] / t_final [F] / psi [F, Lc] (optional) back in
 * host memory when the call returns.  One call at a time per handle. */
int lsb_band_amp_host_f32(lsb_band_design *g, const float *y, int64_t F, const lsb_sparc_amp_params *pp, int32_t *idx,
                          int32_t *t_final, double *psi)
{
    if (F == 0) return LSB_OK;
    LSB_REQUIRE(g && y && pp && idx && t_final, "null pointer");
    LSB_REQUIRE(F > 0 && F <= 65535, "frame count outside [1, 65535]");
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    std::lock_guard<std::mutex> lock(g->host_mu);
    if (!g->host_stream) LSB_CUDA(cudaStreamCreateWithFlags(&g->host_stream, cudaStreamNonBlocking));
    const BandDev &d = g->dev;
    const size_t by = (size_t)F * d.n * sizeof(float), bi = (size_t)F * d.L * sizeof(int32_t), bt = (size_t)F * sizeof(int32_t),
                 bp = (size_t)F * d.Lc * sizeof(double);
    const size_t need = by + bi + bt + bp + 64;
    if (g->host_stage_bytes < need) {
        LSB_CUDA(cudaStreamSynchronize(g->host_stream));
        cudaFree(g->host_stage);
        g->host_stage = nullptr;
        g->host_stage_bytes = 0;
        LSB_CUDA(cudaMalloc(&g->host_stage, need));
        g->host_stage_bytes = need;
    }
    char *base = (char *)g->host_stage;
    float *d_y = (float *)base;
    double *d_psi = (double *)(base + ((by + 15) & ~(size_t)15));
    int32_t *d_idx = (int32_t *)((char *)d_psi + bp);
    int32_t *d_t = (int32_t *)((char *)d_idx + bi);
    cudaStream_t st = g->host_stream;
    LSB_CUDA(cudaMemcpyAsync(d_y, y, by, cudaMemcpyHostToDevice, st));
    lsb_sparc_amp_out o{};
    o.idx = d_idx;
    o.t_final = d_t;
    o.psi = d_psi;
    int rc = lsb_band_amp_dev_f32(g, d_y, F, pp, nullptr, &o, st);
    if (rc) return rc;
    LSB_CUDA(cudaMemcpyAsync(idx, d_idx, bi, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaMemcpyAsync(t_final, d_t, bt, cudaMemcpyDeviceToHost, st));
    if (psi) LSB_CUDA(cudaMemcpyAsync(psi, d_psi, bp, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    return LSB_OK;
}

}  // extern "C"

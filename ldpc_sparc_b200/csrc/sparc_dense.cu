// Dense Gaussian design SPARC (sparc_sophie/sparc_new.py): frame-batched A*beta and
// A^T*z, the AMP residual / denoiser steps and the section <-> bit glue around the BP
// decoder.  First (round 1) version of this row: FP32 / FP64 SIMT kernels, correct and
// coalesced, not yet the tcgen05 path.
//
// Reference per AMP iteration (sparc_new.py:885-912, 975-990):
//   z   = y - A beta + (z / tau^2)(P - |beta|^2 / n)        (integrated: + (z / n) sum d_eta, :492)
//   s   = beta + A^T z ;  tau^2 = |z|^2 / n
//   beta = sqrt(nP_l) softmax_section(sqrt(nP_l) s / tau^2)                   (:1040-1066)
// Glue (sparc_new.py:1118-1138, 1162-1193, 1260-1279, 824-883): section posterior ->
// P(bit = 0) -> LLR -> [BP] -> APP -> P(bit = 0) -> product-form section prior.
#include "common.cuh"

#include <new>
#include <vector>

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ tiled GEMM
// C[F, N] = X[F, K] * op(A):  TRANS_A = true : A is [N, K] row-major (C = X A^T, the A*beta product)
//                             TRANS_A = false: A is [K, N] row-major (C = X A,   the A^T*z product)
template <typename T, int BM, int BN, bool TRANS_A>
__global__ void __launch_bounds__(kThreads)
gemm_kernel(const T *__restrict__ X, const T *__restrict__ A, T *__restrict__ C, int F, int N, int K)
{
    constexpr int BK = 16, TM = 4, TN = BM * BN / (kThreads * TM);
    __shared__ T Xs[BK][BM + 4];
    __shared__ T As[BK][BN + 4];
    const int tid = threadIdx.x;
    const int f0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);   // thread tile: rows ty*TM.., cols tx*TN..
    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = (T)0;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // X tile [BM x BK], k contiguous in memory
        for (int e = tid; e < BM * BK; e += kThreads) {
            const int m = e / BK, k = e % BK;
            const int f = f0 + m, kk = k0 + k;
            Xs[k][m] = (f < F && kk < K) ? X[(size_t)f * K + kk] : (T)0;
        }
        if (TRANS_A) {   // A[n, k]: k contiguous
            for (int e = tid; e < BN * BK; e += kThreads) {
                const int nn = e / BK, k = e % BK;
                const int n = n0 + nn, kk = k0 + k;
                As[k][nn] = (n < N && kk < K) ? A[(size_t)n * K + kk] : (T)0;
            }
        } else {         // A[k, n]: n contiguous
            for (int e = tid; e < BN * BK; e += kThreads) {
                const int k = e / BN, nn = e % BN;
                const int n = n0 + nn, kk = k0 + k;
                As[k][nn] = (n < N && kk < K) ? A[(size_t)kk * N + n] : (T)0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            T xr[TM], ar[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) xr[i] = Xs[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) ar[j] = As[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] += xr[i] * ar[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int f = f0 + ty * TM + i;
        if (f >= F) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n < N) C[(size_t)f * N + n] = acc[i][j];
        }
    }
}

// deterministic block sum (fixed tree), result valid in thread 0
template <typename T>
__device__ T block_sum(T v, T *red /* [kThreads/32] shared */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T tot = (T)0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
    __syncthreads();
    return tot;
}

// z <- y - Ab + z * coef, tau_sqr <- |z|^2 / n.   One CTA per frame.
//   mode 0: coef = (P - beta_sq / n) / tau_sqr_in      (sparc_new.py:904,981)
//   mode 1: coef = onsager_sum / n                     (sparc_new.py:492)
//   first != 0: z = y (t = 0 of sparc_amp, :898) -- Ab and coef ignored
template <typename T>
__global__ void __launch_bounds__(kThreads)
residual_kernel(const T *__restrict__ y, const T *__restrict__ ab, T *__restrict__ z, const double *beta_sq,
                const double *onsager_sum, double *tau_sqr, int n, double P, int mode, int first)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    double coef = 0.0;
    if (!first) coef = mode == 0 ? (P - beta_sq[f] / n) / tau_sqr[f] : onsager_sum[f] / n;
    double part = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const size_t p = f * n + i;
        const T v = first ? y[p] : (T)((y[p] - ab[p]) + z[p] * (T)coef);
        z[p] = v;
        part += (double)v * (double)v;
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) tau_sqr[f] = tot / n;
}

// s = beta_in + atz ; beta <- sqrt_nPl * softmax_section(sqrt_nPl * s / tau_sqr) ; beta_sq = |beta|^2.
// One CTA per frame, one warp per section.  PL = entries per lane (M = 32*PL) or 0 for
// M < 32 (32/M sections per warp pass).  The loads of the warp's next section are issued
// before the current one is reduced (the kernel was bound by the exposed load latency of
// one section at a time).
template <typename T, int PL>
__global__ void __launch_bounds__(kThreads)
denoise_kernel(const T *beta_in, const T *__restrict__ atz, T *beta, T *__restrict__ s_out,
               const double *tau_sqr, double *beta_sq, int L, int M, double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const bool add_beta = beta_in != nullptr;   // beta_in may alias beta (each entry is read before it is written)
    const size_t f = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T scale = (T)(sqrt_nPl / tau_sqr[f]);
    const T amp = (T)sqrt_nPl;
    const size_t frame_base = f * (size_t)L * M;
    double part = 0.0;
    if (PL > 0) {
        constexpr int P = PL > 0 ? PL : 1;
        constexpr int WARPS = kThreads / 32;
        T nx[P];
        int sec = warp;
        if (sec < L) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const size_t a = frame_base + (size_t)sec * M + lane + 32 * k;
                nx[k] = atz[a] + (add_beta ? beta_in[a] : (T)0);
            }
        }
        for (; sec < L; sec += WARPS) {
            T sv[P];
#pragma unroll
            for (int k = 0; k < P; ++k) sv[k] = nx[k];
            const int nsec = sec + WARPS;
            if (nsec < L) {                      // prefetch the next section of this warp
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const size_t a = frame_base + (size_t)nsec * M + lane + 32 * k;
                    nx[k] = atz[a] + (add_beta ? beta_in[a] : (T)0);
                }
            }
            const size_t base = frame_base + (size_t)sec * M;
            T mx = sv[0];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                if (s_out) s_out[base + lane + 32 * k] = sv[k];
                mx = fmax(mx, sv[k]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            T sum = (T)0;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                sv[k] = exp((sv[k] - mx) * scale);
                sum += sv[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const T inv = amp / sum;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const T b = sv[k] * inv;
                beta[base + lane + 32 * k] = b;
                part += (double)b * (double)b;
            }
        }
    } else {
        const int spw = 32 / M;      // sections per warp pass
        for (int sec0 = warp * spw; sec0 < L; sec0 += (kThreads / 32) * spw) {
            const int sec = sec0 + lane / M, j = lane % M;
            const bool live = sec < L;
            const size_t a = frame_base + (size_t)sec * M + j;
            T v = (T)0;
            if (live) {
                v = atz[a] + (add_beta ? beta_in[a] : (T)0);
                if (s_out) s_out[a] = v;
            }
            T mx = v;
            for (int o = M >> 1; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const T e = exp((v - mx) * scale);
            T sum = e;
            for (int o = M >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (live) {
                const T b = amp * (e / sum);
                beta[a] = b;
                part += (double)b * (double)b;
            }
        }
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) beta_sq[f] = tot;
}

// P(bit b of the section index = 0) = sum_{j: bit b of j == 0} beta_j / sqrt_nPl, summed in
// increasing j as the reference (sparc_new.py:1118-1138; b = 0 is the MSB); LLR = log p -
// log(1-p) with p clipped to [1e-15, 1-1e-15] (:1166-1169).  One thread per (frame, section, bit).
template <typename T>
__global__ void section_to_bit_kernel(const T *__restrict__ beta, T *__restrict__ probs, T *__restrict__ llr,
                                      long long total, int M, int logM, double sqrt_nPl)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long sec = i / logM;
    const int b = (int)(i % logM);
    const int shift = logM - 1 - b;
    const T *bs = beta + sec * M;
    T p = (T)0;
    for (int j = 0; j < M; ++j)
        if (((j >> shift) & 1) == 0) p += bs[j] / (T)sqrt_nPl;
    if (probs) probs[i] = p;
    if (llr) {
        if (sizeof(T) == 8) {
            const double pc = fmin(fmax((double)p, 1e-15), 1.0 - 1e-15);
            llr[i] = (T)(log(pc) - log(1.0 - pc));
        } else {
            // 1 - 1e-15 is not an fp32 number, so the clip is applied to the result instead: for p inside
            // [1e-15, 1 - 1e-15] nothing changes, outside it the reference's LLR is +-log((1 - 1e-15) / 1e-15);
            // 1 - p is exact in fp32 for p >= 1/2.  (Was two fp64 logs per bit: 135 -> ~45 us at config 5.)
            // The fp32 sum of softmax entries can round a saturated probability to 1 + ulp (or below 0):
            // log of a negative number is NaN and fminf/fmaxf would turn it into -lim, the wrong sign.
            // Saturate explicitly, as the reference's clip does (p >= 1 -> +lim, p <= 0 -> -lim).
            const float lim = 34.538776f;
            const float pf = (float)p;
            float v;
            if (pf >= 1.0f) v = lim;
            else if (pf <= 0.0f) v = -lim;
            else v = fminf(fmaxf(logf(pf) - logf(1.0f - pf), -lim), lim);
            llr[i] = (T)v;
        }
    }
}

// p_b = e^app / (1 + e^app) (:1190); beta_i = sqrt_nPl * prod_b (p_b or 1 - p_b) (:1260-1279).
// One CTA per frame; beta_sq[f] = |beta_f|^2 (optional) for the next Onsager term.
template <typename T>
__global__ void __launch_bounds__(kThreads)
bit_to_section_kernel(const T *__restrict__ app, T *__restrict__ probs, T *__restrict__ beta, double *beta_sq,
                      int L, int M, int logM, double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    double part = 0.0;
    for (int e = threadIdx.x; e < L * M; e += kThreads) {
    const long long i = (long long)f * L * M + e;
    const long long sec = (long long)f * L + (e >> logM);       // M = 2^logM
    const int idx = e & (M - 1);
    T prod = (T)1;
    for (int j = 0; j < logM; ++j) {
        const T a = app[sec * logM + j];
        T p;
        if (sizeof(T) == 8) {            // the reference's expression (overflows to NaN above 709, :1190)
            const T e = exp(a);
            p = e / ((T)1 + e);
        } else {                         // fp32: same value, without the overflow at a > 88
            p = (T)1 / ((T)1 + exp(-a));
        }
        if (idx == 0 && probs) probs[sec * logM + j] = p;
        prod = ((idx >> (logM - 1 - j)) & 1) ? prod * ((T)1 - p) : prod * p;
    }
    const T bv = prod * (T)sqrt_nPl;
    beta[i] = bv;
    part += (double)bv * (double)bv;
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0 && beta_sq) beta_sq[f] = tot;
}

// hard decisions of the K systematic bits of every LDPC block: bits[blk, k] = app[blk, k] < 0 (:1184-1186)
template <typename T>
__global__ void hard_bits_kernel(const T *__restrict__ app, uint8_t *__restrict__ bits, long long n_blocks, int N, int K)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_blocks * K) return;
    bits[i] = app[(i / K) * N + (i % K)] < (T)0;
}

// sum over the frame of beta_i * main_term_i (sparc_new.py:824-841,871-883), reference
// operation order inside a section.  One CTA per frame, one thread per (section, i).
template <typename T>
__global__ void __launch_bounds__(kThreads)
onsager_kernel(const T *__restrict__ beta, const T *__restrict__ vk, const T *__restrict__ vk0,
               const T *__restrict__ alpha_w /* weighted alpha = sqrt_nPl * alpha */, const double *tau_sqr,
               double *out, int L, int M, int logM, double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    const T c = (T)(sqrt_nPl / tau_sqr[f]);
    const T inv_amp = (T)(1.0 / sqrt_nPl);
    double part = 0.0;
    for (int e = threadIdx.x; e < L * M; e += kThreads) {
        const int l = e / M, i = e % M;
        const T *a = alpha_w + f * (size_t)L * M + (size_t)l * M;
        const T ai = a[i] / (T)sqrt_nPl;
        T main_term = (T)0;
        for (int k = 0; k < logM; ++k) {
            const int shift = logM - 1 - k;
            T sum_term = (T)0;
            for (int q = 0; q < M; ++q) {
                if ((q >> shift) & 1) continue;          // q in S_k: bit k of q is 0
                const T aq = a[q] / (T)sqrt_nPl;
                sum_term += (q == i) ? aq * c * ((T)1 - aq) : aq * c * (-ai);
            }
            const size_t bi = f * (size_t)L * logM + (size_t)l * logM + k;
            // clip in fp64: 1 - 1e-10 is not representable in fp32 (v0 = 1 would divide by zero)
            const double v0 = fmin(fmax((double)vk0[bi], 1e-10), 1.0 - 1e-10);
            const T sub = (T)(1.0 / (v0 * (1.0 - v0))) * sum_term;
            const T v = vk[bi];
            if ((i >> shift) & 1) main_term -= v * sub;
            else main_term += ((T)1 - v) * sub;
        }
        part += (double)(beta[f * (size_t)L * M + e] * main_term);
        (void)inv_amp;
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) out[f] = tot;
}

// Same sum in closed form for the fp32 modes: with alpha = alpha_w / sqrt_nPl and
// A_k = sum_{q in S_k} alpha_q, the inner sum over q of the reference collapses to
//   alpha_i c ([i in S_k] - A_k),
// so a section costs O(M logM) instead of O(M^2 logM).  G = min(M, 32) lanes
// share a section (PER = M / G entries each, compile time, at most 8), A_k by a shuffle
// reduction inside the group; one CTA per frame, fixed summation order.
template <typename T, int PER>
__global__ void __launch_bounds__(kThreads)
onsager_closed_kernel(const T *__restrict__ beta, const T *__restrict__ vk, const T *__restrict__ vk0,
                      const T *__restrict__ alpha_w, const double *tau_sqr, double *out, int L, int M, int logM,
                      double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    const T c = (T)(sqrt_nPl / tau_sqr[f]);
    const T inv_amp = (T)(1.0 / sqrt_nPl);
    const int G = M / PER;                                // lanes per section: M for M <= 32, else 32
    const int logG = 31 - __clz(G);
    const int gl = threadIdx.x & (G - 1);                 // lane inside the section's group
    const int groups = kThreads >> logG;
    const T *a_f = alpha_w + f * (size_t)L * M, *b_f = beta + f * (size_t)L * M;
    const T *v_f = vk + f * (size_t)L * logM, *v0_f = vk0 + f * (size_t)L * logM;
    double part = 0.0;
    for (int l0 = 0; l0 < L; l0 += groups) {
        const int l = l0 + (threadIdx.x >> logG);
        const bool live = l < L;                          // whole groups are live or not: shuffles stay inside a group
        const int e0 = (live ? l : 0) * M + gl;
        T a[PER], acc[PER], b[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            a[u] = live ? a_f[e0 + u * G] * inv_amp : (T)0;
            b[u] = live ? b_f[e0 + u * G] : (T)0;
            acc[u] = (T)0;
        }
        for (int k = 0; k < logM; ++k) {
            const int shift = logM - 1 - k;
            T Ak = (T)0;
#pragma unroll
            for (int u = 0; u < PER; ++u)
                if (!(((gl + u * G) >> shift) & 1)) Ak += a[u];
            for (int o = G >> 1; o > 0; o >>= 1) Ak += __shfl_xor_sync(0xffffffffu, Ak, o);
            const int bi = (live ? l : 0) * logM + k;
            // v0 clipped to [1e-10, 1 - 1e-10] (:881): 1 - v0 is formed in fp64 (1 - 1e-10 is not an fp32 number)
            const T v0 = v0_f[bi];
            const T lo = fmax(v0, (T)1e-10), hi = fmax((T)(1.0 - (double)v0), (T)1e-10);
            const T sub = (T)1 / (lo * hi);
            const T v = v_f[bi];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const bool one = ((gl + u * G) >> shift) & 1;
                acc[u] += (one ? -v : (T)1 - v) * sub * ((one ? (T)0 : (T)1) - Ak);
            }
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) part += (double)(b[u] * (c * a[u] * acc[u]));
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) out[f] = tot;
}

// ------------------------------------------------------------------ M = 2 (one LDPC bit per section), fp32
// BASELINE config 5 runs L = 14400 sections of M = 2.  The general kernels above give a section to a group of lanes
// and reduce across them with shuffles, which for M = 2 leaves a warp with 16 sections per pass and a chain of
// dependent loads per pass (ncu launch list, 256 frames: Onsager 178 us, denoiser 92 us, bit -> section 80 us against
// 10-20 us of memory traffic).  Here a thread owns whole sections (one 8-byte access per array), four in flight.
// Same expressions as the general kernels; only the order of the per-frame fp64 partial sums differs.
__global__ void __launch_bounds__(kThreads)
denoise_m2_kernel(const float *beta_in, const float *__restrict__ atz, float *beta, float *__restrict__ s_out,
                  const double *tau_sqr, double *beta_sq, int L, double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    const float scale = (float)(sqrt_nPl / tau_sqr[f]), amp = (float)sqrt_nPl;
    const float2 *az2 = reinterpret_cast<const float2 *>(atz) + f * (size_t)L;
    const float2 *bi2 = beta_in ? reinterpret_cast<const float2 *>(beta_in) + f * (size_t)L : nullptr;
    float2 *bo2 = reinterpret_cast<float2 *>(beta) + f * (size_t)L;
    float2 *so2 = s_out ? reinterpret_cast<float2 *>(s_out) + f * (size_t)L : nullptr;
    double part = 0.0;
#pragma unroll 4
    for (int l = threadIdx.x; l < L; l += kThreads) {
        float2 v = az2[l];
        if (bi2) { const float2 b = bi2[l]; v.x += b.x; v.y += b.y; }
        if (so2) so2[l] = v;
        const float mx = fmaxf(v.x, v.y);
        const float e0 = expf((v.x - mx) * scale), e1 = expf((v.y - mx) * scale);
        const float sum = e0 + e1;
        const float b0 = amp * (e0 / sum), b1 = amp * (e1 / sum);
        bo2[l] = make_float2(b0, b1);
        part += (double)b0 * (double)b0 + (double)b1 * (double)b1;
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) beta_sq[f] = tot;
}

__global__ void __launch_bounds__(kThreads)
bit_to_section_m2_kernel(const float *__restrict__ app, float *__restrict__ probs, float *__restrict__ beta, double *beta_sq,
                         int L, double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    const float amp = (float)sqrt_nPl;
    float2 *bo2 = reinterpret_cast<float2 *>(beta) + f * (size_t)L;
    double part = 0.0;
#pragma unroll 4
    for (int l = threadIdx.x; l < L; l += kThreads) {
        const float p = 1.0f / (1.0f + expf(-app[f * (size_t)L + l]));
        if (probs) probs[f * (size_t)L + l] = p;
        const float b0 = p * amp, b1 = (1.0f - p) * amp;
        bo2[l] = make_float2(b0, b1);
        part += (double)b0 * (double)b0 + (double)b1 * (double)b1;
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0 && beta_sq) beta_sq[f] = tot;
}

__global__ void __launch_bounds__(kThreads)
onsager_m2_kernel(const float *__restrict__ beta, const float *__restrict__ vk, const float *__restrict__ vk0,
                  const float *__restrict__ alpha_w, const double *tau_sqr, double *out, int L, double sqrt_nPl)
{
    __shared__ double red[kThreads / 32];
    const size_t f = blockIdx.x;
    const float c = (float)(sqrt_nPl / tau_sqr[f]), inv_amp = (float)(1.0 / sqrt_nPl);
    const float2 *a2 = reinterpret_cast<const float2 *>(alpha_w) + f * (size_t)L;
    const float2 *b2 = reinterpret_cast<const float2 *>(beta) + f * (size_t)L;
    const float *v_f = vk + f * (size_t)L, *v0_f = vk0 + f * (size_t)L;
    double part = 0.0;
#pragma unroll 4
    for (int l = threadIdx.x; l < L; l += kThreads) {
        const float2 aw = a2[l], b = b2[l];
        const float a0 = aw.x * inv_amp, a1 = aw.y * inv_amp;
        const float Ak = a0;                                   // S_0 = {0}: entries whose (only) bit is 0
        const float v0 = v0_f[l], v = v_f[l];
        const float lo = fmaxf(v0, 1e-10f), hi = fmaxf((float)(1.0 - (double)v0), 1e-10f);
        const float sub = 1.0f / (lo * hi);
        const float acc0 = (1.0f - v) * sub * (1.0f - Ak);
        const float acc1 = (-v) * sub * (0.0f - Ak);
        part += (double)(b.x * (c * a0 * acc0));
        part += (double)(b.y * (c * a1 * acc1));
    }
    const double tot = block_sum(part, red);
    if (threadIdx.x == 0) out[f] = tot;
}

template <typename T>
void launch_onsager(const T *beta, const T *vk, const T *vk0, const T *alpha_w, const double *tau_sqr, double *out,
                    int64_t F, int L, int M, int logM, double sqrt_nPl, cudaStream_t st)
{
    // fp64 is the parity mode and keeps the reference's operation order; fp32 takes the closed form
#define LSB_ONSAGER_CLOSED(PER) \
    onsager_closed_kernel<T, PER><<<(unsigned)F, kThreads, 0, st>>>(beta, vk, vk0, alpha_w, tau_sqr, out, L, M, logM, sqrt_nPl)
    if (sizeof(T) == 4 && M == 2)
        onsager_m2_kernel<<<(unsigned)F, kThreads, 0, st>>>((const float *)beta, (const float *)vk, (const float *)vk0,
                                                         (const float *)alpha_w, tau_sqr, out, L, sqrt_nPl);
    else if (sizeof(T) == 4 && M <= 32) LSB_ONSAGER_CLOSED(1);
    else if (sizeof(T) == 4 && M == 64) LSB_ONSAGER_CLOSED(2);
    else if (sizeof(T) == 4 && M == 128) LSB_ONSAGER_CLOSED(4);
    else if (sizeof(T) == 4 && M == 256) LSB_ONSAGER_CLOSED(8);
    else
        onsager_kernel<T><<<(unsigned)F, kThreads, 0, st>>>(beta, vk, vk0, alpha_w, tau_sqr, out, L, M, logM, sqrt_nPl);
#undef LSB_ONSAGER_CLOSED
}

}  // namespace

struct lsb_dense_design {
    int device = 0;
    int64_t n = 0, ML = 0;
    float *d_A32 = nullptr;
    double *d_A64 = nullptr;
};

namespace {

int check_device(const lsb_dense_design *g)
{
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    return LSB_OK;
}

template <typename T> const T *mat(const lsb_dense_design *g);
template <> const float *mat<float>(const lsb_dense_design *g) { return g->d_A32; }
template <> const double *mat<double>(const lsb_dense_design *g) { return g->d_A64; }

template <typename T>
int run_gemm(const lsb_dense_design *g, bool transpose, const T *x, T *out, int64_t F, cudaStream_t st)
{
    if (F == 0) return LSB_OK;                     // an empty batch is valid with any pointers
    LSB_REQUIRE(g && x && out, "null pointer");
    LSB_REQUIRE(mat<T>(g), "the design holds no %s copy of A", sizeof(T) == 4 ? "fp32" : "fp64");
    LSB_REQUIRE(F >= 0, "negative frame count");
    if (F == 0) return LSB_OK;
    int rc = check_device(g);
    if (rc) return rc;
    const int n = (int)g->n, ML = (int)g->ML;
    if (!transpose) {   // out[F, n] = x[F, ML] A^T
        if (F > 16) {
            dim3 grid((n + 63) / 64, (unsigned)((F + 63) / 64));
            gemm_kernel<T, 64, 64, true><<<grid, kThreads, 0, st>>>(x, mat<T>(g), out, (int)F, n, ML);
        } else {
            dim3 grid((n + 255) / 256, (unsigned)((F + 15) / 16));
            gemm_kernel<T, 16, 256, true><<<grid, kThreads, 0, st>>>(x, mat<T>(g), out, (int)F, n, ML);
        }
    } else {            // out[F, ML] = x[F, n] A
        if (F > 16) {
            dim3 grid((ML + 63) / 64, (unsigned)((F + 63) / 64));
            gemm_kernel<T, 64, 64, false><<<grid, kThreads, 0, st>>>(x, mat<T>(g), out, (int)F, ML, n);
        } else {
            dim3 grid((ML + 255) / 256, (unsigned)((F + 15) / 16));
            gemm_kernel<T, 16, 256, false><<<grid, kThreads, 0, st>>>(x, mat<T>(g), out, (int)F, ML, n);
        }
    }
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

inline unsigned blocks_for(long long total) { return (unsigned)((total + kThreads - 1) / kThreads); }

}  // namespace

extern "C" {

int lsb_dense_design_create(const double *A, int64_t n, int64_t ML, int want_f32, int want_f64,
                            lsb_dense_design **out)
{
    LSB_REQUIRE(A && out, "null pointer");
    LSB_REQUIRE(n > 0 && ML > 0 && n < (1 << 30) && ML < (1 << 30), "bad design size");
    LSB_REQUIRE(want_f32 || want_f64, "ask for at least one precision");
    lsb_dense_design *g = new (std::nothrow) lsb_dense_design();
    if (!g) return LSB_ERR_NOMEM;
    g->n = n;
    g->ML = ML;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device");
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    const size_t count = (size_t)n * (size_t)ML;
    cudaError_t e = cudaSuccess;
    if (want_f64) {
        e = cudaMalloc((void **)&g->d_A64, count * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(g->d_A64, A, count * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && want_f32) {
        std::vector<float> tmp;
        try { tmp.resize(count); } catch (...) { e = cudaErrorMemoryAllocation; }
        if (e == cudaSuccess) {
            for (size_t i = 0; i < count; ++i) tmp[i] = (float)A[i];
            e = cudaMalloc((void **)&g->d_A32, count * sizeof(float));
            if (e == cudaSuccess) e = cudaMemcpy(g->d_A32, tmp.data(), count * sizeof(float), cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        lsb::set_error("uploading the design matrix: %s", cudaGetErrorString(e));
        lsb_dense_design_destroy(g);
        return LSB_ERR_CUDA;
    }
    *out = g;
    return LSB_OK;
}

void lsb_dense_design_destroy(lsb_dense_design *g)
{
    if (!g) return;
    cudaFree(g->d_A32);
    cudaFree(g->d_A64);
    delete g;
}

int lsb_dense_ab_dev_f32(const lsb_dense_design *g, const float *x, float *out, int64_t F, void *stream)
{ return run_gemm<float>(g, false, x, out, F, (cudaStream_t)stream); }
int lsb_dense_ab_dev_f64(const lsb_dense_design *g, const double *x, double *out, int64_t F, void *stream)
{ return run_gemm<double>(g, false, x, out, F, (cudaStream_t)stream); }
int lsb_dense_az_dev_f32(const lsb_dense_design *g, const float *z, float *out, int64_t F, void *stream)
{ return run_gemm<float>(g, true, z, out, F, (cudaStream_t)stream); }
int lsb_dense_az_dev_f64(const lsb_dense_design *g, const double *z, double *out, int64_t F, void *stream)
{ return run_gemm<double>(g, true, z, out, F, (cudaStream_t)stream); }

#define LSB_DENSE_ELEMWISE(SUF, T)                                                                              \
    int lsb_amp_residual_dev_##SUF(const T *y, const T *ab, T *z, const double *beta_sq,                         \
                                   const double *onsager_sum, double *tau_sqr, int64_t F, int n, double P,       \
                                   int mode, int first, void *stream)                                            \
    {                                                                                                            \
        if (F <= 0) return LSB_OK;                                                                               \
        LSB_REQUIRE(y && z && tau_sqr && (first || ab), "null pointer");                                         \
        LSB_REQUIRE(first || (mode == 0 ? beta_sq != nullptr : onsager_sum != nullptr), "missing coefficient");  \
        if (F <= 0) return LSB_OK;                                                                               \
        residual_kernel<T><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(y, ab, z, beta_sq, onsager_sum,   \
                                                                               tau_sqr, n, P, mode, first);     \
        lsb::count_launch();                                                                                     \
        LSB_CUDA(cudaGetLastError());                                                                            \
        return LSB_OK;                                                                                           \
    }                                                                                                            \
    int lsb_amp_denoise_dev_##SUF(const T *beta_in, const T *atz, T *beta_out, T *s_out,                         \
                                  const double *tau_sqr, double *beta_sq, int64_t F, int L, int M,               \
                                  double sqrt_nPl, void *stream)                                                 \
    {                                                                                                            \
        if (F <= 0) return LSB_OK;                                                                               \
        LSB_REQUIRE(beta_out && atz && tau_sqr && beta_sq, "null pointer");                                      \
        LSB_REQUIRE(M >= 2 && M <= 256 && (M & (M - 1)) == 0, "M must be a power of two in [2, 256]");            \
        if (F <= 0) return LSB_OK;                                                                               \
        if (sizeof(T) == 4 && M == 2 && ((uintptr_t)atz & 7) == 0 && ((uintptr_t)beta_out & 7) == 0 &&           \
            (!beta_in || ((uintptr_t)beta_in & 7) == 0) && (!s_out || ((uintptr_t)s_out & 7) == 0)) {             \
            denoise_m2_kernel<<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(                                \
                (const float *)beta_in, (const float *)atz, (float *)beta_out, (float *)s_out, tau_sqr, beta_sq, L, sqrt_nPl); \
            lsb::count_launch();                                                                                 \
            LSB_CUDA(cudaGetLastError());                                                                        \
            return LSB_OK;                                                                                       \
        }                                                                                                        \
        switch (M >= 32 ? M / 32 : 0) {                                                                          \
            case 0: denoise_kernel<T, 0><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(beta_in, atz, beta_out, s_out, tau_sqr, beta_sq, L, M, sqrt_nPl); break; \
            case 1: denoise_kernel<T, 1><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(beta_in, atz, beta_out, s_out, tau_sqr, beta_sq, L, M, sqrt_nPl); break; \
            case 2: denoise_kernel<T, 2><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(beta_in, atz, beta_out, s_out, tau_sqr, beta_sq, L, M, sqrt_nPl); break; \
            case 4: denoise_kernel<T, 4><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(beta_in, atz, beta_out, s_out, tau_sqr, beta_sq, L, M, sqrt_nPl); break; \
            default: denoise_kernel<T, 8><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(beta_in, atz, beta_out, s_out, tau_sqr, beta_sq, L, M, sqrt_nPl); break; \
        }                                                                                                        \
        lsb::count_launch();                                                                                     \
        LSB_CUDA(cudaGetLastError());                                                                            \
        return LSB_OK;                                                                                           \
    }                                                                                                            \
    int lsb_section_to_bit_dev_##SUF(const T *beta, T *probs, T *llr, int64_t n_sections, int M,                 \
                                     double sqrt_nPl, void *stream)                                              \
    {                                                                                                            \
        if (n_sections <= 0) return LSB_OK;                                                                      \
        LSB_REQUIRE(beta && (probs || llr), "null pointer");                                                     \
        int logM = 0;                                                                                            \
        while ((1 << logM) < M) ++logM;                                                                          \
        const long long total = (long long)n_sections * logM;                                                    \
        if (total <= 0) return LSB_OK;                                                                           \
        section_to_bit_kernel<T><<<blocks_for(total), kThreads, 0, (cudaStream_t)stream>>>(beta, probs, llr,     \
                                                                                          total, M, logM,       \
                                                                                          sqrt_nPl);            \
        lsb::count_launch();                                                                                     \
        LSB_CUDA(cudaGetLastError());                                                                            \
        return LSB_OK;                                                                                           \
    }                                                                                                            \
    int lsb_bit_to_section_dev_##SUF(const T *app, T *probs, T *beta, double *beta_sq, int64_t F, int L, int M,  \
                                     double sqrt_nPl, void *stream)                                              \
    {                                                                                                            \
        if (F <= 0) return LSB_OK;                                                                               \
        LSB_REQUIRE(app && beta, "null pointer");                                                                \
        int logM = 0;                                                                                            \
        while ((1 << logM) < M) ++logM;                                                                          \
        if (F <= 0) return LSB_OK;                                                                               \
        if (sizeof(T) == 4 && M == 2 && ((uintptr_t)beta & 7) == 0) {                                            \
            bit_to_section_m2_kernel<<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(                         \
                (const float *)app, (float *)probs, (float *)beta, beta_sq, L, sqrt_nPl);                         \
            lsb::count_launch();                                                                                 \
            LSB_CUDA(cudaGetLastError());                                                                        \
            return LSB_OK;                                                                                       \
        }                                                                                                        \
        bit_to_section_kernel<T><<<(unsigned)F, kThreads, 0, (cudaStream_t)stream>>>(app, probs, beta, beta_sq,  \
                                                                                    L, M, logM, sqrt_nPl);      \
        lsb::count_launch();                                                                                     \
        LSB_CUDA(cudaGetLastError());                                                                            \
        return LSB_OK;                                                                                           \
    }                                                                                                            \
    int lsb_hard_info_bits_dev_##SUF(const T *app, uint8_t *bits, int64_t n_blocks, int N, int K, void *stream)  \
    {                                                                                                            \
        if (n_blocks <= 0) return LSB_OK;                                                                        \
        LSB_REQUIRE(app && bits, "null pointer");                                                                \
        if (n_blocks <= 0) return LSB_OK;                                                                        \
        hard_bits_kernel<T><<<blocks_for((long long)n_blocks * K), kThreads, 0, (cudaStream_t)stream>>>(         \
            app, bits, (long long)n_blocks, N, K);                                                               \
        lsb::count_launch();                                                                                     \
        LSB_CUDA(cudaGetLastError());                                                                            \
        return LSB_OK;                                                                                           \
    }                                                                                                            \
    int lsb_integrated_onsager_dev_##SUF(const T *beta, const T *vk, const T *vk0, const T *alpha_weighted,      \
                                         const double *tau_sqr, double *out, int64_t F, int L, int M,            \
                                         double sqrt_nPl, void *stream)                                          \
    {                                                                                                            \
        if (F <= 0) return LSB_OK;                                                                               \
        LSB_REQUIRE(beta && vk && vk0 && alpha_weighted && tau_sqr && out, "null pointer");                      \
        int logM = 0;                                                                                            \
        while ((1 << logM) < M) ++logM;                                                                          \
        if (F <= 0) return LSB_OK;                                                                               \
        launch_onsager<T>(beta, vk, vk0, alpha_weighted, tau_sqr, out, F, L, M, logM, sqrt_nPl,                  \
                          (cudaStream_t)stream);                                                                 \
        lsb::count_launch();                                                                                     \
        LSB_CUDA(cudaGetLastError());                                                                            \
        return LSB_OK;                                                                                           \
    }

LSB_DENSE_ELEMWISE(f32, float)
LSB_DENSE_ELEMWISE(f64, double)

}  // extern "C"

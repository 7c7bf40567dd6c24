// State evolution of SPARC AMP decoding: the Monte-Carlo expectation of
// sparc_public/sparc_se.py:82-118 (sparc_se_E); for K = 1 and K = 2 (real samples),
//
//   E(tau) = mean_i  A_i / (A_i + sum_{j>=1} C_ij),
//     K = 1:  A_i = exp (1/tau + u_i0 / sqrt(tau)),  C_ij = exp (u_ij / sqrt(tau))
//     K = 2:  A_i = sinh(1/tau + u_i0 / sqrt(tau)),  C_ij = cosh(u_ij / sqrt(tau))
//
// over mc_samples rows of M standard normals u (drawn once and reused for every
// iteration and column block, as the reference does, :150-153).  One warp per sample,
// lanes over the M entries, every exponent shifted by the row's largest magnitude (the
// reference's plain exp overflows to inf/inf = nan for 1/tau > 709; the shifted form is
// the same number wherever the reference is finite).  Per-CTA partial sums are added in
// a fixed order so the result does not depend on scheduling.
#include "common.cuh"

#include <new>
#include <vector>

namespace {

constexpr int kThreads = 256;
constexpr int kSamplesPerCta = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads)
se_expect_kernel(const T *__restrict__ u, long long mc, int M, int K, const double *__restrict__ tau, double *__restrict__ part)
{
    __shared__ double red[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double itau = 1.0 / tau[blockIdx.y], rtau = sqrt(itau);
    const long long i0 = (long long)blockIdx.x * kSamplesPerCta;
    double acc = 0.0;
    for (long long i = i0 + warp; i < i0 + kSamplesPerCta && i < mc; i += kThreads / 32) {
        const T *row = u + i * M;
        const double a = itau + rtau * (double)row[0];
        double mx = fabs(a);
        for (int j = 1 + lane; j < M; j += 32) mx = fmax(mx, fabs(rtau * (double)row[j]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sum = 0.0;
        for (int j = 1 + lane; j < M; j += 32) {
            const double c = rtau * (double)row[j];
            sum += K == 1 ? exp(c - mx) : 0.5 * (exp(c - mx) + exp(-c - mx));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            const double top = K == 1 ? exp(a - mx) : 0.5 * (exp(a - mx) - exp(-a - mx));
            acc += top / (top + sum);
        }
    }
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
        part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

// Complex designs with K > 2 (sparc_se.py:100-113): u complex,
//   A_i = sum_k Re(c_k) e^{Re(a_i conj c_k)},  B_i = sum_k e^{Re(a_i conj c_k)},  a_i = 1/tau + u_i0 / sqrt(tau),
//   C_ij = sum_k e^{Re(u_ij conj c_k) / sqrt(tau)},        E = mean_i A_i / (B_i + sum_j C_ij)
// (K = 4: A = 2 sinh Re a, B = 2 cosh Re a + 2 cosh Im a, C = 2 cosh Re + 2 cosh Im -- the
// reference's sinh / cosh form; K > 4: its means over k, the factor 1/K cancels).
template <typename T>
__global__ void __launch_bounds__(kThreads)
se_expect_complex_kernel(const T *__restrict__ u /* [mc, M, 2] */, long long mc, int M, int K,
                         const double *__restrict__ tau, double *__restrict__ part)
{
    __shared__ double red[kThreads / 32];
    __shared__ double con_re[64], con_im[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < K) {
        double sn, cs;
        sincospi(2.0 * threadIdx.x / K, &sn, &cs);
        if (K <= 4) { cs = rint(cs); sn = rint(sn); }
        con_re[threadIdx.x] = cs;
        con_im[threadIdx.x] = sn;
    }
    __syncthreads();
    const double itau = 1.0 / tau[blockIdx.y], rtau = sqrt(itau);
    const long long i0 = (long long)blockIdx.x * kSamplesPerCta;
    double acc = 0.0;
    for (long long i = i0 + warp; i < i0 + kSamplesPerCta && i < mc; i += kThreads / 32) {
        const T *row = u + i * M * 2;
        const double ar = itau + rtau * (double)row[0], ai = rtau * (double)row[1];
        double mx = -INFINITY;
        for (int k = 0; k < K; ++k) mx = fmax(mx, ar * con_re[k] + ai * con_im[k]);
        for (int j = 1 + lane; j < M; j += 32) {
            const double vr = rtau * (double)row[2 * j], vi = rtau * (double)row[2 * j + 1];
            for (int k = 0; k < K; ++k) mx = fmax(mx, vr * con_re[k] + vi * con_im[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double sum = 0.0;
        for (int j = 1 + lane; j < M; j += 32) {
            const double vr = rtau * (double)row[2 * j], vi = rtau * (double)row[2 * j + 1];
            for (int k = 0; k < K; ++k) sum += exp(vr * con_re[k] + vi * con_im[k] - mx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            double top = 0.0, bot = 0.0;
            for (int k = 0; k < K; ++k) {
                const double e = exp(ar * con_re[k] + ai * con_im[k] - mx);
                top += con_re[k] * e;
                bot += e;
            }
            acc += top / (bot + sum);
        }
    }
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
        part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

// E[t] = (sum of the n_part partials of tau t, in index order) / mc
__global__ void se_finish_kernel(const double *__restrict__ part, int n_part, long long mc, double *__restrict__ E)
{
    const int t = blockIdx.x;
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < n_part; ++k) tot += part[(size_t)t * n_part + k];
        E[t] = tot / (double)mc;
    }
}

template <typename T>
int run_se(const T *u, int64_t mc, int M, int K, const double *tau, int n_tau, double *E, cudaStream_t st, bool cplx)
{
    LSB_REQUIRE(u && tau && E, "null pointer");
    LSB_REQUIRE(mc > 0 && M >= 1 && n_tau >= 1, "bad mc_samples / M / n_tau");
    if (cplx) LSB_REQUIRE(K >= 4 && K <= 64 && (K & (K - 1)) == 0, "complex samples need K-PSK with K = 4..64");
    else LSB_REQUIRE(K == 1 || K == 2, "real samples need K = 1 or K = 2");
    const int n_part = (int)((mc + kSamplesPerCta - 1) / kSamplesPerCta);
    double *d_tau = nullptr, *d_part = nullptr, *d_E = nullptr;
    LSB_CUDA(cudaMallocAsync((void **)&d_tau, sizeof(double) * n_tau, st));
    LSB_CUDA(cudaMallocAsync((void **)&d_part, sizeof(double) * (size_t)n_tau * n_part, st));
    LSB_CUDA(cudaMallocAsync((void **)&d_E, sizeof(double) * n_tau, st));
    LSB_CUDA(cudaMemcpyAsync(d_tau, tau, sizeof(double) * n_tau, cudaMemcpyHostToDevice, st));
    const dim3 grid((unsigned)n_part, (unsigned)n_tau);
    if (cplx) se_expect_complex_kernel<T><<<grid, kThreads, 0, st>>>(u, (long long)mc, M, K, d_tau, d_part);
    else se_expect_kernel<T><<<grid, kThreads, 0, st>>>(u, (long long)mc, M, K, d_tau, d_part);
    lsb::count_launch();
    se_finish_kernel<<<n_tau, 32, 0, st>>>(d_part, n_part, (long long)mc, d_E);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    LSB_CUDA(cudaMemcpyAsync(E, d_E, sizeof(double) * n_tau, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaFreeAsync(d_tau, st));
    LSB_CUDA(cudaFreeAsync(d_part, st));
    LSB_CUDA(cudaFreeAsync(d_E, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    return LSB_OK;
}

}  // namespace

extern "C" {

int lsb_sparc_se_expect_f64(const double *u_dev, int64_t mc_samples, int M, int K, const double *tau_host, int n_tau,
                            double *E_host, void *stream)
{
    return run_se<double>(u_dev, mc_samples, M, K, tau_host, n_tau, E_host, (cudaStream_t)stream, false);
}

int lsb_sparc_se_expect_f32(const float *u_dev, int64_t mc_samples, int M, int K, const double *tau_host, int n_tau,
                            double *E_host, void *stream)
{
    return run_se<float>(u_dev, mc_samples, M, K, tau_host, n_tau, E_host, (cudaStream_t)stream, false);
}

int lsb_sparc_se_expect_c128(const double *u_dev, int64_t mc_samples, int M, int K, const double *tau_host, int n_tau,
                             double *E_host, void *stream)
{
    return run_se<double>(u_dev, mc_samples, M, K, tau_host, n_tau, E_host, (cudaStream_t)stream, true);
}

int lsb_sparc_se_expect_c64(const float *u_dev, int64_t mc_samples, int M, int K, const double *tau_host, int n_tau,
                            double *E_host, void *stream)
{
    return run_se<float>(u_dev, mc_samples, M, K, tau_host, n_tau, E_host, (cudaStream_t)stream, true);
}

}  // extern "C"

// State evolution of SPARC AMP decoding: the Monte-Carlo expectation of
// sparc_public/sparc_se.py:82-118 (sparc_se_E) for K = 1 and real K = 2,
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
int run_se(const T *u, int64_t mc, int M, int K, const double *tau, int n_tau, double *E, cudaStream_t st)
{
    LSB_REQUIRE(u && tau && E, "null pointer");
    LSB_REQUIRE(mc > 0 && M >= 1 && n_tau >= 1, "bad mc_samples / M / n_tau");
    LSB_REQUIRE(K == 1 || K == 2, "state evolution is built for K = 1 and real K = 2 only");
    const int n_part = (int)((mc + kSamplesPerCta - 1) / kSamplesPerCta);
    double *d_tau = nullptr, *d_part = nullptr, *d_E = nullptr;
    LSB_CUDA(cudaMallocAsync((void **)&d_tau, sizeof(double) * n_tau, st));
    LSB_CUDA(cudaMallocAsync((void **)&d_part, sizeof(double) * (size_t)n_tau * n_part, st));
    LSB_CUDA(cudaMallocAsync((void **)&d_E, sizeof(double) * n_tau, st));
    LSB_CUDA(cudaMemcpyAsync(d_tau, tau, sizeof(double) * n_tau, cudaMemcpyHostToDevice, st));
    se_expect_kernel<T><<<dim3((unsigned)n_part, (unsigned)n_tau), kThreads, 0, st>>>(u, (long long)mc, M, K, d_tau, d_part);
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
    return run_se<double>(u_dev, mc_samples, M, K, tau_host, n_tau, E_host, (cudaStream_t)stream);
}

int lsb_sparc_se_expect_f32(const float *u_dev, int64_t mc_samples, int M, int K, const double *tau_host, int n_tau,
                            double *E_host, void *stream)
{
    return run_se<float>(u_dev, mc_samples, M, K, tau_host, n_tau, E_host, (cudaStream_t)stream);
}

}  // extern "C"

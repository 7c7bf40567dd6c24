// Pipe-throughput micro-benchmarks behind the AMP kernel's design choices (sm_100a):
//   packed fp32 (FFMA2 / FADD2) against scalar FFMA, legacy mma.sync (HMMA) and LDS.128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench tools/ubench/ubench.cu
// Prints warp-instructions per clock per SM and the equivalent lane-operations.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kIters = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) fp_kernel(float *out, float seed, long long *cycles)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-6f, -1e-6f);
    int n = threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }      // 2 FFMA
            if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);                                             // 1 FFMA2
            if (MODE == 2) a[i] = __fadd2_rn(a[i], c);                                                // 1 FADD2
            if (MODE == 3) { a[i].x += c.x; a[i].y += c.y; }                                          // 2 FADD
            if (MODE == 4) { a[i] = __ffma2_rn(a[i], b, c); n = (n ^ (n << 1)) + i; }                 // FFMA2 + integer
            if (MODE == 5) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); n = (n ^ (n << 1)) + i; }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + n;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) hmma_kernel(float *out, long long *cycles)
{
    unsigned a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x38003800u, 0x38003800u};
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-3f;
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int BYTES, bool STORE>
__global__ void __launch_bounds__(256) lds_kernel(float *out, long long *cycles)
{
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += 256) sm[i] = i;
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < kIters / 4; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int slot = (threadIdx.x + 256 * i + it) & 2047;
            if (BYTES == 16) {
                float4 *p = reinterpret_cast<float4 *>(sm) + slot;
                if (STORE) *p = acc; else { float4 v = *p; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            } else {
                float2 *p = reinterpret_cast<float2 *>(sm) + slot;
                if (STORE) *p = make_float2(acc.x, acc.y); else { float2 v = *p; acc.x += v.x; acc.y += v.y; }
            }
        }
        if (STORE) acc.x += 1.f;
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w + sm[threadIdx.x];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static long long h_cycles[148 * 8];

template <typename F>
void run(const char *name, F launch, int ctas_per_sm, double warp_instr_per_thread_iter, double lane_ops_per_instr, int iters)
{
    float *out;
    long long *cyc;
    const int grid = 148 * ctas_per_sm;
    CK(cudaMalloc(&out, grid * 256 * sizeof(float)));
    CK(cudaMalloc(&cyc, grid * sizeof(long long)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch(grid, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    launch(grid, out, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    CK(cudaMemcpy(h_cycles, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < grid; ++i) mean += h_cycles[i];
    mean /= grid;
    const double warp_instr = (double)iters * warp_instr_per_thread_iter * 8 /*warps*/ * ctas_per_sm;
    printf("%-28s ctas/SM %d  %.0f cycles  %.3f warp-instr/clk/SM  %.1f lane-ops/clk/SM  (%.3f ms)\n", name, ctas_per_sm,
           mean, warp_instr / mean, warp_instr * 32 * lane_ops_per_instr / mean, ms);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    float *dummy;
    CK(cudaMalloc(&dummy, 4));
    for (int occ : {1, 2, 4}) {
        run("FFMA x2 (scalar)", [](int g, float *o, long long *c) { fp_kernel<0><<<g, 256>>>(o, 1.f, c); }, occ, 16, 1, kIters);
        run("FFMA2 (packed)", [](int g, float *o, long long *c) { fp_kernel<1><<<g, 256>>>(o, 1.f, c); }, occ, 8, 2, kIters);
        run("FADD2 (packed)", [](int g, float *o, long long *c) { fp_kernel<2><<<g, 256>>>(o, 1.f, c); }, occ, 8, 2, kIters);
        run("FADD x2 (scalar)", [](int g, float *o, long long *c) { fp_kernel<3><<<g, 256>>>(o, 1.f, c); }, occ, 16, 1, kIters);
        run("FFMA2 + 2 int ops", [](int g, float *o, long long *c) { fp_kernel<4><<<g, 256>>>(o, 1.f, c); }, occ, 8, 2, kIters);
        run("FFMA x2 + 2 int ops", [](int g, float *o, long long *c) { fp_kernel<5><<<g, 256>>>(o, 1.f, c); }, occ, 16, 1, kIters);
        run("HMMA m16n8k16 f16->f32", [](int g, float *o, long long *c) { hmma_kernel<<<g, 256>>>(o, c); }, occ, 8, 16 * 8 * 16 / 32.0, kIters);
        run("LDS.128", [](int g, float *o, long long *c) { lds_kernel<16, false><<<g, 256, 32768>>>(o, c); }, occ, 8, 16, kIters / 4);
        run("STS.128", [](int g, float *o, long long *c) { lds_kernel<16, true><<<g, 256, 32768>>>(o, c); }, occ, 8, 16, kIters / 4);
        run("LDS.64", [](int g, float *o, long long *c) { lds_kernel<8, false><<<g, 256, 32768>>>(o, c); }, occ, 8, 8, kIters / 4);
    }
    return 0;
}

// Tensor-core path of the dense-design SPARC products (sm_100a: tcgen05 + TMEM + TMA).
//
//   C[f, r] = sum_k W[r, k] * X[f, k]        W = A (r = row of A, k < ML)    -> A beta   (sparc_new.py:903)
//                                            W = A^T (r = column of A, k < n) -> A^T z   (sparc_new.py:907)
// for a batch of F frames that share the design matrix.  The reference computes these
// in float64; north_star asks for 1e-4 relative on tau^2 / beta, which one 16-bit operand
// (8 or 11 mantissa bits) cannot give.  Each operand is therefore split into two fp16 terms
// (x = hi + lo, 22 mantissa bits; the design is stored times a power of two so that its lo
// terms stay normal numbers) and three MMAs accumulate into the same fp32 TMEM tile:
// W_hi X_hi + W_lo X_hi + W_hi X_lo  (the lo*lo term is below 2^-22 relative).  The tensor
// core rounds its accumulator toward zero, so K is accumulated in TMEM 8 K blocks at a time
// and the chunks are summed in registers with round-to-nearest (tc_gemm.cuh, kChunkKb).
//
// Kernel anatomy (persistent, one CTA per SM; a work item is one 128 x 128 output tile over
// one K range, K streamed in 64-element blocks; K is split when the tile count alone would
// leave the last wave mostly idle, partials summed in fixed order by reduce_splits_kernel):
//   warp 0, one lane : TMA producer  - 4 tiles (W_hi, W_lo, X_hi, X_lo; 128 rows x 128 B,
//                      SWIZZLE_128B) per stage, 3-stage ring, mbarrier expect_tx
//   warp 1, one lane : MMA issuer    - 4 k-steps x 3 tcgen05.mma.cta_group::1.kind::f16
//                      (M = 128, N = 128, K = 16) per stage, tcgen05.commit frees the stage
//   warp 2           : TMEM allocator (ring of 4 accumulators x 128 fp32 columns)
//   warps 4..7       : epilogue      - tcgen05.ld 32x32b.x32, transposed store C[f, r]
//                      (lanes = consecutive r: coalesced); takes over the accumulator of K chunk j
//                      (summing the chunks of an item in registers) while the MMAs of the next
//                      chunks fill the other three
#include "tc_gemm.cuh"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <new>


// One operand W [rows, K] kept as fp16 hi/lo + its TMA maps
struct TcOperand {
    tc_t *hi = nullptr, *lo = nullptr;
    CUtensorMap map_hi, map_lo;
    int64_t rows = 0, K = 0;
};

struct lsb_tc_design {
    int device = 0;
    int64_t n = 0, ML = 0;
    // The stored operands are A * 2^shift (exact), shift chosen so that the largest entry lands in [8, 16): the lo
    // terms of an fp16 pair then stay in the normal range (an i.i.d. N(0, 1/n) design has entries ~ 0.01, whose lo
    // terms would be fp16 subnormals with 10 bits fewer).  Epilogues multiply the accumulators by out_scale = 2^-shift.
    float out_scale = 1.0f;
    TcOperand A, At;              // A: [n, ML] (K = ML) for A beta;  At: [ML, n] (K = n) for A^T z
    std::mutex mu;
    // Staging (split frame operand, K-split partials) is kept per stream: calls on one stream are
    // ordered by the stream itself, calls on different streams (two half-batches whose element-wise
    // kernels overlap each other's GEMMs) never share a buffer.
    struct Staging {
        tc_t *x_hi = nullptr, *x_lo = nullptr;
        size_t x_cap = 0;
        float *part = nullptr;
        size_t part_cap = 0;
        // fused AMP chain (lsb_tc_amp_dev_f32): z and its fp16 pair, beta's fp16 pair, scalars
        float *z = nullptr, *bsq_part = nullptr;
        tc_t *z_hi = nullptr, *z_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
        double *tau = nullptr;
        size_t amp_frames = 0;
    };
    std::map<cudaStream_t, Staging> staging;
    // host-pointer entry (lsb_tc_amp_host_f32): its own stream and staging, one call at a time
    std::mutex host_mu;
    cudaStream_t host_stream = nullptr;
    void *host_stage = nullptr;
    size_t host_stage_bytes = 0;
};

namespace {

// K splits of a product with n_tiles output tiles and num_kb K blocks: few tiles (A beta: n/128 x F/128) leave most
// of the last wave idle; pick the split count that minimises rounds-per-split, charging 3% per extra partial
int pick_splits(int n_tiles, int num_kb, int sms)
{
    int splits = 1;
    double best = 1e30;
    for (int sp = 1; sp <= 8 && sp * 8 <= num_kb; ++sp) {
        const double cost = (double)((n_tiles * sp + sms - 1) / sms) / sp + 0.03 * (sp - 1);
        if (cost < best - 1e-9) {
            best = cost;
            splits = sp;
        }
    }
    return splits;
}

int build_operand(TcOperand *op, const double *d_src, int64_t rows, int64_t K, double pre)
{
    const size_t count = (size_t)rows * K;
    // TMA needs the row pitch to be a multiple of 16 bytes
    LSB_REQUIRE((K * sizeof(tc_t)) % 16 == 0, "inner dimension %lld must be a multiple of 8", (long long)K);
    LSB_CUDA(cudaMalloc((void **)&op->hi, count * sizeof(tc_t)));
    LSB_CUDA(cudaMalloc((void **)&op->lo, count * sizeof(tc_t)));
    split_kernel<double><<<lsb::sm_count() * 8, 256>>>(d_src, op->hi, op->lo, (long long)count, pre);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    op->rows = rows;
    op->K = K;
    int rc = make_map(&op->map_hi, op->hi, rows, K);
    if (!rc) rc = make_map(&op->map_lo, op->lo, rows, K);
    return rc;
}

int tc_product(lsb_tc_design *g, const TcOperand &W, const float *x, float *out, int64_t F, cudaStream_t st)
{
    if (F == 0) return LSB_OK;                     // an empty batch is valid with any pointers
    LSB_REQUIRE(g && x && out, "null pointer");
    LSB_REQUIRE(F >= 0, "negative frame count");
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    std::unique_lock<std::mutex> lock(g->mu);
    LSB_REQUIRE(g->staging.size() < 32 || g->staging.count(st), "more than 32 streams use this design");
    lsb_tc_design::Staging &sg = g->staging[st];
    lock.unlock();                                       // entries are never erased; each stream touches only its own
    const size_t need = (size_t)F * (size_t)(g->n > g->ML ? g->n : g->ML);
    if (sg.x_cap < need) {
        LSB_CUDA(cudaStreamSynchronize(st));
        cudaFree(sg.x_hi);
        cudaFree(sg.x_lo);
        sg.x_hi = sg.x_lo = nullptr;
        sg.x_cap = 0;
        LSB_CUDA(cudaMalloc((void **)&sg.x_hi, need * sizeof(tc_t)));
        LSB_CUDA(cudaMalloc((void **)&sg.x_lo, need * sizeof(tc_t)));
        sg.x_cap = need;
    }
    const long long count = (long long)F * W.K;
    if (count % 8 == 0 && ((uintptr_t)x & 15) == 0)
        split8_kernel<<<lsb::sm_count() * 8, 256, 0, st>>>((const float4 *)x, (uint4 *)sg.x_hi, (uint4 *)sg.x_lo, count / 8);
    else
        split_kernel<float><<<lsb::sm_count() * 4, 256, 0, st>>>(x, sg.x_hi, sg.x_lo, count, 1.0);
    lsb::count_launch();
    CUtensorMap mx_hi, mx_lo;
    int rc = make_map(&mx_hi, sg.x_hi, (uint64_t)F, (uint64_t)W.K);
    if (!rc) rc = make_map(&mx_lo, sg.x_lo, (uint64_t)F, (uint64_t)W.K);
    if (rc) return rc;
    LSB_CUDA(lsb::ensure_dynamic_smem(tc_gemm_kernel, kSmemBytes));      // tracked per device
    const int n_row_tiles = (int)((W.rows + kTileRows - 1) / kTileRows), n_f_tiles = (int)((F + kTileN - 1) / kTileN);
    const int n_tiles = n_row_tiles * n_f_tiles, sms = lsb::sm_count();
    const int num_kb = (int)((W.K + kBlockK - 1) / kBlockK);
    int splits = pick_splits(n_tiles, num_kb, sms);
    const size_t out_count = (size_t)F * (size_t)W.rows;
    if (((uintptr_t)out & 15) != 0) splits = 1;     // the partial sum runs on float4
    float *dst = out;
    if (splits > 1) {
        const size_t need_part = out_count * (size_t)splits;
        if (sg.part_cap < need_part) {
            LSB_CUDA(cudaStreamSynchronize(st));
            cudaFree(sg.part);
            sg.part = nullptr;
            sg.part_cap = 0;
            LSB_CUDA(cudaMalloc((void **)&sg.part, need_part * sizeof(float)));
            sg.part_cap = need_part;
        }
        dst = sg.part;
    }
    const int n_items = n_tiles * splits;
    tc_gemm_kernel<<<n_items < sms ? n_items : sms, kThreadsTc, kSmemBytes, st>>>(
        W.map_hi, W.map_lo, mx_hi, mx_lo, dst, (int)W.rows, (int)F, (int)W.K, n_row_tiles, n_f_tiles, splits,
        (long long)out_count, g->out_scale);
    if (splits > 1) {
        lsb::count_launch();
        LSB_CUDA(cudaGetLastError());
        reduce_splits_kernel<<<sms * 8, 256, 0, st>>>((const float4 *)sg.part, (float4 *)out, (long long)(out_count / 4),
                                                     splits);
    }
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

template <int M>
void launch_denoise_gemm(int grid, cudaStream_t st, const CUtensorMap &w_hi, const CUtensorMap &w_lo, const CUtensorMap &x_hi,
                         const CUtensorMap &x_lo, float *beta, int has_beta, tc_t *b_hi, tc_t *b_lo,
                         float *s_out, const double *tau, float *bsq_part, float sq, int rows, int F, int K, int n_row_tiles,
                         int n_f_tiles, float out_scale, int32_t *idx_out)
{
    lsb::ensure_dynamic_smem(tc_gemm_denoise_kernel<M>, kSmemBytes);
    tc_gemm_denoise_kernel<M><<<grid, kThreadsTc, kSmemBytes, st>>>(w_hi, w_lo, x_hi, x_lo, beta, has_beta, b_hi, b_lo, s_out,
                                                                    tau, bsq_part, sq, rows, F, K, n_row_tiles, n_f_tiles,
                                                                    out_scale, idx_out);
}

}  // namespace

extern "C" {

// Whole dense-design AMP decode (sparc_new.py:885-912) for F frames sharing the design, three launches per
// iteration: A beta (K-split partials) -> tc_residual_kernel -> A^T z with the denoiser as its epilogue.
// beta_out [F, ML] is the iterate itself (zeroed here, soft estimate on return); s_out (optional) receives the
// last iteration's s = beta + A^T z; tau_trace (optional) [t_max, F] the tau^2 of every iteration.
int lsb_tc_amp_dev_f32(lsb_tc_design *g, const float *y, int64_t F, int L, int M, double P, double sqrt_nPl, int t_max,
                       float *beta_out, float *s_out, double *tau_trace, int32_t *idx_out, void *stream)
{
    if (F == 0) return LSB_OK;
    LSB_REQUIRE(g && y && beta_out, "null pointer");
    LSB_REQUIRE(F > 0 && F < (1 << 24) && t_max >= 1, "bad frame count or t_max");
    LSB_REQUIRE((int64_t)L * M == g->ML, "L * M = %lld does not match the design (ML = %lld)", (long long)L * M,
                (long long)g->ML);
    LSB_REQUIRE(M >= 2 && M <= 128 && (M & (M - 1)) == 0, "the fused denoiser handles M = 2 .. 128 (power of two), got %d", M);
    LSB_REQUIRE(g->ML % 32 == 0, "the fused denoiser needs L * M to be a multiple of 32 (whole 32-column accumulator chunks)");
    LSB_REQUIRE(((uintptr_t)y & 15) == 0 && ((uintptr_t)beta_out & 15) == 0 && (!s_out || ((uintptr_t)s_out & 15) == 0),
                "y, beta_out and s_out must be 16-byte aligned");
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    cudaStream_t st = (cudaStream_t)stream;
    std::unique_lock<std::mutex> lock(g->mu);
    LSB_REQUIRE(g->staging.size() < 32 || g->staging.count(st), "more than 32 streams use this design");
    lsb_tc_design::Staging &sg = g->staging[st];
    lock.unlock();

    const int n = (int)g->n, ML = (int)g->ML, sms = lsb::sm_count();
    const int n_f_tiles = (int)((F + kTileN - 1) / kTileN);
    const int ab_row_tiles = (n + kTileRows - 1) / kTileRows, az_row_tiles = (ML + kTileRows - 1) / kTileRows;
    const int splits = pick_splits(ab_row_tiles * n_f_tiles, (ML + kBlockK - 1) / kBlockK, sms);
    if (sg.amp_frames < (size_t)F) {
        LSB_CUDA(cudaStreamSynchronize(st));
        cudaFree(sg.z); cudaFree(sg.bsq_part); cudaFree(sg.z_hi); cudaFree(sg.z_lo); cudaFree(sg.b_hi); cudaFree(sg.b_lo);
        cudaFree(sg.tau);
        sg.z = sg.bsq_part = nullptr;
        sg.z_hi = sg.z_lo = sg.b_hi = sg.b_lo = nullptr;
        sg.tau = nullptr;
        sg.amp_frames = 0;
        LSB_CUDA(cudaMalloc((void **)&sg.z, (size_t)F * n * sizeof(float)));
        LSB_CUDA(cudaMalloc((void **)&sg.z_hi, (size_t)F * n * sizeof(tc_t)));
        LSB_CUDA(cudaMalloc((void **)&sg.z_lo, (size_t)F * n * sizeof(tc_t)));
        LSB_CUDA(cudaMalloc((void **)&sg.b_hi, (size_t)F * ML * sizeof(tc_t)));
        LSB_CUDA(cudaMalloc((void **)&sg.b_lo, (size_t)F * ML * sizeof(tc_t)));
        LSB_CUDA(cudaMalloc((void **)&sg.bsq_part, (size_t)F * az_row_tiles * sizeof(float)));
        LSB_CUDA(cudaMalloc((void **)&sg.tau, (size_t)F * sizeof(double)));
        sg.amp_frames = (size_t)F;
    }
    const size_t out_count = (size_t)F * (size_t)n;
    if (sg.part_cap < out_count * (size_t)splits) {
        LSB_CUDA(cudaStreamSynchronize(st));
        cudaFree(sg.part);
        sg.part = nullptr;
        sg.part_cap = 0;
        LSB_CUDA(cudaMalloc((void **)&sg.part, out_count * (size_t)splits * sizeof(float)));
        sg.part_cap = out_count * (size_t)splits;
    }
    CUtensorMap mb_hi, mb_lo, mz_hi, mz_lo;
    int rc = make_map(&mb_hi, sg.b_hi, (uint64_t)F, (uint64_t)ML);
    if (!rc) rc = make_map(&mb_lo, sg.b_lo, (uint64_t)F, (uint64_t)ML);
    if (!rc) rc = make_map(&mz_hi, sg.z_hi, (uint64_t)F, (uint64_t)n);
    if (!rc) rc = make_map(&mz_lo, sg.z_lo, (uint64_t)F, (uint64_t)n);
    if (rc) return rc;
    LSB_CUDA(lsb::ensure_dynamic_smem(tc_gemm_kernel, kSmemBytes));

    const int ab_items = ab_row_tiles * n_f_tiles * splits, az_items = az_row_tiles * n_f_tiles;
    for (int t = 0; t < t_max; ++t) {
        if (t > 0) {
            tc_gemm_kernel<<<ab_items < sms ? ab_items : sms, kThreadsTc, kSmemBytes, st>>>(
                g->A.map_hi, g->A.map_lo, mb_hi, mb_lo, sg.part, n, (int)F, ML, ab_row_tiles, n_f_tiles, splits,
                (long long)out_count, g->out_scale);
            lsb::count_launch();
        }
        tc_residual_kernel<<<(unsigned)F, kThreadsTc, 0, st>>>(y, sg.part, splits, (long long)out_count, sg.z, sg.z_hi, sg.z_lo,
                                                             sg.bsq_part, az_row_tiles, sg.tau,
                                                             tau_trace ? tau_trace + (size_t)t * F : nullptr, (int)F, n, P,
                                                             t == 0);
        lsb::count_launch();
        float *so = (t == t_max - 1) ? s_out : nullptr;
        int32_t *io = (t == t_max - 1) ? idx_out : nullptr;
        const int grid = az_items < sms ? az_items : sms;
#define LSB_DN(MM)                                                                                                       \
    case MM:                                                                                                             \
        launch_denoise_gemm<MM>(grid, st, g->At.map_hi, g->At.map_lo, mz_hi, mz_lo, beta_out, t > 0, sg.b_hi, sg.b_lo, so, \
                                sg.tau, sg.bsq_part, (float)sqrt_nPl, ML, (int)F, n, az_row_tiles, n_f_tiles,            \
                                g->out_scale, io);                                                                       \
        break;
        switch (M) {
            LSB_DN(2) LSB_DN(4) LSB_DN(8) LSB_DN(16) LSB_DN(32) LSB_DN(64) LSB_DN(128)
        }
#undef LSB_DN
        lsb::count_launch();
        LSB_CUDA(cudaGetLastError());
    }
    return LSB_OK;
}

/* Host-pointer form: y [F, n] in host memory (pinned memory makes the copies asynchronous) -> MAP section decisions
 * idx [F, L] (argmax of the last iteration's s, sparc_new.py:79-80 / 1099-1116) back in host memory on return. */
int lsb_tc_amp_host_f32(lsb_tc_design *g, const float *y, int64_t F, int L, int M, double P, double sqrt_nPl, int t_max,
                        int32_t *idx)
{
    if (F == 0) return LSB_OK;
    LSB_REQUIRE(g && y && idx && F > 0, "null pointer or negative frame count");
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    std::lock_guard<std::mutex> lock(g->host_mu);
    if (!g->host_stream) LSB_CUDA(cudaStreamCreateWithFlags(&g->host_stream, cudaStreamNonBlocking));
    const size_t by = (size_t)F * g->n * sizeof(float), bb = (size_t)F * g->ML * sizeof(float), bi = (size_t)F * L * sizeof(int32_t);
    const size_t need = ((by + 255) & ~(size_t)255) + ((bb + 255) & ~(size_t)255) + bi;
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
    float *d_beta = (float *)(base + ((by + 255) & ~(size_t)255));
    int32_t *d_idx = (int32_t *)((char *)d_beta + ((bb + 255) & ~(size_t)255));
    cudaStream_t st = g->host_stream;
    LSB_CUDA(cudaMemcpyAsync(d_y, y, by, cudaMemcpyHostToDevice, st));
    int rc = lsb_tc_amp_dev_f32(g, d_y, F, L, M, P, sqrt_nPl, t_max, d_beta, nullptr, nullptr, d_idx, st);
    if (rc) return rc;
    LSB_CUDA(cudaMemcpyAsync(idx, d_idx, bi, cudaMemcpyDeviceToHost, st));
    LSB_CUDA(cudaStreamSynchronize(st));
    return LSB_OK;
}

int lsb_tc_design_create(const double *A, int64_t n, int64_t ML, lsb_tc_design **out)
{
    LSB_REQUIRE(A && out, "null pointer");
    LSB_REQUIRE(n > 0 && ML > 0 && n % 8 == 0 && ML % 8 == 0, "n and ML must be positive multiples of 8");
    lsb_tc_design *g = new (std::nothrow) lsb_tc_design();
    if (!g) return LSB_ERR_NOMEM;
    g->n = n;
    g->ML = ML;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device");
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    const size_t count = (size_t)n * (size_t)ML;
    double *d_A = nullptr, *d_At = nullptr;
    int rc = LSB_OK;
    cudaError_t e = cudaMalloc((void **)&d_A, count * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_At, count * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_A, A, count * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dim3 grid((unsigned)((ML + 31) / 32), (unsigned)((n + 31) / 32)), block(32, 8);
        transpose_kernel<<<grid, block>>>(d_A, d_At, (int)n, (int)ML);
        lsb::count_launch();
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        lsb::set_error("uploading the design matrix: %s", cudaGetErrorString(e));
        rc = LSB_ERR_CUDA;
    }
    double pre = 1.0;
#if LSB_TC_HALF
    {
        double mx = 0.0;
        for (size_t i = 0; i < count; ++i) mx = std::max(mx, std::fabs(A[i]));
        if (mx > 0.0 && std::isfinite(mx)) {
            int e = 0;
            std::frexp(mx, &e);                       // mx = m 2^e, m in [0.5, 1)
            const int shift = std::min(24, std::max(-24, 4 - e));     // mx 2^shift in [8, 16)
            pre = std::ldexp(1.0, shift);
        }
    }
#endif
    g->out_scale = (float)(1.0 / pre);
    if (!rc) rc = build_operand(&g->A, d_A, n, ML, pre);
    if (!rc) rc = build_operand(&g->At, d_At, ML, n, pre);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) {
        lsb::set_error("building the fp16 operands: %s", cudaGetErrorString(cudaGetLastError()));
        rc = LSB_ERR_CUDA;
    }
    cudaFree(d_A);
    cudaFree(d_At);
    if (rc) {
        lsb_tc_design_destroy(g);
        return rc;
    }
    *out = g;
    return LSB_OK;
}

void lsb_tc_design_destroy(lsb_tc_design *g)
{
    if (!g) return;
    cudaDeviceSynchronize();
    cudaFree(g->A.hi);
    cudaFree(g->A.lo);
    cudaFree(g->At.hi);
    cudaFree(g->At.lo);
    for (auto &kv : g->staging) {
        cudaFree(kv.second.x_hi);
        cudaFree(kv.second.x_lo);
        cudaFree(kv.second.part);
        cudaFree(kv.second.z);
        cudaFree(kv.second.bsq_part);
        cudaFree(kv.second.z_hi);
        cudaFree(kv.second.z_lo);
        cudaFree(kv.second.b_hi);
        cudaFree(kv.second.b_lo);
        cudaFree(kv.second.tau);
    }
    cudaFree(g->host_stage);
    if (g->host_stream) cudaStreamDestroy(g->host_stream);
    delete g;
}

int lsb_tc_ab_dev_f32(lsb_tc_design *g, const float *x, float *out, int64_t F, void *stream)
{
    LSB_REQUIRE(g, "null design");
    return tc_product(g, g->A, x, out, F, (cudaStream_t)stream);
}

int lsb_tc_az_dev_f32(lsb_tc_design *g, const float *z, float *out, int64_t F, void *stream)
{
    LSB_REQUIRE(g, "null design");
    return tc_product(g, g->At, z, out, F, (cudaStream_t)stream);
}

}  // extern "C"

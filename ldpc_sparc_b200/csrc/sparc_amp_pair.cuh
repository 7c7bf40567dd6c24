// fp32 throughput kernels on the pair transforms of sparc_pair.cuh: AMP decoder and Ab / Az operators.
// Included by sparc_amp.cu inside its anonymous namespace (uses DesignDev, AmpParams, AmpOut, warp_sum).
//
// amp_pair_kernel: one CTA decodes TWO frames in lockstep (reference sparc_public/sparc.py:883-999,
// same iteration as amp_kernel above).  Differences from the one-frame kernel that matter for the
// machine:
//  * beta lives in a per-CTA global scratch as float2 (frame A, frame B) per entry -- half of what the
//    one-frame kernel kept (it also stored s every iteration; here s only exists in registers and the
//    MAP decision argmax(s) is taken inside the fused softmax of every iteration, overwriting the
//    frame's decision row while the frame is active), so the scratch of all resident CTAs stays in L2;
//  * psi / nmse partial sums go through fixed-order per-warp slots (no atomics): results are
//    bit-reproducible run to run;
//  * frames stop independently: a stopped frame's outputs are frozen, the pair ends when both stopped.
#pragma once
#include "sparc_pair.cuh"

#ifndef LSB_AMP_PAIR_OCC
#define LSB_AMP_PAIR_OCC 3
#endif

using lsb_sparc::pairfft::CP;
using lsb_sparc::pairfft::P;
using lsb_sparc::pairfft::PairBlock;
using lsb_sparc::pairfft::PairTables;
using lsb_sparc::pairfft::kPairThreads;

struct PairDev {
    const PairBlock *blocks;       // same order as DesignDev::blocks (sorted by column, then row)
    PairTables tb;
};

struct SrcPair {                   // column block stored as float2 (A, B) per entry
    const float2 *p;
    __device__ __forceinline__ P get(unsigned i) const { return __ldcg(p + i); }     // written by this CTA: L2, no L1 allocation
};
struct SrcStage {                  // column block staged in shared memory (BSTAGE variant of the AMP kernel)
    const float2 *p;
    __device__ __forceinline__ P get(unsigned i) const { return p[i]; }
};
struct SrcTwo {                    // the two frames' column blocks in two separate arrays
    const float *a, *b;
    __device__ __forceinline__ P get(unsigned i) const { return make_float2(__ldg(a + i), __ldg(b + i)); }
};

__device__ __forceinline__ float fast_ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// per-CTA scalar state of the two frames, carved from dynamic shared memory behind the residual
// (sized by the design's Lr / Lc, so that three CTAs leave the L1 room for the transform tables)
struct PairScalars {                          // accessors instead of eleven live pointers: the kernel runs at its register limit
    double *dp;
    int Lc, Lr;
    __device__ __forceinline__ double *psi() const { return dp; }                  // [2][Lc]
    __device__ __forceinline__ double *psi_prev() const { return dp + 2 * Lc; }
    __device__ __forceinline__ double *tau() const { return dp + 4 * Lc; }
    __device__ __forceinline__ double *nm() const { return dp + 6 * Lc; }
    __device__ __forceinline__ double *phi() const { return dp + 8 * Lc; }         // [2][Lr]
    __device__ __forceinline__ double *gam() const { return dp + 8 * Lc + 2 * Lr; }
    __device__ __forceinline__ double *zsq() const { return dp + 8 * Lc + 4 * Lr; }
    __device__ __forceinline__ float2 *fbase() const { return reinterpret_cast<float2 *>(dp + 8 * Lc + 6 * Lr); }
    __device__ __forceinline__ float2 *tau_f() const { return fbase(); }           // [Lc] tau_c, log2(e) / tau_c          (A, B)
    __device__ __forceinline__ float2 *itau_f() const { return fbase() + Lc; }
    __device__ __forceinline__ float2 *iphi_f() const { return fbase() + 2 * Lc; } // [Lr] 1 / phi_r, gamma_r / phi_prev_r (A, B)
    __device__ __forceinline__ float2 *coef_f() const { return fbase() + 2 * Lc + Lr; }
};
__host__ __device__ inline size_t pair_scalar_bytes(int Lr, int Lc)
{
    return (size_t)(4 * 2 * Lc + 3 * 2 * Lr) * sizeof(double) + (size_t)(2 * Lc + 2 * Lr) * sizeof(float2);
}
constexpr int kBetaStageBytes = 8192;     // BSTAGE variant of the AMP kernel: one column block of beta (float2 per entry)
// FFT buffer, table staging area (sparc_pair.cuh: LSB_PAIR_STAGE), residual
__host__ __device__ inline size_t pair_vec_bytes(int h, int n)
{
    return ((size_t)16 << h) + (h == 11 ? (size_t)16 * pairfft::stage_slots<11>() : 0) + (((size_t)n * sizeof(float2) + 15) & ~(size_t)15);
}
size_t pair_op_dyn_smem(int h, int n) { return pair_vec_bytes(h, n); }
size_t pair_amp_dyn_smem(int h, int n, int Lr, int Lc, bool bstage)
{
    return pair_vec_bytes(h, n) + (bstage ? ((pair_scalar_bytes(Lr, Lc) + 15) & ~(size_t)15) + kBetaStageBytes : pair_scalar_bytes(Lr, Lc));
}

// ------------------------------------------------------------------ operators
// out[f, n] = A x[f, LM] for frame pairs (reference Ab, sparc.py:786-789,811-819,851-862)
template <int LOGH>
__global__ void __launch_bounds__(kPairThreads)
ab_pair_kernel(DesignDev d, PairDev pd, const float *__restrict__ x, float *__restrict__ out, long long n_frames)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int H = 1 << LOGH;
    float4 *buf = reinterpret_cast<float4 *>(smem);
    float2 *acc2 = reinterpret_cast<float2 *>(buf + H + pairfft::stage_slots<LOGH>());       // [n]
    const int tid = threadIdx.x;
    const long long LM = (long long)d.L * d.M;
    for (long long fA = 2LL * blockIdx.x; fA < n_frames; fA += 2LL * gridDim.x) {
        const long long fB = fA + 1 < n_frames ? fA + 1 : fA;
        for (int i = tid; i < d.n; i += kPairThreads) acc2[i] = make_float2(0.f, 0.f);
        PairBlock b = pd.blocks[0];
        pairfft::FwdIn<LOGH> fin = pairfft::forward_issue<LOGH>(pd.tb, b, tid, buf);
        for (int bi = 0; bi < d.n_blocks; ++bi) {
            const size_t off = (size_t)b.c * d.Mc;
            const pairfft::GatPre g =
                pairfft::forward_run<LOGH>(buf, pd.tb, b, SrcTwo{x + fA * LM + off, x + fB * LM + off}, fin, tid);
            PairBlock bn = b;
            if (bi + 1 < d.n_blocks) {
                bn = pd.blocks[bi + 1];
                fin = pairfft::forward_issue<LOGH>(pd.tb, bn, tid, buf);
            }
            pairfft::gather_add(acc2 + b.r * d.Mr, buf, g, b.fgat, b.scale, b.gat_chunks, tid);
            __syncthreads();
            b = bn;
        }
        for (int i = tid; i < d.n; i += kPairThreads) {
            const float2 v = acc2[i];
            out[fA * d.n + i] = v.x;
            if (fB != fA) out[fB * d.n + i] = v.y;
        }
        __syncthreads();
    }
}

// out[f, LM] = A^T y[f, n] for frame pairs (reference Az, sparc.py:791-794,821-829,864-875)
template <int LOGH>
__global__ void __launch_bounds__(kPairThreads)
az_pair_kernel(DesignDev d, PairDev pd, const float *__restrict__ y, float *__restrict__ out, long long n_frames)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int H = 1 << LOGH;
    float4 *buf = reinterpret_cast<float4 *>(smem);
    float2 *y2 = reinterpret_cast<float2 *>(buf + H + pairfft::stage_slots<LOGH>());         // [n]
    const int tid = threadIdx.x;
    const long long LM = (long long)d.L * d.M;
    for (long long fA = 2LL * blockIdx.x; fA < n_frames; fA += 2LL * gridDim.x) {
        const long long fB = fA + 1 < n_frames ? fA + 1 : fA;
        for (int i = tid; i < d.n; i += kPairThreads) y2[i] = make_float2(__ldg(y + fA * d.n + i), __ldg(y + fB * d.n + i));
        __syncthreads();
        for (int c = 0; c < d.Lc; ++c) {
            float *oA = out + fA * LM + (size_t)c * d.Mc, *oB = out + fB * LM + (size_t)c * d.Mc;
            const int b0 = d.col_first[c], b1 = d.col_first[c + 1];
            for (int bi = b0; bi < b1; ++bi) {
                const PairBlock b = pd.blocks[bi];
                const pairfft::TrIn tin = pairfft::transposed_issue<LOGH>(pd.tb, b, tid, buf);
                pairfft::transposed_run<LOGH>(buf, pd.tb, b, y2 + b.r * d.Mr, tin, tid, [] { return 0; });
                const float2 *bf2 = reinterpret_cast<const float2 *>(buf);
                for (int ch = 0; ch < pairfft::chunk_count(d.Mc); ++ch) {
                    const uint2 e = pairfft::ldg_na(b.tgat + ch * kPairThreads + tid);
                    const unsigned id[4] = {e.x & 0xFFFFu, e.x >> 16, e.y & 0xFFFFu, e.y >> 16};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = (ch * kPairThreads + tid) * 4 + u;          // tgat: four consecutive columns per thread
                        if (j >= d.Mc) continue;
                        const float2 v = bf2[id[u]];
                        const float va = b.scale * v.x, vb = b.scale * v.y;
                        if (bi == b0) {
                            oA[j] = va;
                            if (fB != fA) oB[j] = vb;
                        } else {
                            oA[j] += va;
                            if (fB != fA) oB[j] += vb;
                        }
                    }
                }
                __syncthreads();
            }
            if (b0 == b1)
                for (int j = tid; j < d.Mc; j += kPairThreads) {
                    oA[j] = 0.f;
                    if (fB != fA) oB[j] = 0.f;
                }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ AMP
// Base matrix, column offsets and block descriptors of the design, copied into shared memory at kernel start when
// they fit (CACHED): the block loop reads the next descriptor and immediately uses its table pointers, the column
// loop branches on col_first, the gamma / tau loops walk W -- from global memory each is an exposed L2 round trip
// (ncu: long-scoreboard stalls at these loads), from shared memory a 30-cycle load.
// (Static arrays: their addresses are compile-time constants, so the cache costs no registers; a pointer into dynamic
// shared memory behind the run-time sized vectors did, and the kernel spilled.)
#ifndef LSB_PAIR_CACHE
#define LSB_PAIR_CACHE 1
#endif
constexpr int kCacheBlocks = 48, kCacheCols = 63, kCacheW = 64;
inline bool pair_design_cached(int Lr, int Lc, int n_blocks) { return n_blocks <= kCacheBlocks && Lc <= kCacheCols && Lr * Lc <= kCacheW; }

struct OutPre {                    // first chunk of the gather after a transposed transform: 4 slots + the 4 addends
    uint2 ids;
    float2 pv[4];
};
__device__ __forceinline__ void ld4_cg(const float2 *p, float2 *o)          // four consecutive float2 (32-byte aligned), L2 only
{
    const float4 a = __ldcg(reinterpret_cast<const float4 *>(p)), b = __ldcg(reinterpret_cast<const float4 *>(p) + 1);
    o[0] = make_float2(a.x, a.y); o[1] = make_float2(a.z, a.w); o[2] = make_float2(b.x, b.y); o[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void st4_cg(float2 *p, const float2 *v)
{
    __stcg(reinterpret_cast<float4 *>(p), make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
    __stcg(reinterpret_cast<float4 *>(p) + 1, make_float4(v[2].x, v[2].y, v[3].x, v[3].y));
}

// Requirements checked by the host: fp32, unmodulated (K = 1), 4 <= M <= 32, every column block has at least
// one non-zero base-matrix entry, no soft outputs (beta_soft / s_last), H = 2^LOGH.
// scratch per CTA: [LM] float2 beta (A, B) then [Mc] float2 partial s of the current column.
// BSTAGE: the column block of beta that the NEXT forward transform gathers from travels L2 -> shared memory with
// cp.async (issued next to that transform's index table, one row gather ahead) into 8 KB behind the scalar state, so the
// first pass gathers from shared memory instead of waiting for 16 random L2 loads per thread.  Needs Mc <= 1024 and
// runs 2 CTAs per SM (the third CTA's shared memory pays for the buffer; the 2-CTA build is as fast).
#ifndef LSB_PAIR_BSTAGE
#define LSB_PAIR_BSTAGE 1
#endif
template <int LOGH, bool CACHED, bool BSTAGE>
__global__ void __launch_bounds__(kPairThreads, BSTAGE ? 2 : LSB_AMP_PAIR_OCC)
amp_pair_kernel(DesignDev d, PairDev pd, AmpParams p, const float *__restrict__ y_all,
                const int32_t *__restrict__ idx0, AmpOut o, float2 *__restrict__ scratch, long long n_frames,
                unsigned long long *__restrict__ queue)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int H = 1 << LOGH;
    constexpr int kWarps = kPairThreads / 32;
    float4 *buf = reinterpret_cast<float4 *>(smem);           // [H] (re_A, re_B, im_A, im_B)
    float2 *z2 = reinterpret_cast<float2 *>(buf + H + pairfft::stage_slots<LOGH>());         // [n] residual of frame A / frame B
    const PairScalars S{reinterpret_cast<double *>(smem + pair_vec_bytes(LOGH, d.n)), d.Lc, d.Lr};
    float4 *bstage = reinterpret_cast<float4 *>(smem + pair_vec_bytes(LOGH, d.n) + ((pair_scalar_bytes(d.Lr, d.Lc) + 15) & ~(size_t)15));
    // thread t copies 16-byte pieces t and 256 + t of column block c (Mc float2 = Mc / 2 pieces)
    auto beta_issue = [&](int c) {
        const float4 *src = reinterpret_cast<const float4 *>(scratch + (size_t)blockIdx.x * ((size_t)d.L * d.M + d.Mc) + (size_t)c * d.Mc);
        const int pieces = d.Mc >> 1, t = threadIdx.x;
        if (t < pieces)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(bstage + t)), "l"(src + t) : "memory");
        if (t + kPairThreads < pieces)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(bstage + kPairThreads + t)),
                         "l"(src + kPairThreads + t)
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __shared__ __align__(16) PairBlock c_blocks[CACHED ? kCacheBlocks : 1];
    __shared__ int c_col_first[CACHED ? kCacheCols + 1 : 1];
    __shared__ double c_W[CACHED ? kCacheW : 1];
    if (CACHED) {
        for (int i = threadIdx.x; i < d.Lr * d.Lc; i += kPairThreads) c_W[i] = d.W[i];
        for (int i = threadIdx.x; i <= d.Lc; i += kPairThreads) c_col_first[i] = d.col_first[i];
        const int words = d.n_blocks * (int)(sizeof(PairBlock) / sizeof(uint2));
        for (int i = threadIdx.x; i < words; i += kPairThreads)
            reinterpret_cast<uint2 *>(c_blocks)[i] = reinterpret_cast<const uint2 *>(pd.blocks)[i];
        __syncthreads();
    }
#define LSB_DV_BLOCKS (CACHED ? (const PairBlock *)c_blocks : pd.blocks)
#define LSB_DV_COLS (CACHED ? (const int *)c_col_first : d.col_first)
#define LSB_DV_W (CACHED ? (const double *)c_W : d.W)
    __shared__ double part_psi[2][kWarps][2], part_nm[2][kWarps][2];   // [column parity][warp][frame]
    __shared__ double part_z[kWarps][2];
    __shared__ long long s_frame;
    __shared__ int s_active[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = (tid >> 6) & 1, ix = tid & 63;             // scalar work: threads 0..127 = (frame, index)
    const int Lr = d.Lr, Lc = d.Lc;
    const int LM = d.L * d.M;
    float2 *beta2 = scratch + (size_t)blockIdx.x * ((size_t)LM + d.Mc);
    float2 *s2col = beta2 + LM;
    const int secs_per_col = d.Mc / d.M;
    const double sec_norm = (double)d.L / d.Lc;
    // denoiser layout: a thread holds four consecutive entries of a section (both frames), LS = M/4 lanes a section
    const int M = d.M, logM = d.logM, LS = M >> 2;
    const unsigned grp_base = (unsigned)lane & ~(unsigned)(LS - 1);
    const unsigned grp_mask = ((1u << LS) - 1u) << grp_base;
    const int col_lines = (d.Mc * (int)sizeof(float2) + 127) / 128;     // 128-byte lines of one column block of beta

    for (;;) {
        if (tid == 0) s_frame = (long long)atomicAdd(queue, 2ull);
        __syncthreads();
        const long long fA = s_frame;
        if (fA >= n_frames) break;
        const bool hasB = fA + 1 < n_frames;
        const long long fB = hasB ? fA + 1 : fA;
        const long long fsel = fr ? fB : fA;                   // frame of this thread's scalar work
        const float *yA = y_all + fA * d.n, *yB = y_all + fB * d.n;

        for (int i = tid; i < d.n; i += kPairThreads) z2[i] = make_float2(yA[i], yB[i]);      // z = y (:915)
        if (tid < 128) {
            if (ix < Lc) S.psi()[fr * Lc + ix] = 1.0;                                           // gamma = W.1/Lc (:918,926)
            if (ix < Lr) S.phi()[fr * Lr + ix] = 0.0;
        }
        if (tid < 2) s_active[tid] = tid == 0 ? 1 : (hasB ? 1 : 0);
        __syncthreads();

        int t = 0;
        for (; t < p.t_max - 1; ++t) {
            // ---- gamma_r, the Onsager coefficient gamma_r / phi_prev_r (phi still holds the previous iteration's)
            if (tid < 128) {
                if (ix < Lr) {
                    double g = 0.0;
                    for (int c = 0; c < Lc; ++c) g += LSB_DV_W[ix * Lc + c] * S.psi()[fr * Lc + c];
                    g /= Lc;
                    S.gam()[fr * Lr + ix] = g;
                    reinterpret_cast<float *>(&S.coef_f()[ix])[fr] = t > 0 ? (float)(g / S.phi()[fr * Lr + ix]) : 0.f;
                    S.zsq()[fr * Lr + ix] = 0.0;
                }
                if (ix < Lc) S.psi_prev()[fr * Lc + ix] = S.psi()[fr * Lc + ix];
            }
            __syncthreads();
            if (t > 0) {
                // z = y - A beta + b z (:944-946): z <- y + b z, then every block subtracts its A_rc beta_c.
                // (the barriers inside the first transform order this update before the first row gather)
                PairBlock b = LSB_DV_BLOCKS[0];
                pairfft::FwdIn<LOGH> fin = pairfft::forward_issue<LOGH>(pd.tb, b, tid, buf);
                if (d.n_blocks > 1 && tid < col_lines) pairfft::prefetch_l2(beta2 + (size_t)LSB_DV_BLOCKS[1].c * d.Mc + tid * 16);
                if (Lr == 1) {
                    const float2 cf = S.coef_f()[0];
                    for (int i = tid; i < d.n; i += kPairThreads) z2[i] = pairfft::pfma(z2[i], cf, make_float2(yA[i], yB[i]));
                } else {
                    for (int i = tid; i < d.n; i += kPairThreads)
                        z2[i] = pairfft::pfma(z2[i], S.coef_f()[i / d.Mr], make_float2(yA[i], yB[i]));
                }
                if (BSTAGE) {                        // (issued after the table of block 0; the update of z above ran under both)
                    beta_issue(b.c);
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncthreads();
                }
                for (int bi = 0; bi < d.n_blocks; ++bi) {
                    const pairfft::GatPre g =
                        BSTAGE ? pairfft::forward_run<LOGH>(buf, pd.tb, b, SrcStage{reinterpret_cast<const float2 *>(bstage)}, fin, tid)
                               : pairfft::forward_run<LOGH>(buf, pd.tb, b, SrcPair{beta2 + (size_t)b.c * d.Mc}, fin, tid);
                    PairBlock bn = b;
                    if (bi + 1 < d.n_blocks) {
                        bn = LSB_DV_BLOCKS[bi + 1];
                        fin = pairfft::forward_issue<LOGH>(pd.tb, bn, tid, buf);
                        if (BSTAGE) beta_issue(bn.c);
                        // the column block after that: pull it into L2 (beta streams through DRAM between iterations)
                        if (bi + 2 < d.n_blocks && tid < col_lines)
                            pairfft::prefetch_l2(beta2 + (size_t)LSB_DV_BLOCKS[bi + 2].c * d.Mc + tid * 16);
                    }
                    pairfft::gather_add(z2 + b.r * d.Mr, buf, g, b.fgat, -b.scale, b.gat_chunks, tid);
                    if (BSTAGE) asm volatile("cp.async.wait_group 0;" ::: "memory");   // the next column block, before the barrier publishes it
                    __syncthreads();
                    b = bn;
                }
            }
            // ---- phi (:949-955)
            if (p.phi_method == 2) {
                for (int r = 0; r < Lr; ++r) {
                    double pa = 0.0, pb = 0.0;
                    for (int i = tid; i < d.Mr; i += kPairThreads) {
                        const float2 v = z2[r * d.Mr + i];
                        pa += (double)v.x * (double)v.x;
                        pb += (double)v.y * (double)v.y;
                    }
                    pa = warp_sum(pa);
                    pb = warp_sum(pb);
                    if (lane == 0) { part_z[warp][0] = pa; part_z[warp][1] = pb; }
                    __syncthreads();
                    if (tid < 2) {
                        double a = 0.0;
                        for (int w = 0; w < kWarps; ++w) a += part_z[w][tid];
                        S.zsq()[tid * Lr + r] = a;
                    }
                    __syncthreads();
                }
            }
            if (tid < 128 && ix < Lr) {
                const double ph = p.phi_method == 1 ? p.awgn_var + S.gam()[fr * Lr + ix] : S.zsq()[fr * Lr + ix] / d.Mr;
                S.phi()[fr * Lr + ix] = ph;
                reinterpret_cast<float *>(&S.iphi_f()[ix])[fr] = (float)(1.0 / ph);
            }
            __syncthreads();
            // ---- tau (:958-969)
            if (tid < 128 && ix < Lc) {
                double tc;
                if (Lr == 1) {
                    tc = ((double)d.L * S.phi()[fr * Lr] / d.n) / LSB_DV_W[ix];
                } else {
                    double acc_w = 0.0;
                    for (int r = 0; r < Lr; ++r) acc_w += LSB_DV_W[r * Lc + ix] / S.phi()[fr * Lr + r];
                    tc = ((double)d.L / d.Mr) / acc_w;
                }
                S.tau()[fr * Lc + ix] = tc;
                reinterpret_cast<float *>(&S.tau_f()[ix])[fr] = (float)tc;
                reinterpret_cast<float *>(&S.itau_f()[ix])[fr] = (float)(1.4426950408889634 / tc);
                if (o.tau_trace && s_active[fr]) o.tau_trace[(fsel * p.t_max + t) * Lc + ix] = tc;
            }
            __syncthreads();
            const bool actA = s_active[0] != 0, actB = s_active[1] != 0;
            // ---- s = beta + tau * Az(z / phi) column block by column block (:972), softmax denoiser (:429-432),
            //      MAP decision argmax(s) of this iteration (:997)
            PairBlock b = LSB_DV_BLOCKS[0];
            pairfft::TrIn tin = pairfft::transposed_issue<LOGH>(pd.tb, b, tid, buf);
            for (int c = 0; c < Lc; ++c) {
                const float2 tau2 = S.tau_f()[c], itl = S.itau_f()[c];
                float2 *beta2_c = beta2 + (size_t)c * d.Mc;
                const int b0 = LSB_DV_COLS[c], b1 = LSB_DV_COLS[c + 1];
                if (t > 0 && c + 1 < Lc && tid < col_lines) pairfft::prefetch_l2(beta2_c + d.Mc + tid * 16);
                double psA = 0.0, psB = 0.0, nmA = 0.0, nmB = 0.0;
                for (int bi = b0; bi < b1; ++bi) {
                    const float2 *prev = bi == b0 ? (t > 0 ? beta2_c : nullptr) : s2col;
                    const OutPre pre = pairfft::transposed_run<LOGH>(buf, pd.tb, b, z2 + b.r * d.Mr, tin, tid, [&] {
                        OutPre q;
                        q.ids = pairfft::ldg_na(b.tgat + tid);
                        if (prev && 4 * tid < d.Mc) ld4_cg(prev + 4 * tid, q.pv);
                        else q.pv[0] = q.pv[1] = q.pv[2] = q.pv[3] = make_float2(0.f, 0.f);
                        return q;
                    });
                    const float2 *bf2 = reinterpret_cast<const float2 *>(buf);
                    const float2 sc2 = pairfft::pmul(pairfft::pscale(tau2, b.scale), S.iphi_f()[b.r]);   // scale tau_c / phi_r
                    const uint2 *tgat = b.tgat;
                    if (bi + 1 < d.n_blocks) {                   // next block (this column or the next): start its table loads
                        b = LSB_DV_BLOCKS[bi + 1];
                        tin = pairfft::transposed_issue<LOGH>(pd.tb, b, tid, buf);
                    }
                    const bool last = bi == b1 - 1;
                    for (int base = 0, ch = 0; base < d.Mc; base += kPairThreads * 4, ++ch) {     // warp-uniform trip count
                        const int j0 = base + 4 * tid;
                        const bool live = j0 < d.Mc;               // Mc is a multiple of M >= 4: uniform over a section's lanes
                        uint2 ide = pre.ids;
                        float2 pv[4] = {pre.pv[0], pre.pv[1], pre.pv[2], pre.pv[3]};
                        if (ch > 0) {
                            ide = pairfft::ldg_na(tgat + ch * kPairThreads + tid);
                            if (prev && live) ld4_cg(prev + j0, pv);
                            else pv[0] = pv[1] = pv[2] = pv[3] = make_float2(0.f, 0.f);
                        }
                        const unsigned id[4] = {ide.x & 0xFFFFu, ide.x >> 16, ide.y & 0xFFFFu, ide.y >> 16};
                        float2 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = pairfft::pfma(bf2[live ? id[u] : 0], sc2, pv[u]);
                        if (!last) {
                            if (live) st4_cg(s2col + j0, v);
                            continue;
                        }
                        // last block of the column: softmax of the sections on the gathered values in registers
                        float mxA = live ? fmaxf(fmaxf(v[0].x, v[1].x), fmaxf(v[2].x, v[3].x)) : -INFINITY;
                        float mxB = live ? fmaxf(fmaxf(v[0].y, v[1].y), fmaxf(v[2].y, v[3].y)) : -INFINITY;
                        if (M == 32) {
#pragma unroll
                            for (int off = 4; off > 0; off >>= 1) {
                                mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, off));
                                mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, off));
                            }
                        } else {
                            for (int off = LS >> 1; off > 0; off >>= 1) {
                                mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, off));
                                mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, off));
                            }
                        }
                        // decision of this iteration: lowest index attaining the section maximum
                        const int luA = v[0].x == mxA ? 0 : v[1].x == mxA ? 1 : v[2].x == mxA ? 2 : v[3].x == mxA ? 3 : 4;
                        const int luB = v[0].y == mxB ? 0 : v[1].y == mxB ? 1 : v[2].y == mxB ? 2 : v[3].y == mxB ? 3 : 4;
                        const unsigned balA = __ballot_sync(0xffffffffu, live && luA < 4) & grp_mask;
                        const unsigned balB = __ballot_sync(0xffffffffu, live && luB < 4) & grp_mask;
                        const int srcA = balA ? __ffs(balA) - 1 : (int)grp_base, srcB = balB ? __ffs(balB) - 1 : (int)grp_base;
                        const int wuA = __shfl_sync(0xffffffffu, luA, srcA), wuB = __shfl_sync(0xffffffffu, luB, srcB);
                        // e = exp((v - mx) / tau) as 2^((v - mx) log2(e) / tau)
                        const float2 mx2 = make_float2(mxA, mxB);
                        float2 e[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float2 arg = pairfft::pmul(pairfft::psub(v[u], mx2), itl);
                            e[u] = make_float2(live ? fast_ex2(arg.x) : 0.f, live ? fast_ex2(arg.y) : 0.f);
                        }
                        float2 sum = pairfft::padd(pairfft::padd(e[0], e[1]), pairfft::padd(e[2], e[3]));
                        if (M == 32) {
#pragma unroll
                            for (int off = 4; off > 0; off >>= 1)
                                sum = pairfft::padd(sum, make_float2(__shfl_xor_sync(0xffffffffu, sum.x, off),
                                                                     __shfl_xor_sync(0xffffffffu, sum.y, off)));
                        } else {
                            for (int off = LS >> 1; off > 0; off >>= 1)
                                sum = pairfft::padd(sum, make_float2(__shfl_xor_sync(0xffffffffu, sum.x, off),
                                                                     __shfl_xor_sync(0xffffffffu, sum.y, off)));
                        }
                        if (live) {
                            const float2 inv = make_float2(fast_rcp(sum.x), fast_rcp(sum.y));
                            float2 bta[4];
                            float2 tsum = make_float2(0.f, 0.f);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                bta[u] = pairfft::pmul(e[u], inv);
                                // 1 - sum beta^2 = sum beta (1 - beta), 1 - beta = (sum - e) / sum: no cancellation
                                tsum = pairfft::pfma(bta[u], pairfft::pmul(pairfft::psub(sum, e[u]), inv), tsum);
                            }
                            st4_cg(beta2_c + j0, bta);
                            psA += (double)tsum.x;
                            psB += (double)tsum.y;
                            const int sec = c * secs_per_col + (j0 >> logM);
                            if (((unsigned)lane & (unsigned)(LS - 1)) == 0u) {
                                if (actA) o.idx[fA * d.L + sec] = balA && wuA < 4 ? (srcA - (int)grp_base) * 4 + wuA : 0;
                                if (actB) o.idx[fB * d.L + sec] = balB && wuB < 4 ? (srcB - (int)grp_base) * 4 + wuB : 0;
                            }
                            if (o.nmse) {
                                const int k0 = j0 & (M - 1), tA = idx0[fA * d.L + sec], tB = idx0[fB * d.L + sec];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float dA = bta[u].x - (k0 + u == tA ? 1.f : 0.f);
                                    const float dB = bta[u].y - (k0 + u == tB ? 1.f : 0.f);
                                    nmA += (double)dA * (double)dA;
                                    nmB += (double)dB * (double)dB;
                                }
                            }
                        }
                    }
                    if (last) {
                        psA = warp_sum(psA);
                        psB = warp_sum(psB);
                        if (o.nmse) { nmA = warp_sum(nmA); nmB = warp_sum(nmB); }
                        if (lane == 0) {
                            part_psi[c & 1][warp][0] = psA;
                            part_psi[c & 1][warp][1] = psB;
                            part_nm[c & 1][warp][0] = nmA;
                            part_nm[c & 1][warp][1] = nmB;
                        }
                    }
                    __syncthreads();
                }
                // column totals in fixed warp order (the slots of parity c & 1 are rewritten two columns later,
                // several barriers after this read)
                if (tid < 2) {
                    double a = 0.0, bsum = 0.0;
                    for (int w = 0; w < kWarps; ++w) { a += part_psi[c & 1][w][tid]; bsum += part_nm[c & 1][w][tid]; }
                    S.psi()[tid * Lc + c] = a;
                    S.nm()[tid * Lc + c] = bsum;
                }
            }
            // ---- psi, nmse, early stop (:976-986), per frame
            if (tid < 2 && s_active[tid]) {
                const long long f = tid ? fB : fA;
                double *psi = S.psi() + tid * Lc, *psi_prev = S.psi_prev() + tid * Lc, *nm = S.nm() + tid * Lc;
                int stop = t > 0;
                for (int c = 0; c < Lc; ++c) {
                    const double ps = psi[c] / sec_norm;
                    psi[c] = ps;
                    if (!(fabs(ps - psi_prev[c]) <= p.atol + p.rtol * fabs(psi_prev[c]))) stop = 0;
                    if (o.psi_trace) o.psi_trace[(f * p.t_max + t) * Lc + c] = ps;
                    if (o.nmse) o.nmse[(f * p.t_max + t + 1) * Lc + c] = (float)(nm[c] / sec_norm);
                }
                if (stop) {
                    if (o.nmse)                                       // nmse[t:] = nmse[t]  (:985)
                        for (int tt = t + 1; tt < p.t_max; ++tt)
                            for (int c = 0; c < Lc; ++c)
                                o.nmse[(f * p.t_max + tt) * Lc + c] = o.nmse[(f * p.t_max + t) * Lc + c];
                    o.t_final[f] = t + 1;
                    for (int c = 0; c < Lc; ++c) o.psi[f * Lc + c] = psi[c];
                    s_active[tid] = 0;
                }
            }
            __syncthreads();
            if (!s_active[0] && !s_active[1]) break;
        }
        // frames that ran all t_max - 1 iterations
        if (tid < 2 && s_active[tid]) {
            const long long f = tid ? fB : fA;
            o.t_final[f] = p.t_max - 1;
            for (int c = 0; c < Lc; ++c) o.psi[f * Lc + c] = S.psi()[tid * Lc + c];
        }
        __syncthreads();
    }
}
#undef LSB_DV_BLOCKS
#undef LSB_DV_COLS
#undef LSB_DV_W

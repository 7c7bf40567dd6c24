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

// dst2[i] += sc * Y[tab[i]] for the rows of one block (Y: float2 view of the pair buffer)
__device__ __forceinline__ void pair_gather_add(float2 *dst2, const float4 *buf, const uint16_t *__restrict__ tab, float sc,
                                                int count, int tid)
{
    const float2 *bf2 = reinterpret_cast<const float2 *>(buf);
    for (int i0 = tid; i0 < count; i0 += kPairThreads * kBatch) {
        int id[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kPairThreads;
            id[u] = i < count ? (int)__ldg(tab + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kPairThreads;
            if (i < count) dst2[i] = pairfft::pfmas(bf2[id[u]], sc, dst2[i]);
        }
    }
}

size_t pair_amp_dyn_smem(int h, int n) { return ((size_t)16 << h) + (size_t)n * sizeof(float2); }

// ------------------------------------------------------------------ operators
// out[f, n] = A x[f, LM] for frame pairs (reference Ab, sparc.py:786-789,811-819,851-862)
template <int LOGH>
__global__ void __launch_bounds__(kPairThreads)
ab_pair_kernel(DesignDev d, PairDev pd, const float *__restrict__ x, float *__restrict__ out, long long n_frames)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int H = 1 << LOGH;
    float4 *buf = reinterpret_cast<float4 *>(smem);
    float2 *acc2 = reinterpret_cast<float2 *>(buf + H);       // [n]
    const int tid = threadIdx.x;
    const long long LM = (long long)d.L * d.M;
    for (long long fA = 2LL * blockIdx.x; fA < n_frames; fA += 2LL * gridDim.x) {
        const long long fB = fA + 1 < n_frames ? fA + 1 : fA;
        for (int i = tid; i < d.n; i += kPairThreads) acc2[i] = make_float2(0.f, 0.f);
        for (int bi = 0; bi < d.n_blocks; ++bi) {
            const PairBlock b = pd.blocks[bi];
            const size_t off = (size_t)b.c * d.Mc;
            pairfft::forward<LOGH>(buf, pd.tb, b, SrcTwo{x + fA * LM + off, x + fB * LM + off}, tid);
            pair_gather_add(acc2 + b.r * d.Mr, buf, b.fgat, b.scale, d.Mr, tid);
            __syncthreads();
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
    float2 *y2 = reinterpret_cast<float2 *>(buf + H);         // [n]
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
                pairfft::transposed<LOGH>(buf, pd.tb, b, y2 + b.r * d.Mr, tid);
                const float2 *bf2 = reinterpret_cast<const float2 *>(buf);
                for (int j = tid; j < d.Mc; j += kPairThreads) {
                    const float2 v = bf2[__ldg(b.tgat + j)];
                    const float va = b.scale * v.x, vb = b.scale * v.y;
                    if (bi == b0) {
                        oA[j] = va;
                        if (fB != fA) oB[j] = vb;
                    } else {
                        oA[j] += va;
                        if (fB != fA) oB[j] += vb;
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
// Requirements checked by the host: fp32, unmodulated (K = 1), M <= 32, every column block has at least
// one non-zero base-matrix entry, no soft outputs (beta_soft / s_last), H = 2^LOGH.
// scratch per CTA: [LM] float2 beta (A, B) then [Mc] float2 partial s of the current column.
template <int LOGH>
__global__ void __launch_bounds__(kPairThreads, LSB_AMP_PAIR_OCC)
amp_pair_kernel(DesignDev d, PairDev pd, AmpParams p, const float *__restrict__ y_all,
                const int32_t *__restrict__ idx0, AmpOut o, float2 *__restrict__ scratch, long long n_frames,
                unsigned long long *__restrict__ queue)
{
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int H = 1 << LOGH;
    constexpr int kWarps = kPairThreads / 32;
    float4 *buf = reinterpret_cast<float4 *>(smem);           // [H] (re_A, re_B, im_A, im_B)
    float2 *z2 = reinterpret_cast<float2 *>(buf + H);         // [n] residual of frame A / frame B
    __shared__ double psi[2][kMaxLc], psi_prev[2][kMaxLc], tau[2][kMaxLc], nm[2][kMaxLc];
    __shared__ double phi[2][kMaxLr], gam[2][kMaxLr], zsq[2][kMaxLr];
    __shared__ float2 tau_f[kMaxLc], itau_f[kMaxLc];          // tau_c, log2(e) / tau_c          (A, B)
    __shared__ float2 iphi_f[kMaxLr], coef_f[kMaxLr];         // 1 / phi_r, gamma_r / phi_prev_r (A, B)
    __shared__ double part_psi[2][kWarps][2], part_nm[2][kWarps][2];   // [column parity][warp][frame]
    __shared__ double part_z[kWarps][2];
    __shared__ long long s_frame;
    __shared__ int s_active[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = (tid >> 6) & 1, ix = tid & 63;             // scalar work: threads 0..127 = (frame, index)
    const int LM = d.L * d.M;
    float2 *beta2 = scratch + (size_t)blockIdx.x * ((size_t)LM + d.Mc);
    float2 *s2col = beta2 + LM;
    const int secs_per_col = d.Mc / d.M;
    const double sec_norm = (double)d.L / d.Lc;
    const int M = d.M, logM = d.logM;
    const unsigned grp_base = (unsigned)lane & ~(unsigned)(M - 1);
    const unsigned grp_mask = M == 32 ? 0xffffffffu : (((1u << M) - 1u) << grp_base);

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
            if (ix < d.Lc) psi[fr][ix] = 1.0;                                                 // gamma = W.1/Lc (:918,926)
            phi[fr][ix] = 0.0;
        }
        if (tid < 2) s_active[tid] = tid == 0 ? 1 : (hasB ? 1 : 0);
        __syncthreads();

        int t = 0;
        for (; t < p.t_max - 1; ++t) {
            // ---- gamma_r, the Onsager coefficient gamma_r / phi_prev_r (phi still holds the previous iteration's)
            if (tid < 128) {
                if (ix < d.Lr) {
                    double g = 0.0;
                    for (int c = 0; c < d.Lc; ++c) g += d.W[ix * d.Lc + c] * psi[fr][c];
                    g /= d.Lc;
                    gam[fr][ix] = g;
                    reinterpret_cast<float *>(&coef_f[ix])[fr] = t > 0 ? (float)(g / phi[fr][ix]) : 0.f;
                    zsq[fr][ix] = 0.0;
                }
                if (ix < d.Lc) psi_prev[fr][ix] = psi[fr][ix];
            }
            __syncthreads();
            if (t > 0) {
                // z = y - A beta + b z (:944-946): z <- y + b z, then every block subtracts its A_rc beta_c.
                // Row i is owned by thread i mod 256 here and in pair_gather_add: no barrier in between.
                if (d.Lr == 1) {
                    const float2 cf = coef_f[0];
                    for (int i = tid; i < d.n; i += kPairThreads) z2[i] = pairfft::pfma(z2[i], cf, make_float2(yA[i], yB[i]));
                } else {
                    for (int i = tid; i < d.n; i += kPairThreads)
                        z2[i] = pairfft::pfma(z2[i], coef_f[i / d.Mr], make_float2(yA[i], yB[i]));
                }
                for (int bi = 0; bi < d.n_blocks; ++bi) {
                    const PairBlock b = pd.blocks[bi];
                    pairfft::forward<LOGH>(buf, pd.tb, b, SrcPair{beta2 + (size_t)b.c * d.Mc}, tid);
                    pair_gather_add(z2 + b.r * d.Mr, buf, b.fgat, -b.scale, d.Mr, tid);
                    __syncthreads();
                }
            }
            // ---- phi (:949-955)
            if (p.phi_method == 2) {
                for (int r = 0; r < d.Lr; ++r) {
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
                        zsq[tid][r] = a;
                    }
                    __syncthreads();
                }
            }
            if (tid < 128 && ix < d.Lr) {
                const double ph = p.phi_method == 1 ? p.awgn_var + gam[fr][ix] : zsq[fr][ix] / d.Mr;
                phi[fr][ix] = ph;
                reinterpret_cast<float *>(&iphi_f[ix])[fr] = (float)(1.0 / ph);
            }
            __syncthreads();
            // ---- tau (:958-969)
            if (tid < 128 && ix < d.Lc) {
                double tc;
                if (d.Lr == 1) {
                    tc = ((double)d.L * phi[fr][0] / d.n) / d.W[ix];
                } else {
                    double acc_w = 0.0;
                    for (int r = 0; r < d.Lr; ++r) acc_w += d.W[r * d.Lc + ix] / phi[fr][r];
                    tc = ((double)d.L / d.Mr) / acc_w;
                }
                tau[fr][ix] = tc;
                reinterpret_cast<float *>(&tau_f[ix])[fr] = (float)tc;
                reinterpret_cast<float *>(&itau_f[ix])[fr] = (float)(1.4426950408889634 / tc);
                if (o.tau_trace && s_active[fr]) o.tau_trace[(fsel * p.t_max + t) * d.Lc + ix] = tc;
            }
            __syncthreads();
            const bool actA = s_active[0] != 0, actB = s_active[1] != 0;
            // ---- s = beta + tau * Az(z / phi) column block by column block (:972), softmax denoiser (:429-432),
            //      MAP decision argmax(s) of this iteration (:997)
            for (int c = 0; c < d.Lc; ++c) {
                const float2 tau2 = tau_f[c], itl = itau_f[c];
                float2 *beta2_c = beta2 + (size_t)c * d.Mc;
                const int b0 = d.col_first[c], b1 = d.col_first[c + 1];
                double psA = 0.0, psB = 0.0, nmA = 0.0, nmB = 0.0;
                for (int bi = b0; bi < b1; ++bi) {
                    const PairBlock b = pd.blocks[bi];
                    pairfft::transposed<LOGH>(buf, pd.tb, b, z2 + b.r * d.Mr, tid);
                    const float2 *bf2 = reinterpret_cast<const float2 *>(buf);
                    const float2 sc2 = pairfft::pmul(pairfft::pscale(tau2, b.scale), iphi_f[b.r]);   // scale tau_c / phi_r
                    const float2 *prev = bi == b0 ? (t > 0 ? beta2_c : nullptr) : s2col;
                    if (bi < b1 - 1) {
                        for (int j0 = tid; j0 < d.Mc; j0 += kPairThreads * kBatch) {
                            int id[kBatch];
                            float2 pv[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = j0 + u * kPairThreads;
                                id[u] = j < d.Mc ? (int)__ldg(b.tgat + j) : 0;
                                pv[u] = (prev && j < d.Mc) ? prev[j] : make_float2(0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = j0 + u * kPairThreads;
                                if (j < d.Mc) s2col[j] = pairfft::pfma(bf2[id[u]], sc2, pv[u]);
                            }
                        }
                    } else {
                        // last block of the column: 32 consecutive j of a warp hold whole sections, the softmax
                        // runs on the gathered values in registers
                        for (int base = 0; base < d.Mc; base += kPairThreads * kBatch) {     // warp-uniform trip count
                            int id[kBatch];
                            float2 pv[kBatch];
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = base + tid + u * kPairThreads;
                                id[u] = j < d.Mc ? (int)__ldg(b.tgat + j) : 0;
                                pv[u] = (prev && j < d.Mc) ? prev[j] : make_float2(0.f, 0.f);
                            }
#pragma unroll
                            for (int u = 0; u < kBatch; ++u) {
                                const int j = base + tid + u * kPairThreads;
                                const bool live = j < d.Mc;                       // uniform over a section's lanes
                                const float2 v = pairfft::pfma(bf2[id[u]], sc2, pv[u]);
                                float mxA = live ? v.x : -INFINITY, mxB = live ? v.y : -INFINITY;
                                if (M == 32) {
#pragma unroll
                                    for (int off = 16; off > 0; off >>= 1) {
                                        mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, off));
                                        mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, off));
                                    }
                                } else {
                                    for (int off = M >> 1; off > 0; off >>= 1) {
                                        mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, off));
                                        mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, off));
                                    }
                                }
                                // decision of this iteration: lowest index attaining the section maximum
                                const unsigned balA = __ballot_sync(0xffffffffu, live && v.x == mxA) & grp_mask;
                                const unsigned balB = __ballot_sync(0xffffffffu, live && v.y == mxB) & grp_mask;
                                // e = exp((v - mx) / tau) as 2^((v - mx) log2(e) / tau)
                                const float2 arg = pairfft::pmul(pairfft::psub(v, make_float2(mxA, mxB)), itl);
                                float2 e = make_float2(live ? fast_ex2(arg.x) : 0.f, live ? fast_ex2(arg.y) : 0.f);
                                float2 sum = e;
                                if (M == 32) {
#pragma unroll
                                    for (int off = 16; off > 0; off >>= 1)
                                        sum = pairfft::padd(sum, make_float2(__shfl_xor_sync(0xffffffffu, sum.x, off),
                                                                             __shfl_xor_sync(0xffffffffu, sum.y, off)));
                                } else {
                                    for (int off = M >> 1; off > 0; off >>= 1)
                                        sum = pairfft::padd(sum, make_float2(__shfl_xor_sync(0xffffffffu, sum.x, off),
                                                                             __shfl_xor_sync(0xffffffffu, sum.y, off)));
                                }
                                if (live) {
                                    const float2 inv = make_float2(fast_rcp(sum.x), fast_rcp(sum.y));
                                    const float2 bta = pairfft::pmul(e, inv);
                                    beta2_c[j] = bta;
                                    // 1 - sum beta^2 = sum beta (1 - beta), 1 - beta = (sum - e) / sum: no cancellation
                                    const float2 term = pairfft::pmul(bta, pairfft::pmul(pairfft::psub(sum, e), inv));
                                    psA += (double)term.x;
                                    psB += (double)term.y;
                                    const int sec = c * secs_per_col + (j >> logM);
                                    if (((unsigned)lane & (unsigned)(M - 1)) == 0u) {
                                        if (actA) o.idx[fA * d.L + sec] = balA ? (int)(__ffs(balA) - 1 - (int)grp_base) : 0;
                                        if (actB) o.idx[fB * d.L + sec] = balB ? (int)(__ffs(balB) - 1 - (int)grp_base) : 0;
                                    }
                                    if (o.nmse) {
                                        const int k = j & (M - 1);
                                        const float dA = bta.x - (k == idx0[fA * d.L + sec] ? 1.f : 0.f);
                                        const float dB = bta.y - (k == idx0[fB * d.L + sec] ? 1.f : 0.f);
                                        nmA += (double)dA * (double)dA;
                                        nmB += (double)dB * (double)dB;
                                    }
                                }
                            }
                        }
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
                    psi[tid][c] = a;
                    nm[tid][c] = bsum;
                }
            }
            // ---- psi, nmse, early stop (:976-986), per frame
            if (tid < 2 && s_active[tid]) {
                const long long f = tid ? fB : fA;
                int stop = t > 0;
                for (int c = 0; c < d.Lc; ++c) {
                    const double ps = psi[tid][c] / sec_norm;
                    psi[tid][c] = ps;
                    if (!(fabs(ps - psi_prev[tid][c]) <= p.atol + p.rtol * fabs(psi_prev[tid][c]))) stop = 0;
                    if (o.psi_trace) o.psi_trace[(f * p.t_max + t) * d.Lc + c] = ps;
                    if (o.nmse) o.nmse[(f * p.t_max + t + 1) * d.Lc + c] = (float)(nm[tid][c] / sec_norm);
                }
                if (stop) {
                    if (o.nmse)                                       // nmse[t:] = nmse[t]  (:985)
                        for (int tt = t + 1; tt < p.t_max; ++tt)
                            for (int c = 0; c < d.Lc; ++c)
                                o.nmse[(f * p.t_max + tt) * d.Lc + c] = o.nmse[(f * p.t_max + t) * d.Lc + c];
                    o.t_final[f] = t + 1;
                    for (int c = 0; c < d.Lc; ++c) o.psi[f * d.Lc + c] = psi[tid][c];
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
            for (int c = 0; c < d.Lc; ++c) o.psi[f * d.Lc + c] = psi[tid][c];
        }
        __syncthreads();
    }
}

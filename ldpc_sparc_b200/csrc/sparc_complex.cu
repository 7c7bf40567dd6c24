// Complex SPARCs: sub-sampled DFT design operators and the AMP decoder with K-PSK
// modulated sections (sm_100a).
//
// Reference (sparc_public/sparc.py): sub_fft :593-646 and sparc_transforms :703-880 with
// csparc=True -- block (r, c) of the design is
//     A_rc[i, j] = sqrt(W[r,c]/L) * exp(-2 pi i o0[i] o1[j] / w),
// o0 / o1 drawn from {1..w-1} \ {w/2};  Ab = A x by zero-padded scatter -> FFT_w -> gather,
// Az = A^H y by scatter -> conj(FFT(conj .)) -> gather.  sparc_amp :883-999 runs unchanged
// on complex z / s; the denoiser (msg_vector_mmse_estimator :402-465) halves tau for complex
// observations and, with c_k the K-PSK constellation (c = {1} when unmodulated),
//     beta_j = sum_k c_k e^{t_jk} / sum_{j' in section} sum_k e^{t_j'k},
//     t_jk = Re(s_j conj(c_k)) / (tau / 2),
// which is what the reference's K = 1, 2, 4 and K > 4 branches all evaluate; the MAP
// decision (:467-512) is the (j, k) with the largest Re(s_j conj(c_k)).
//
// Same kernel plan as the real design (sparc_amp.cu): one CTA per frame from a device
// queue, the FFT buffer (w complex slots, bank-swizzled, bit-reversed output) and the
// residual z in shared memory, beta / s in a per-CTA global scratch, permutations folded
// into four uint16 slot tables per block.  This path is sized for correctness and a
// reasonable rate, not tuned like the real-design kernel.
#include "common.cuh"
#include "sparc_dct.cuh"

#include <atomic>
#include <cmath>
#include <mutex>
#include <new>
#include <vector>

using namespace lsb_sparc;

namespace {

constexpr int kThreads = 256;
constexpr int kMaxLr = 64, kMaxLc = 64;
constexpr int kMaxK = 64;
constexpr int kCounters = 64;

struct CBlock {
    int r, c;
    float scale;         // sqrt(W[r,c] / L)
    double scale_d;
    const uint16_t *ab_scatter, *ab_gather, *az_scatter, *az_gather;   // complex slots of the FFT buffer
};

struct CDesign {
    int L, M, n, Lr, Lc, Mr, Mc, w, h, n_blocks;
    const CBlock *blocks;     // sorted by (c, r)
    const int *col_first;     // [Lc + 1]
    const double *W;          // [Lr * Lc]
};

struct CParams {
    double awgn_var, rtol, atol;
    int t_max, phi_method, K;
};

struct COut {
    int32_t *idx, *t_final;
    double *psi;
    float *nmse;
    void *beta_soft, *s_last;
    double *tau_trace, *psi_trace;
};

template <typename T> struct BScale;
template <> struct BScale<float> { static __device__ __forceinline__ float get(const CBlock &b) { return b.scale; } };
template <> struct BScale<double> { static __device__ __forceinline__ double get(const CBlock &b) { return b.scale_d; } };

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

// buf <- FFT_w (forward) or its unnormalised inverse of the zero-padded scatter of src
template <typename T, bool INV>
__device__ void transform(typename Cplx<T>::type *buf, const FftTables<T> &tb, const uint16_t *__restrict__ tab,
                          const typename Cplx<T>::type *src, T src_scale, int count, int w, int tid)
{
    typedef typename Cplx<T>::type T2;
    for (int i = tid; i < w; i += kThreads) buf[i] = T2{(T)0, (T)0};
    __syncthreads();
    for (int i = tid; i < count; i += kThreads) {
        const T2 v = src[i];
        buf[tab[i]] = T2{v.x * src_scale, v.y * src_scale};
    }
    __syncthreads();
    if constexpr (sizeof(T) == 4) {          // compile-time-size passes for the common fp32 sizes (sparc_dct.cuh)
        switch (tb.h) {
            case 11: fft_fixed<T, 11, INV, kThreads>(buf, tb, tid); return;
            case 12: fft_fixed<T, 12, INV, kThreads>(buf, tb, tid); return;
            case 13: fft_fixed<T, 13, INV, kThreads>(buf, tb, tid); return;
            default: break;
        }
    }
    fft_inplace<T, INV>(buf, tb, tid, kThreads);
}

// out[f] = A x[f]  (TRANS = false)  or  A^H y[f]  (TRANS = true)
template <typename T, bool TRANS>
__global__ void __launch_bounds__(kThreads)
cop_kernel(CDesign d, FftTables<T> tb, const typename Cplx<T>::type *__restrict__ in,
           typename Cplx<T>::type *__restrict__ out, long long n_frames, typename Cplx<T>::type *gbuf)
{
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename Cplx<T>::type T2;
    // gbuf != null: the FFT buffer does not fit shared memory and lives in a per-CTA global scratch (L2)
    T2 *buf = gbuf ? gbuf + (size_t)blockIdx.x * d.w : reinterpret_cast<T2 *>(smem);
    T2 *acc = gbuf ? reinterpret_cast<T2 *>(smem) : buf + d.w;     // [n], TRANS = false only
    const int tid = threadIdx.x;
    const long long LM = (long long)d.L * d.M;
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        if (!TRANS) {
            for (int i = tid; i < d.n; i += kThreads) acc[i] = T2{(T)0, (T)0};
            for (int bi = 0; bi < d.n_blocks; ++bi) {
                const CBlock b = d.blocks[bi];
                transform<T, false>(buf, tb, b.ab_scatter, in + f * LM + (long long)b.c * d.Mc, (T)1, d.Mc, d.w, tid);
                const T sc = BScale<T>::get(b);
                T2 *dst = acc + b.r * d.Mr;
                for (int i = tid; i < d.Mr; i += kThreads) {
                    const T2 v = buf[b.ab_gather[i]];
                    dst[i].x += sc * v.x;
                    dst[i].y += sc * v.y;
                }
                __syncthreads();
            }
            for (int i = tid; i < d.n; i += kThreads) out[f * d.n + i] = acc[i];
            __syncthreads();
        } else {
            T2 *o = out + f * LM;
            for (int c = 0; c < d.Lc; ++c) {
                for (int bi = d.col_first[c]; bi < d.col_first[c + 1]; ++bi) {
                    const CBlock b = d.blocks[bi];
                    transform<T, true>(buf, tb, b.az_scatter, in + f * d.n + b.r * d.Mr, (T)1, d.Mr, d.w, tid);
                    const T sc = BScale<T>::get(b);
                    const bool first = bi == d.col_first[c];
                    for (int j = tid; j < d.Mc; j += kThreads) {
                        const T2 v = buf[b.az_gather[j]];
                        T2 *dst = o + (long long)c * d.Mc + j;
                        const T2 prev = first ? T2{(T)0, (T)0} : *dst;
                        *dst = T2{prev.x + sc * v.x, prev.y + sc * v.y};
                    }
                    __syncthreads();
                }
                if (d.col_first[c] == d.col_first[c + 1])
                    for (int j = tid; j < d.Mc; j += kThreads) o[(long long)c * d.Mc + j] = T2{(T)0, (T)0};
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 4 : 2)
camp_kernel(CDesign d, FftTables<T> tb, CParams p, const typename Cplx<T>::type *__restrict__ y_all,
            const int32_t *__restrict__ idx0, COut o, typename Cplx<T>::type *__restrict__ scratch,
            long long n_frames, unsigned long long *__restrict__ queue, int gbuf)
{
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename Cplx<T>::type T2;
    const int LM = d.L * d.M;
    // per-CTA global scratch: [beta LM][s LM] (+ [FFT buffer w] when it does not fit shared memory)
    T2 *beta = scratch + (size_t)blockIdx.x * (2 * (size_t)LM + (gbuf ? (size_t)d.w : 0));
    T2 *s = beta + LM;
    T2 *buf = gbuf ? s + LM : reinterpret_cast<T2 *>(smem);
    T2 *z = gbuf ? reinterpret_cast<T2 *>(smem) : buf + d.w;       // [n]
    __shared__ double psi[kMaxLc], psi_prev[kMaxLc], tau[kMaxLc], nm[kMaxLc];
    __shared__ double phi[kMaxLr], phi_prev[kMaxLr], gam[kMaxLr], zsq[kMaxLr];
    // per-warp partial sums added in fixed warp order by one thread (bit-reproducible run to run)
    __shared__ double part_psi[2][kThreads / 32], part_nm[2][kThreads / 32], part_z[kThreads / 32];
    __shared__ T2 con[kMaxK];                  // constellation
    __shared__ long long s_frame;
    __shared__ int s_stop;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int secs_per_col = d.Mc / d.M;
    const double sec_norm = (double)d.L / d.Lc;
    const int K = p.K;
    if (tid < K) {
        // psk_constel (:225-239): exact for K <= 4, cos / sin of 2 pi k / K otherwise
        double cs, sn;
        sincospi(2.0 * tid / K, &sn, &cs);
        if (K <= 4) { cs = rint(cs); sn = rint(sn); }
        con[tid] = T2{(T)cs, (T)sn};
    }
    __syncthreads();

    for (;;) {
        if (tid == 0) s_frame = (long long)atomicAdd(queue, 1ull);
        __syncthreads();
        const long long f = s_frame;
        if (f >= n_frames) break;
        const T2 *y = y_all + f * d.n;
        for (int i = tid; i < d.n; i += kThreads) z[i] = y[i];
        for (int i = tid; i < LM; i += kThreads) beta[i] = T2{(T)0, (T)0};
        if (tid < d.Lc) psi[tid] = 1.0;
        if (tid < kMaxLr) phi[tid] = 0.0;
        __syncthreads();

        int t = 0;
        for (; t < p.t_max - 1; ++t) {
            if (tid < d.Lr) {
                double g = 0.0;
                for (int c = 0; c < d.Lc; ++c) g += d.W[tid * d.Lc + c] * psi[c];
                gam[tid] = g / d.Lc;
                phi_prev[tid] = phi[tid];
                zsq[tid] = 0.0;
            }
            if (tid < d.Lc) psi_prev[tid] = psi[tid];
            __syncthreads();
            if (t > 0) {          // z = y - A beta + b z (:942-946), accumulated in place
                for (int i = tid; i < d.n; i += kThreads) {
                    const T coef = (T)(gam[i / d.Mr] / phi_prev[i / d.Mr]);
                    const T2 yv = y[i], zv = z[i];
                    z[i] = T2{yv.x + coef * zv.x, yv.y + coef * zv.y};
                }
                for (int bi = 0; bi < d.n_blocks; ++bi) {
                    const CBlock b = d.blocks[bi];
                    transform<T, false>(buf, tb, b.ab_scatter, beta + (size_t)b.c * d.Mc, (T)1, d.Mc, d.w, tid);
                    const T sc = BScale<T>::get(b);
                    T2 *dst = z + b.r * d.Mr;
                    for (int i = tid; i < d.Mr; i += kThreads) {
                        const T2 v = buf[b.ab_gather[i]];
                        dst[i].x -= sc * v.x;
                        dst[i].y -= sc * v.y;
                    }
                    __syncthreads();
                }
            }
            if (p.phi_method == 2) {            // phi = mean |z|^2 per row block (:949-955)
                for (int r = 0; r < d.Lr; ++r) {
                    double part = 0.0;
                    for (int i = tid; i < d.Mr; i += kThreads) {
                        const T2 v = z[r * d.Mr + i];
                        part += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
                    }
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
            if (tid < d.Lc) {                   // tau (:958-969)
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
            for (int c = 0; c < d.Lc; ++c) {
                const T tau_c = (T)tau[c];
                T2 *s_c = s + (size_t)c * d.Mc;
                T2 *beta_c = beta + (size_t)c * d.Mc;
                const int b0 = d.col_first[c], b1 = d.col_first[c + 1];
                for (int bi = b0; bi < b1; ++bi) {     // s = beta + tau A^H (z / phi)  (:972)
                    const CBlock b = d.blocks[bi];
                    transform<T, true>(buf, tb, b.az_scatter, z + b.r * d.Mr, (T)(1.0 / phi[b.r]), d.Mr, d.w, tid);
                    const T sc = BScale<T>::get(b) * tau_c;
                    const T2 *prev = bi == b0 ? beta_c : s_c;
                    for (int j = tid; j < d.Mc; j += kThreads) {
                        const T2 v = buf[b.az_gather[j]], pv = prev[j];
                        s_c[j] = T2{pv.x + sc * v.x, pv.y + sc * v.y};
                    }
                    __syncthreads();
                }
                if (b0 == b1)
                    for (int j = tid; j < d.Mc; j += kThreads) s_c[j] = beta_c[j];
                // denoiser: one warp per section; exponent t_jk = Re(s_j conj(c_k)) / (tau / 2)
                const T inv_tau = (T)2 / tau_c;
                double psi_part = 0.0, nm_part = 0.0;
                for (int sec = warp; sec < secs_per_col; sec += kThreads / 32) {
                    const T2 *ss = s_c + (size_t)sec * d.M;
                    T mx = -INFINITY;
                    for (int j = lane; j < d.M; j += 32) {
                        const T2 v = ss[j];
                        for (int k = 0; k < K; ++k) mx = fmax(mx, (v.x * con[k].x + v.y * con[k].y) * inv_tau);
                    }
                    mx = warp_max(mx);
                    T bot = (T)0;
                    for (int j = lane; j < d.M; j += 32) {
                        const T2 v = ss[j];
                        for (int k = 0; k < K; ++k) bot += exp((v.x * con[k].x + v.y * con[k].y) * inv_tau - mx);
                    }
                    bot = warp_sum(bot);
                    const T inv = (T)1 / bot;
                    const int truth = idx0 ? idx0[f * d.L + c * secs_per_col + sec] : -1;
                    const int t_loc = truth >= 0 ? (truth & (d.M - 1)) : -1;
                    const T2 t_val = truth >= 0 ? con[truth / d.M] : T2{(T)0, (T)0};
                    for (int j = lane; j < d.M; j += 32) {
                        const T2 v = ss[j];
                        T tx = (T)0, ty = (T)0, row = (T)0;
                        for (int k = 0; k < K; ++k) {
                            const T e = exp((v.x * con[k].x + v.y * con[k].y) * inv_tau - mx);
                            tx += e * con[k].x;
                            ty += e * con[k].y;
                            row += e;
                        }
                        const T2 bta = T2{tx * inv, ty * inv};
                        beta_c[(size_t)sec * d.M + j] = bta;
                        // 1 - sum |beta|^2 = sum p (1 - p) + (p^2 - |beta|^2),  p_j = row_j / bot
                        const double pj = (double)(row * inv), om = (double)((bot - row) * inv);
                        const double mag2 = (double)bta.x * (double)bta.x + (double)bta.y * (double)bta.y;
                        psi_part += pj * om + (pj * pj - mag2);
                        const double dx = (double)bta.x - (j == t_loc ? (double)t_val.x : 0.0);
                        const double dy = (double)bta.y - (j == t_loc ? (double)t_val.y : 0.0);
                        nm_part += dx * dx + dy * dy;
                    }
                }
                psi_part = warp_sum(psi_part);
                nm_part = warp_sum(nm_part);
                if (lane == 0) {
                    part_psi[c & 1][warp] = psi_part;
                    part_nm[c & 1][warp] = nm_part;
                }
                __syncthreads();
                if (tid == 0) {         // slots of parity c & 1 are rewritten two columns (several barriers) later
                    double a = 0.0, bsum = 0.0;
                    for (int w = 0; w < kThreads / 32; ++w) { a += part_psi[c & 1][w]; bsum += part_nm[c & 1][w]; }
                    psi[c] = a;
                    nm[c] = bsum;
                }
            }
            if (tid == 0) {                     // psi, nmse, early stop (:976-986)
                int stop = t > 0;
                for (int c = 0; c < d.Lc; ++c) {
                    psi[c] /= sec_norm;
                    if (!(fabs(psi[c] - psi_prev[c]) <= p.atol + p.rtol * fabs(psi_prev[c]))) stop = 0;
                    if (o.psi_trace) o.psi_trace[(f * p.t_max + t) * d.Lc + c] = psi[c];
                    if (o.nmse) o.nmse[(f * p.t_max + t + 1) * d.Lc + c] = (float)(nm[c] / sec_norm);
                }
                s_stop = stop;
                if (stop && o.nmse)
                    for (int tt = t + 1; tt < p.t_max; ++tt)
                        for (int c = 0; c < d.Lc; ++c)
                            o.nmse[(f * p.t_max + tt) * d.Lc + c] = o.nmse[(f * p.t_max + t) * d.Lc + c];
            }
            __syncthreads();
            if (s_stop) break;
        }
        const int t_final = (t < p.t_max - 1) ? t + 1 : p.t_max - 1;
        if (tid == 0) o.t_final[f] = t_final;
        if (tid < d.Lc) o.psi[f * d.Lc + tid] = psi[tid];
        // MAP decision (:467-512): the (location, symbol) with the largest Re(s_j conj(c_k))
        for (int sec = warp; sec < d.L; sec += kThreads / 32) {
            const T2 *ss = s + (size_t)sec * d.M;
            T best = -INFINITY;
            int arg = 0x7fffffff;
            for (int j = lane; j < d.M; j += 32) {
                const T2 v = ss[j];
                for (int k = 0; k < K; ++k) {
                    const T corr = v.x * con[k].x + v.y * con[k].y;
                    if (corr > best) { best = corr; arg = j * K + k; }     // row-major (j, k) order breaks ties
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (lane == 0) o.idx[f * d.L + sec] = (arg / K) + d.M * (arg % K);
        }
        if (o.beta_soft) {
            T2 *dst = reinterpret_cast<T2 *>(o.beta_soft) + f * LM;
            for (int i = tid; i < LM; i += kThreads) dst[i] = beta[i];
        }
        if (o.s_last) {
            T2 *dst = reinterpret_cast<T2 *>(o.s_last) + f * LM;
            for (int i = tid; i < LM; i += kThreads) dst[i] = s[i];
        }
        __syncthreads();
    }
}

}  // namespace

// ------------------------------------------------------------------ host side
struct lsb_csparc_design {
    int device = 0;
    CDesign dev{};
    std::vector<void *> allocs;
    FftTables<float> tb32{};
    FftTables<double> tb64{};
    unsigned long long *d_queue = nullptr;
    std::atomic<unsigned> next_queue{0};
    std::mutex mu;
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t scratch_event = nullptr;
};

namespace {

template <typename T>
int upload(lsb_csparc_design *g, const T **dst, const std::vector<T> &src)
{
    void *p = nullptr;
    LSB_CUDA(cudaMalloc(&p, src.size() * sizeof(T)));
    g->allocs.push_back(p);
    LSB_CUDA(cudaMemcpy(p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T *)p;
    return LSB_OK;
}

template <typename T, typename T2>
int make_twiddles(lsb_csparc_design *g, FftTables<T> *tb, int h)
{
    const int H = 1 << h;
    const double PI = 3.14159265358979323846;
    std::vector<T2> tw(H);
    for (int k = 0; k < H; ++k) tw[k] = T2{(T)cos(2 * PI * k / H), (T)-sin(2 * PI * k / H)};
    tb->pt = nullptr;
    tb->ptr = nullptr;
    tb->h = h;
    return upload(g, &tb->tw, tw);
}

int brev_bits(int k, int h)
{
    int r = 0;
    for (int i = 0; i < h; ++i) r |= ((k >> i) & 1) << (h - 1 - i);
    return r;
}

template <typename T> const FftTables<T> &tables(const lsb_csparc_design *g);
template <> const FftTables<float> &tables<float>(const lsb_csparc_design *g) { return g->tb32; }
template <> const FftTables<double> &tables<double>(const lsb_csparc_design *g) { return g->tb64; }

size_t op_smem(const CDesign &d, size_t elem2, bool gbuf) { return ((gbuf ? 0 : (size_t)d.w) + (size_t)d.n) * elem2; }
bool needs_gbuf(const CDesign &d, size_t elem2) { return op_smem(d, elem2, false) + 8192 > (size_t)lsb::max_smem_optin(); }

int ensure_scratch(lsb_csparc_design *g, size_t need, cudaStream_t st)
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

int check_device(const lsb_csparc_design *g)
{
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "design was created on device %d, current device is %d", g->device, dev);
    return LSB_OK;
}

template <typename T, bool TRANS>
int run_cop(lsb_csparc_design *g, const void *in, void *out, int64_t n_frames, cudaStream_t st)
{
    typedef typename Cplx<T>::type T2;
    if (n_frames == 0) return LSB_OK;              // an empty batch is valid with any pointers
    LSB_REQUIRE(g && in && out, "null pointer");
    LSB_REQUIRE(n_frames >= 0, "negative frame count");
    if (n_frames == 0) return LSB_OK;
    int rc = check_device(g);
    if (rc) return rc;
    const bool gbuf = needs_gbuf(g->dev, sizeof(T2));
    const size_t smem = op_smem(g->dev, sizeof(T2), gbuf);
    LSB_REQUIRE(smem + 8192 <= (size_t)lsb::max_smem_optin(),
                "complex design needs %zu B of shared memory per CTA (n = %d): unsupported", smem, g->dev.n);
    LSB_CUDA(lsb::ensure_dynamic_smem(cop_kernel<T, TRANS>, smem));
    long long grid = 2LL * lsb::sm_count();
    if (grid > n_frames) grid = n_frames;
    std::lock_guard<std::mutex> lock(g->mu);
    if (gbuf && (rc = ensure_scratch(g, (size_t)grid * g->dev.w * sizeof(T2), st))) return rc;
    cop_kernel<T, TRANS><<<(unsigned)grid, kThreads, smem, st>>>(g->dev, tables<T>(g), (const T2 *)in, (T2 *)out,
                                                                 (long long)n_frames, gbuf ? (T2 *)g->d_scratch : nullptr);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    if (gbuf) LSB_CUDA(cudaEventRecord(g->scratch_event, st));
    return LSB_OK;
}

template <typename T>
int run_camp(lsb_csparc_design *g, const void *y, int64_t n_frames, const lsb_sparc_amp_params *pp, const int32_t *idx0,
             const lsb_sparc_amp_out *po, cudaStream_t st)
{
    typedef typename Cplx<T>::type T2;
    if (n_frames == 0) return LSB_OK;              // an empty batch is valid with any pointers
    LSB_REQUIRE(g && y && pp && po && po->idx && po->t_final && po->psi, "null pointer");
    LSB_REQUIRE(pp->t_max > 1, "t_max must be > 1");
    LSB_REQUIRE(pp->phi_est_method == 1 || pp->phi_est_method == 2, "phi_est_method must be 1 or 2");
    const int K = pp->modulation <= 1 ? 1 : pp->modulation;
    LSB_REQUIRE(K <= kMaxK && (K & (K - 1)) == 0, "modulation K must be a power of two <= %d", kMaxK);
    LSB_REQUIRE(n_frames >= 0, "negative frame count");
    LSB_REQUIRE(!po->nmse || idx0, "nmse output needs the true section symbols idx0");
    if (n_frames == 0) return LSB_OK;
    int rc = check_device(g);
    if (rc) return rc;
    const bool gbuf = needs_gbuf(g->dev, sizeof(T2));
    const size_t smem = op_smem(g->dev, sizeof(T2), gbuf);
    LSB_REQUIRE(smem + 8192 <= (size_t)lsb::max_smem_optin(),
                "complex design needs %zu B of shared memory per CTA (n = %d): unsupported", smem, g->dev.n);
    LSB_CUDA(lsb::ensure_dynamic_smem(camp_kernel<T>, smem));
    int per_sm = 0;
    LSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, camp_kernel<T>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    const long long resident = (long long)per_sm * lsb::sm_count();
    const long long grid = resident < n_frames ? resident : n_frames;
    const size_t need = (size_t)grid * (2 * (size_t)g->dev.L * g->dev.M + (gbuf ? (size_t)g->dev.w : 0)) * sizeof(T2);
    std::lock_guard<std::mutex> lock(g->mu);
    if ((rc = ensure_scratch(g, need, st))) return rc;
    unsigned long long *queue = g->d_queue + (g->next_queue.fetch_add(1) % kCounters);
    LSB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
    CParams p{pp->awgn_var, pp->rtol, pp->atol, pp->t_max, pp->phi_est_method, K};
    COut o{po->idx, po->t_final, po->psi, po->nmse, po->beta_soft, po->s_last, po->tau_trace, po->psi_trace};
    camp_kernel<T><<<(unsigned)grid, kThreads, smem, st>>>(g->dev, tables<T>(g), p, (const T2 *)y, idx0, o,
                                                           (T2 *)g->d_scratch, (long long)n_frames, queue, gbuf ? 1 : 0);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    LSB_CUDA(cudaEventRecord(g->scratch_event, st));
    return LSB_OK;
}

}  // namespace

extern "C" {

void lsb_csparc_design_destroy(lsb_csparc_design *g)
{
    if (!g) return;
    cudaDeviceSynchronize();
    for (void *p : g->allocs) cudaFree(p);
    cudaFree(g->d_queue);
    cudaFree(g->d_scratch);
    if (g->scratch_event) cudaEventDestroy(g->scratch_event);
    delete g;
}

int lsb_csparc_design_create(const double *W, int Lr, int Lc, int L, int M, int n, const uint32_t *order0,
                             const uint32_t *order1, lsb_csparc_design **out)
{
    LSB_REQUIRE(W && order0 && order1 && out, "null pointer");
    LSB_REQUIRE(Lr >= 1 && Lr <= kMaxLr && Lc >= 1 && Lc <= kMaxLc, "base matrix %dx%d outside [1,%d]x[1,%d]", Lr,
                Lc, kMaxLr, kMaxLc);
    LSB_REQUIRE(L > 0 && M > 0 && (M & (M - 1)) == 0 && n > 0, "bad L / M / n");
    LSB_REQUIRE(((long long)L * M) % Lc == 0 && n % Lr == 0 && L % Lc == 0, "Lc must divide L and LM, Lr must divide n");
    const int Mr = n / Lr, Mc = (int)((long long)L * M / Lc);
    LSB_REQUIRE(Mc % M == 0, "column blocks must hold whole sections");
    int w = 1, h = 0;
    while (w < (Mr > Mc ? Mr : Mc) + 2) { w <<= 1; ++h; }      // sparc.py:747
    LSB_REQUIRE(w >= 8 && w <= 65536, "transform size %d outside [8, 65536] (slot tables are 16-bit)", w);

    lsb_csparc_design *g = new (std::nothrow) lsb_csparc_design();
    if (!g) return LSB_ERR_NOMEM;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    int rc = LSB_OK;
    std::vector<CBlock> blocks;
    std::vector<int> col_first(Lc + 1, 0);
    std::vector<uint16_t> t_abs(Mc), t_abg(Mr), t_azs(Mr), t_azg(Mc);
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
                if (p < 0 || p >= w) { lsb::set_error("order1 entry %d outside [0, %d)", p, w); rc = LSB_ERR_INVALID; break; }
                t_abs[j] = (uint16_t)swz(p);
                t_azg[j] = (uint16_t)swz(brev_bits(p, h));
            }
            for (int i = 0; i < Mr && !rc; ++i) {
                const int q = (int)o0[i];
                if (q < 0 || q >= w) { lsb::set_error("order0 entry %d outside [0, %d)", q, w); rc = LSB_ERR_INVALID; break; }
                t_abg[i] = (uint16_t)swz(brev_bits(q, h));
                t_azs[i] = (uint16_t)swz(q);
            }
            if (rc) break;
            CBlock b;
            b.r = r;
            b.c = c;
            b.scale_d = std::sqrt(wrc / L);
            b.scale = (float)b.scale_d;
            if ((rc = upload(g, &b.ab_scatter, t_abs))) break;
            if ((rc = upload(g, &b.ab_gather, t_abg))) break;
            if ((rc = upload(g, &b.az_scatter, t_azs))) break;
            if ((rc = upload(g, &b.az_gather, t_azg))) break;
            blocks.push_back(b);
        }
    }
    col_first[Lc] = (int)blocks.size();
    if (!rc && blocks.empty()) { lsb::set_error("base matrix has no non-zero entry"); rc = LSB_ERR_INVALID; }
    CDesign &d = g->dev;
    d.L = L; d.M = M; d.n = n; d.Lr = Lr; d.Lc = Lc; d.Mr = Mr; d.Mc = Mc; d.w = w; d.h = h;
    d.n_blocks = (int)blocks.size();
    if (!rc) rc = upload(g, &d.blocks, blocks);
    if (!rc) rc = upload(g, &d.col_first, col_first);
    if (!rc) rc = upload(g, &d.W, std::vector<double>(W, W + (size_t)Lr * Lc));
    if (!rc) rc = make_twiddles<float, float2>(g, &g->tb32, h);
    if (!rc) rc = make_twiddles<double, double2>(g, &g->tb64, h);
    if (!rc) {
        cudaError_t e = cudaMalloc((void **)&g->d_queue, kCounters * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(g->d_queue, 0, kCounters * sizeof(unsigned long long));
        if (e != cudaSuccess) { lsb::set_error("queue alloc: %s", cudaGetErrorString(e)); rc = LSB_ERR_CUDA; }
    }
    if (rc) {
        lsb_csparc_design_destroy(g);
        return rc;
    }
    *out = g;
    return LSB_OK;
}

int lsb_csparc_ab_dev_f32(lsb_csparc_design *g, const void *x, void *out, int64_t n_frames, void *stream)
{
    return run_cop<float, false>(g, x, out, n_frames, (cudaStream_t)stream);
}
int lsb_csparc_ab_dev_f64(lsb_csparc_design *g, const void *x, void *out, int64_t n_frames, void *stream)
{
    return run_cop<double, false>(g, x, out, n_frames, (cudaStream_t)stream);
}
int lsb_csparc_az_dev_f32(lsb_csparc_design *g, const void *y, void *out, int64_t n_frames, void *stream)
{
    return run_cop<float, true>(g, y, out, n_frames, (cudaStream_t)stream);
}
int lsb_csparc_az_dev_f64(lsb_csparc_design *g, const void *y, void *out, int64_t n_frames, void *stream)
{
    return run_cop<double, true>(g, y, out, n_frames, (cudaStream_t)stream);
}
int lsb_csparc_amp_dev_f32(lsb_csparc_design *g, const void *y, int64_t n_frames, const lsb_sparc_amp_params *p,
                           const int32_t *idx0, const lsb_sparc_amp_out *out, void *stream)
{
    return run_camp<float>(g, y, n_frames, p, idx0, out, (cudaStream_t)stream);
}
int lsb_csparc_amp_dev_f64(lsb_csparc_design *g, const void *y, int64_t n_frames, const lsb_sparc_amp_params *p,
                           const int32_t *idx0, const lsb_sparc_amp_out *out, void *stream)
{
    return run_camp<double>(g, y, n_frames, p, idx0, out, (cudaStream_t)stream);
}

}  // extern "C"

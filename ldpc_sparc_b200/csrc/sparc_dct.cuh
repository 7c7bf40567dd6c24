// Shared-memory sub-sampled DCT operators for SPARC design matrices (sm_100a).
//
// Reference semantics (sparc_public/sparc.py:648-701 sub_dct, :703-880
// sparc_transforms): block (r, c) of the design matrix is
//     A_rc[i, j] = sqrt(W[r,c]/L) * sqrt(2) * cos(pi * o0[i] * (2*o1[j] + 1) / (2w))
// with o0 (Mr row indices) and o1 (Mc column indices) drawn from 1..w-1.
//   Ab block:  out_r[i] += scale * Y[o0[i]],  Y[k] = sum_n x[n] cos(pi k (2n+1)/(2w)),
//              x = the column block's entries scattered to positions o1  (DCT-II)
//   Az block:  out_c[j] += scale * X[o1[j]],  X[n] = sum_k y[k] cos(pi k (2n+1)/(2w)),
//              y = the row block's entries scattered to positions o0     (DCT-III)
//
// Each length-w real transform runs as ONE complex FFT of H = w/2 points held in
// shared memory (Makhoul's even/odd permutation + the real-FFT split), in place:
// decimation-in-frequency radix-8/4/2 passes, natural order in, bit-reversed
// order out.  The permutations (o0/o1, Makhoul, bit reversal) are folded on the
// host into four uint16 index tables per block, so the device scatters and
// gathers straight between the frame vectors and the FFT buffer:
//   ab_scatter[j] : float slot of x[o1[j]] in the FFT input
//   ab_gather[i]  : float slot of Y[o0[i]] after dct2_post
//   az_scatter[i] : float slot of y[o0[i]] before dct3_pre
//   az_gather[j]  : float slot of X[o1[j]] after the inverse FFT
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lsb_sparc {

template <typename T> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

template <typename T2> __device__ __forceinline__ T2 cadd(T2 a, T2 b) { return {a.x + b.x, a.y + b.y}; }
template <typename T2> __device__ __forceinline__ T2 csub(T2 a, T2 b) { return {a.x - b.x, a.y - b.y}; }
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b)
{
    return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T2> __device__ __forceinline__ T2 cconj(T2 a) { return {a.x, -a.y}; }
// multiply by -i (forward) or +i (inverse)
template <bool INV, typename T2> __device__ __forceinline__ T2 mul_mi(T2 a)
{
    return INV ? T2{-a.y, a.x} : T2{a.y, -a.x};
}

// Twiddle tables of one transform size (device pointers):
//   tw[k] = exp(-2 pi i k / H), k < H          (FFT passes)
//   pt[k] = exp(-i pi k / (2N)), k <= H        (DCT phase), N = 2H = w; the real-FFT
//           split factor exp(-2 pi i k / N) is pt[k]^4, pt[H-k] = e^{-i pi/4} conj(pt[k])
template <typename T>
struct FftTables {
    const typename Cplx<T>::type *tw, *pt;
    const typename Cplx<T>::type *ptr;   // ptr[t] = pt[brev_{h-1}(t)], t < H/2; ptr[H/2] = pt[H/2]
    int h;   // log2(H)
};

// Shared-memory slot of complex element i.  The XOR swizzle spreads the last FFT
// passes (sub-transform sizes 32 and 4, strides 4 and 1) over all banks: without
// it every 64-bit access of those passes is a 4-way conflict per half-warp.  The
// host applies the same map when it builds the scatter/gather tables.
__host__ __device__ __forceinline__ int swz(int i) { return i ^ ((i >> 4) & 3) ^ (((i >> 5) & 3) << 2); }

template <bool INV, typename T2> __device__ __forceinline__ T2 twiddle(const T2 *tw, int idx)
{
    T2 w = __ldg(tw + idx);
    return INV ? cconj(w) : w;
}

// radix-2 DIF butterfly: a' = a + b, b' = (a - b) * w
template <typename T2> __device__ __forceinline__ void bfly(T2 &a, T2 &b, T2 w)
{
    T2 d = csub(a, b);
    a = cadd(a, b);
    b = cmul(d, w);
}
template <typename T2> __device__ __forceinline__ void bfly1(T2 &a, T2 &b)   // w = 1
{
    T2 d = csub(a, b);
    a = cadd(a, b);
    b = d;
}

// One in-place DIF pass of radix 2^LOGR over sub-transforms of size S (S >= R).
// Thread q handles positions g*S + j + m*(S/R), m < R (g = q / (S/R), j = q % (S/R)).
template <typename T, int LOGR, bool INV>
__device__ __forceinline__ void fft_pass(typename Cplx<T>::type *buf, const typename Cplx<T>::type *tw,
                                         int H, int logS, int tid, int nthreads)
{
    typedef typename Cplx<T>::type T2;
    constexpr int R = 1 << LOGR;
    const int logQ = logS - LOGR;           // log2(S/R)
    const int stride = 1 << logQ;
    const int tw_step = H >> logS;          // table step of w_S
    const T kR = (T)0.70710678118654752440;
    for (int q = tid; q < (H >> LOGR); q += nthreads) {
        const int j = q & (stride - 1);
        const int base = ((q >> logQ) << logS) + j;
        T2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = buf[swz(base + m * stride)];
        const T2 w1 = twiddle<INV>(tw, j * tw_step);
        if (LOGR == 1) {
            bfly(v[0], v[1], w1);
        } else if (LOGR == 2) {
            const T2 w2 = cmul(w1, w1);                 // w_S^{2j}
            bfly(v[0], v[2], w1);
            bfly(v[1], v[3], mul_mi<INV>(w1));          // w1 * w_4^1
            bfly(v[0], v[1], w2);
            bfly(v[2], v[3], w2);
        } else {
            const T2 w2 = cmul(w1, w1);                 // w_S^{2j}
            const T2 w4 = cmul(w2, w2);                 // w_S^{4j}
            const T2 r8 = INV ? T2{kR, kR} : T2{kR, -kR};  // w_8^1
            const T2 w1b = cmul(w1, r8);
            bfly(v[0], v[4], w1);
            bfly(v[1], v[5], w1b);
            bfly(v[2], v[6], mul_mi<INV>(w1));          // w1 * w_8^2
            bfly(v[3], v[7], mul_mi<INV>(w1b));         // w1 * w_8^3
            const T2 w2b = mul_mi<INV>(w2);             // w2 * w_4^1
            bfly(v[0], v[2], w2);
            bfly(v[1], v[3], w2b);
            bfly(v[4], v[6], w2);
            bfly(v[5], v[7], w2b);
            bfly(v[0], v[1], w4);
            bfly(v[2], v[3], w4);
            bfly(v[4], v[5], w4);
            bfly(v[6], v[7], w4);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) buf[swz(base + m * stride)] = v[m];
    }
}

// Complete in-place FFT of H = 2^h points (h >= 1).  Ends with __syncthreads().
template <typename T, bool INV>
__device__ void fft_inplace(typename Cplx<T>::type *buf, const FftTables<T> &tb, int tid, int nthreads)
{
    const int h = tb.h, H = 1 << h;
    int n8 = h / 3, rem = h - 3 * n8, n4 = 0, n2 = 0;
    if (rem == 2) n4 = 1;
    else if (rem == 1) { if (n8 >= 1) { n8 -= 1; n4 = 2; } else n2 = 1; }
    int logS = h;
    for (int i = 0; i < n8; ++i) { fft_pass<T, 3, INV>(buf, tb.tw, H, logS, tid, nthreads); logS -= 3; __syncthreads(); }
    for (int i = 0; i < n4; ++i) { fft_pass<T, 2, INV>(buf, tb.tw, H, logS, tid, nthreads); logS -= 2; __syncthreads(); }
    for (int i = 0; i < n2; ++i) { fft_pass<T, 1, INV>(buf, tb.tw, H, logS, tid, nthreads); logS -= 1; __syncthreads(); }
}

// ---------------------------------------------------------------- fixed-size FFT
// Same passes with the transform size known at compile time (the AMP kernel is bound by
// issue slots, ncu: 77 % issue-active, 43 % of its instructions integer): all strides,
// loop counts and twiddle steps become immediates, and because swz() is linear over GF(2)
// (shifts, masks and XORs only) the slot of element base + m*stride is
// swz(base) ^ swz(m*stride) with the second term a constant.  The radix-8 butterfly is a
// trivial-twiddle DFT-8 followed by one multiply per output (w^k, k = 1..7, from three
// table entries), the last pass (stride 1) has no twiddles at all.  Bit-reversed output,
// identical to fft_inplace.
__host__ __device__ constexpr int swz_c(int i) { return i ^ ((i >> 4) & 3) ^ (((i >> 5) & 3) << 2); }
__host__ __device__ constexpr int brev3_c(int m) { return ((m & 1) << 2) | (m & 2) | ((m >> 2) & 1); }

template <bool INV, typename T2> __device__ __forceinline__ T2 mul_w8(T2 a)   // a * exp(-+ i pi/4)
{
    typedef decltype(a.x) T;
    const T kR = (T)0.70710678118654752440;
    return INV ? T2{(a.x - a.y) * kR, (a.x + a.y) * kR} : T2{(a.x + a.y) * kR, (a.y - a.x) * kR};
}

template <typename T, int LOGH, int LOGS, int LOGR, bool INV, int NT>
__device__ __forceinline__ void fft_pass_fixed(typename Cplx<T>::type *buf, const typename Cplx<T>::type *tw, int tid)
{
    typedef typename Cplx<T>::type T2;
    constexpr int H = 1 << LOGH, R = 1 << LOGR, LOGQ = LOGS - LOGR, STRIDE = 1 << LOGQ;
    constexpr int NB = H >> LOGR;                 // butterflies in this pass
    constexpr int TW_STEP = H >> LOGS;
#pragma unroll
    for (int q0 = 0; q0 < NB; q0 += NT) {
        const int q = q0 + tid;
        if (NB % NT != 0 && q >= NB) break;
        const int j = q & (STRIDE - 1);
        const int sb = swz(((q >> LOGQ) << LOGS) | j);
        T2 v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = buf[sb ^ swz_c(m << LOGQ)];
        if (LOGR == 1) {
            bfly1(v[0], v[1]);
        } else if (LOGR == 2) {
            bfly1(v[0], v[2]);
            bfly1(v[1], v[3]);
            v[3] = mul_mi<INV>(v[3]);
            bfly1(v[0], v[1]);
            bfly1(v[2], v[3]);
        } else {
            bfly1(v[0], v[4]);
            bfly1(v[1], v[5]);
            bfly1(v[2], v[6]);
            bfly1(v[3], v[7]);
            v[5] = mul_w8<INV>(v[5]);
            v[6] = mul_mi<INV>(v[6]);
            v[7] = mul_mi<INV>(mul_w8<INV>(v[7]));
            bfly1(v[0], v[2]);
            bfly1(v[1], v[3]);
            bfly1(v[4], v[6]);
            bfly1(v[5], v[7]);
            v[3] = mul_mi<INV>(v[3]);
            v[7] = mul_mi<INV>(v[7]);
            bfly1(v[0], v[1]);
            bfly1(v[2], v[3]);
            bfly1(v[4], v[5]);
            bfly1(v[6], v[7]);
        }
        if (LOGQ > 0) {     // output m holds frequency k = brev(m) of the radix-R DFT: times w_S^(j k)
            const T2 w1 = twiddle<INV>(tw, j * TW_STEP);
            if (LOGR == 1) {
                v[1] = cmul(v[1], w1);
            } else if (LOGR == 2) {
                const T2 w2 = twiddle<INV>(tw, 2 * j * TW_STEP);
                v[1] = cmul(v[1], w2);
                v[2] = cmul(v[2], w1);
                v[3] = cmul(v[3], cmul(w1, w2));
            } else {
                const T2 w2 = twiddle<INV>(tw, 2 * j * TW_STEP);
                const T2 w4 = twiddle<INV>(tw, 4 * j * TW_STEP);
                const T2 w3 = cmul(w1, w2);
                v[1] = cmul(v[1], w4);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], cmul(w2, w4));
                v[4] = cmul(v[4], w1);
                v[5] = cmul(v[5], cmul(w1, w4));
                v[6] = cmul(v[6], w3);
                v[7] = cmul(v[7], cmul(w3, w4));
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) buf[sb ^ swz_c(m << LOGQ)] = v[m];
    }
}

// radix of the pass that starts at sub-transform size 2^LOGS: 8 unless that would leave a
// single radix-2 pass (then two radix-4 passes close the plan), as in fft_inplace
__host__ __device__ constexpr int pass_logr(int logs) { return (logs >= 3 && logs - 3 != 1) ? 3 : (logs >= 2 ? 2 : 1); }

template <typename T, int LOGH, int LOGS, bool INV, int NT>
__device__ __forceinline__ void fft_fixed_from(typename Cplx<T>::type *buf, const typename Cplx<T>::type *tw, int tid)
{
    if constexpr (LOGS > 0) {
        constexpr int LOGR = pass_logr(LOGS);
        fft_pass_fixed<T, LOGH, LOGS, LOGR, INV, NT>(buf, tw, tid);
        __syncthreads();
        fft_fixed_from<T, LOGH, LOGS - LOGR, INV, NT>(buf, tw, tid);
    }
}

// Complete in-place FFT of 2^LOGH points, NT threads.  Ends with __syncthreads().
template <typename T, int LOGH, bool INV, int NT>
__device__ __forceinline__ void fft_fixed(typename Cplx<T>::type *buf, const FftTables<T> &tb, int tid)
{
    fft_fixed_from<T, LOGH, LOGH, INV, NT>(buf, tb.tw, tid);
}

// ---------------------------------------------------------------- staged FFT (buffer in global memory)
// Transforms whose buffer lives in global memory (w = 65536: 256 KB, more than one SM's shared
// memory) would touch L2 once per radix pass.  The same in-place passes are regrouped so that
// a 4096-point tile is staged through shared memory and several passes run on it there:
//   stage A: the passes above sub-transform size 512 (element i = j + 512 m: a tile holds the
//            64 values of m for 64 consecutive j -- rows of 512 B in global memory);
//   stage B: the passes of the last 512-point sub-transforms, 8 of them per contiguous tile.
// Element order, swizzle and twiddles are exactly those of fft_inplace, so the index tables
// and the DCT pre/post-processing are unchanged.  Built for H = 2^15 (two radix-8 passes in
// stage A, three in stage B): 2 L2 round trips instead of 5.
constexpr int kTilePts = 4096;

// one radix-8 DIF pass over a 4096-point tile in shared memory; slot(l) = tile slot of local index l,
// twi(jl) = twiddle-table index of the butterfly whose first local index has offset jl inside its sub-transform
template <typename T, bool INV, typename SlotFn, typename TwFn>
__device__ __forceinline__ void tile_pass8(typename Cplx<T>::type *tile, const typename Cplx<T>::type *tw, int logS,
                                           SlotFn slot, TwFn twi, int tid, int nthreads)
{
    typedef typename Cplx<T>::type T2;
    const int logQ = logS - 3, stride = 1 << logQ;
    for (int q = tid; q < (kTilePts >> 3); q += nthreads) {
        const int jl = q & (stride - 1);
        const int base = ((q >> logQ) << logS) + jl;
        T2 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = tile[slot(base + m * stride)];
        bfly1(v[0], v[4]);
        bfly1(v[1], v[5]);
        bfly1(v[2], v[6]);
        bfly1(v[3], v[7]);
        v[5] = mul_w8<INV>(v[5]);
        v[6] = mul_mi<INV>(v[6]);
        v[7] = mul_mi<INV>(mul_w8<INV>(v[7]));
        bfly1(v[0], v[2]);
        bfly1(v[1], v[3]);
        bfly1(v[4], v[6]);
        bfly1(v[5], v[7]);
        v[3] = mul_mi<INV>(v[3]);
        v[7] = mul_mi<INV>(v[7]);
        bfly1(v[0], v[1]);
        bfly1(v[2], v[3]);
        bfly1(v[4], v[5]);
        bfly1(v[6], v[7]);
        const int ti = twi(jl);
        if (ti != 0) {
            const T2 w1 = twiddle<INV>(tw, ti), w2 = twiddle<INV>(tw, 2 * ti), w4 = twiddle<INV>(tw, 4 * ti);
            const T2 w3 = cmul(w1, w2);
            v[1] = cmul(v[1], w4);
            v[2] = cmul(v[2], w2);
            v[3] = cmul(v[3], cmul(w2, w4));
            v[4] = cmul(v[4], w1);
            v[5] = cmul(v[5], cmul(w1, w4));
            v[6] = cmul(v[6], w3);
            v[7] = cmul(v[7], cmul(w3, w4));
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) tile[slot(base + m * stride)] = v[m];
    }
}

// In-place FFT of 2^15 points held in global memory at gbuf (slots as fft_inplace), staged through the
// 4096-point shared-memory tile.  Ends with __syncthreads().
template <typename T, bool INV>
__device__ void fft_staged_h15(typename Cplx<T>::type *gbuf, typename Cplx<T>::type *tile, const FftTables<T> &tb,
                               int tid, int nthreads)
{
    typedef typename Cplx<T>::type T2;
    const T2 *tw = tb.tw;
    // ---- stage A: passes with sub-transform sizes 2^15 and 2^12 (strides 4096 and 512)
    for (int t = 0; t < 8; ++t) {
        for (int e = tid; e < kTilePts; e += nthreads) tile[e] = gbuf[512 * (e >> 6) + 64 * t + (e & 63)];
        __syncthreads();
        // local index l = 64 m + jj  <->  global element i = 512 m + 64 t + jj; the global swizzle only
        // permutes the low 4 bits as a function of bits 4..6 of i (bit 6 = bit 0 of t)
        auto slot = [t](int l) { return (l & ~63) | (swz(((t & 1) << 6) | (l & 63)) & 63); };
        tile_pass8<T, INV>(tile, tw, 12, slot, [t](int jl) { return 64 * t + (jl & 63) + 512 * (jl >> 6); }, tid, nthreads);
        __syncthreads();
        tile_pass8<T, INV>(tile, tw, 9, slot, [t](int jl) { return (64 * t + jl) * 8; }, tid, nthreads);
        __syncthreads();
        for (int e = tid; e < kTilePts; e += nthreads) gbuf[512 * (e >> 6) + 64 * t + (e & 63)] = tile[e];
        __syncthreads();
    }
    // ---- stage B: the 512-point sub-transforms (sizes 2^9, 2^6, 2^3), 8 per contiguous tile
    for (int c = 0; c < 8; ++c) {
        T2 *src = gbuf + (size_t)c * kTilePts;
        for (int e = tid; e < kTilePts; e += nthreads) tile[e] = src[e];
        __syncthreads();
        auto slot = [](int l) { return swz(l); };
        tile_pass8<T, INV>(tile, tw, 9, slot, [](int jl) { return jl * 64; }, tid, nthreads);
        __syncthreads();
        tile_pass8<T, INV>(tile, tw, 6, slot, [](int jl) { return jl * 512; }, tid, nthreads);
        __syncthreads();
        tile_pass8<T, INV>(tile, tw, 3, slot, [](int) { return 0; }, tid, nthreads);
        __syncthreads();
        for (int e = tid; e < kTilePts; e += nthreads) src[e] = tile[e];
        __syncthreads();
    }
}

__device__ __forceinline__ int bitrev(int k, int h) { return (int)(__brev((unsigned)k) >> (32 - h)); }

// After the forward FFT of the packed sequence: slot brev(k) <- (Y[k], Y[N-k]),
// slot 0 <- (Y[0], Y[H]).  Ends with __syncthreads().
template <typename T>
__device__ void dct2_post(typename Cplx<T>::type *buf, const FftTables<T> &tb, int tid, int nthreads)
{
    typedef typename Cplx<T>::type T2;
    const int h = tb.h, H = 1 << h;
    // Pair index k is visited in bit-reversed order (k = brev_{h-1}(t)): the FFT left
    // Z[k] at slot brev_h(k) = 2t and Z[H-k] at a slot that runs backwards with t, so
    // consecutive lanes touch consecutive slots (in natural k order every lane of a
    // warp hits the same bank).  ptr is pt in that same order, so it loads coalesced.
    const int Hh = H >> 1;
    for (int t = tid; t <= Hh; t += nthreads) {
        const int k = (t == Hh) ? Hh : (h > 1 ? bitrev(t, h - 1) : 0);
        if (k == 0) {
            const T2 z = buf[0];
            buf[0] = T2{z.x + z.y, (z.x - z.y) * (T)0.70710678118654752440};
            continue;
        }
        const int pa = swz(bitrev(k, h)), pb = swz(bitrev(H - k, h));
        const T2 za = buf[pa], zb = buf[pb];
        const T2 ptk = __ldg(tb.ptr + t);
        const T2 pt2 = cmul(ptk, ptk);
        const T2 qk = cmul(pt2, pt2);                                  // q[k] = pt[k]^4
        const T kR = (T)0.70710678118654752440;
        const T2 ptm = cmul(T2{kR, -kR}, cconj(ptk));                  // pt[H-k] = e^{-i pi/4} conj(pt[k])
        // E = (za + conj(zb))/2, O = -i (za - conj(zb))/2
        const T2 E = {(T)0.5 * (za.x + zb.x), (T)0.5 * (za.y - zb.y)};
        const T2 D = {(T)0.5 * (za.x - zb.x), (T)0.5 * (za.y + zb.y)};
        const T2 O = {D.y, -D.x};
        const T2 qO = cmul(qk, O);
        const T2 Vk = cadd(E, qO);                 // V[k]
        const T2 Vm = cconj(csub(E, qO));          // V[H-k]
        const T2 ta = cmul(ptk, Vk);
        buf[pa] = T2{ta.x, -ta.y};
        if (pb != pa) {
            const T2 t2 = cmul(ptm, Vm);
            buf[pb] = T2{t2.x, -t2.y};
        }
    }
    __syncthreads();
}

// Before the inverse FFT: natural slot k holds (c[k], c[N-k]), slot 0 holds
// (c[0], c[H]); rewritten in place with the packed spectrum Zs[k].  Ends with
// __syncthreads().
template <typename T>
__device__ void dct3_pre(typename Cplx<T>::type *buf, const FftTables<T> &tb, int tid, int nthreads)
{
    typedef typename Cplx<T>::type T2;
    const int h = tb.h, H = 1 << h;
    for (int k = tid; k <= (H >> 1); k += nthreads) {
        if (k == 0) {
            const T2 c = buf[0];
            const T wH = c.y * (T)0.70710678118654752440;
            buf[0] = T2{c.x + wH, c.x - wH};
            continue;
        }
        const int kb = H - k;
        const int sa = swz(k), sb = swz(kb);
        const T2 ca = buf[sa], cb = buf[sb];
        const T2 ptk = __ldg(tb.pt + k);
        const T2 pt2 = cmul(ptk, ptk);
        const T2 qk = cmul(pt2, pt2);                                  // q[k] = pt[k]^4
        const T kR = (T)0.70710678118654752440;
        const T2 ptm = cmul(T2{kR, -kR}, cconj(ptk));                  // pt[H-k]
        // W[k] = conj(pt[k]) (c[k] - i c[N-k]) / 2
        const T2 Wa = cmul(cconj(ptk), T2{(T)0.5 * ca.x, (T)-0.5 * ca.y});
        const T2 Wb = cmul(cconj(ptm), T2{(T)0.5 * cb.x, (T)-0.5 * cb.y});
        const T2 A = cadd(Wa, cconj(Wb));
        const T2 B = cmul(csub(Wa, cconj(Wb)), cconj(qk));
        buf[sa] = T2{A.x - B.y, A.y + B.x};                     // A + iB
        if (kb != k) buf[sb] = T2{A.x + B.y, -A.y + B.x};      // conj(A) + i conj(B)
    }
    __syncthreads();
}

// Device view of one non-zero block of the design.
struct BlockDev {
    int r, c;
    float scale;                       // sqrt(2 W[r,c] / L)
    double scale_d;
    const uint16_t *ab_scatter, *ab_gather, *az_scatter, *az_gather;
};

}  // namespace lsb_sparc

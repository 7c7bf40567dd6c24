// Two-frames-per-CTA ("pair") sub-sampled DCT transforms for the fp32 throughput path, sm_100a.
//
// Same mathematics as sparc_dct.cuh (reference sparc_public/sparc.py:648-701 sub_dct,
// :703-880 sparc_transforms); what changes is how the work is laid out for the machine:
//  * every shared-memory slot of the FFT buffer is a float4 (re_A, re_B, im_A, im_B) holding one
//    complex element of TWO frames.  All arithmetic runs on the packed-fp32 pipe forms
//    (FADD2 / FMUL2 / FFMA2, sm_100): one issue slot per two frames, twiddles enter as a scalar
//    broadcast operand.  Index tables, twiddle loads, slot arithmetic and loop control are shared
//    by the two frames, shared-memory accesses are 128-bit.
//  * the forward transform has no zero-fill / scatter: its first radix-8 pass GATHERS its inputs
//    straight from the column block in global memory through an inverse table (0xFFFF = zero);
//  * the transposed transform gathers the row block from shared memory inside the real-FFT
//    pre-processing pass (again an inverse table with a sentinel), no zero-fill / scatter;
//  * real-FFT twiddles come from (0.5 pt, pt^4) tables, the bit-reversed pair slots of the
//    post-processing from a slot table.
// Compile-time transform sizes H = w/2 = 2^10, 2^11, 2^12 (w = 2048, 4096, 8192).
#pragma once
#include "sparc_dct.cuh"

namespace lsb_sparc {
namespace pairfft {

constexpr int kPairThreads = 256;
constexpr uint16_t kNone = 0xFFFFu;

typedef float2 P;                  // .x: frame A, .y: frame B
struct CP { P re, im; };           // one complex value of each frame

__device__ __forceinline__ P padd(P a, P b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ P psub(P a, P b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ P pneg(P a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ P pscale(P a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ P pfmas(P a, float s, P c) { return __ffma2_rn(a, make_float2(s, s), c); }   // a*s + c
__device__ __forceinline__ P pmul(P a, P b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ P pfma(P a, P b, P c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ CP cadd(CP a, CP b) { return {padd(a.re, b.re), padd(a.im, b.im)}; }
__device__ __forceinline__ CP csub(CP a, CP b) { return {psub(a.re, b.re), psub(a.im, b.im)}; }
// a * w, w one complex scalar shared by both frames
__device__ __forceinline__ CP cmulw(CP a, float2 w)
{
    return {pfmas(a.im, -w.y, pscale(a.re, w.x)), pfmas(a.im, w.x, pscale(a.re, w.y))};
}
template <bool INV> __device__ __forceinline__ CP mul_mi(CP a)          // times -i (forward) / +i (inverse)
{
    return INV ? CP{pneg(a.im), a.re} : CP{a.im, pneg(a.re)};
}
template <bool INV> __device__ __forceinline__ CP mul_w8(CP a)          // times exp(-+ i pi/4)
{
    const float kR = 0.70710678118654752440f;
    return INV ? CP{pscale(psub(a.re, a.im), kR), pscale(padd(a.re, a.im), kR)}
               : CP{pscale(padd(a.re, a.im), kR), pscale(psub(a.im, a.re), kR)};
}
__device__ __forceinline__ void bfly1(CP &a, CP &b)
{
    const CP d = csub(a, b);
    a = cadd(a, b);
    b = d;
}
__device__ __forceinline__ CP ld(const float4 *buf, int s)
{
    const float4 v = buf[s];
    return {make_float2(v.x, v.y), make_float2(v.z, v.w)};
}
__device__ __forceinline__ void st(float4 *buf, int s, CP v) { buf[s] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }

__device__ __forceinline__ float2 cmul_s(float2 a, float2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// Slot swizzle of the float4 buffer.  A 128-bit warp access is served 8 lanes at a time, so the 8
// lanes of a phase must hit 8 different slots mod 8.  Lanes walk consecutive slots in the passes of
// stride >= 8 (no swizzle needed); the last passes walk bits {0,1,LOGS_a} and {LOGR.., } of the
// index, which the XOR terms below fold into the low three bits (linear over GF(2), so
// swz(base + m*stride) = swz(base) ^ swz(m*stride) whenever the two have disjoint bits).
//   h = 12 (8,8,8,8):  bits 3,4,5 -> 0,1,2
//   h = 11 (8,8,8,4):  bits 3,4 -> 0,1; bit 5 -> 2
//   h = 10 (8,8,4,4):  bit 3 -> 0; bit 4 -> 1 and 2
template <int LOGH> __host__ __device__ constexpr int pswz(int i)
{
    return LOGH == 12 ? (i ^ ((i >> 3) & 7))
         : LOGH == 11 ? (i ^ ((i >> 3) & 3) ^ (((i >> 5) & 1) << 2))
                      : (i ^ ((i >> 3) & 1) ^ (((i >> 4) & 1) * 6));
}
inline int pswz_rt(int logh, int i) { return logh == 12 ? pswz<12>(i) : logh == 11 ? pswz<11>(i) : pswz<10>(i); }

// One radix-8 / radix-4 DIF pass over sub-transforms of size 2^LOGS on the pair buffer.
// GATHER: first pass of the forward transform -- inputs come from the column block in global
// memory (x.get(i): entry i of frame A / frame B) through src[q][2R] (uint16, kNone = zero).
struct NoSrc {
    __device__ __forceinline__ P get(unsigned) const { return make_float2(0.f, 0.f); }
};
template <int LOGH, int LOGS, int LOGR, bool INV, bool GATHER, typename Src = NoSrc>
__device__ __forceinline__ void pass(float4 *buf, const float2 *__restrict__ tw, int tid,
                                     const uint16_t *__restrict__ src = nullptr, const Src x = Src())
{
    constexpr int H = 1 << LOGH, R = 1 << LOGR, LOGQ = LOGS - LOGR;
    constexpr int NB = H >> LOGR;
    constexpr int TW_STEP = H >> LOGS;
    static_assert(!GATHER || (LOGS == LOGH && LOGR == 3), "the gathering pass is the first, radix-8 one");
#pragma unroll
    for (int q0 = 0; q0 < NB; q0 += kPairThreads) {
        const int q = q0 + tid;
        if (NB % kPairThreads != 0 && q >= NB) break;
        const int j = q & ((1 << LOGQ) - 1);
        const int sb = pswz<LOGH>(((q >> LOGQ) << LOGS) | j);
        float2 w1 = {1.f, 0.f}, w2 = w1, w4 = w1;
        if (LOGQ > 0) {
            w1 = __ldg(tw + j * TW_STEP);
            w2 = __ldg(tw + 2 * j * TW_STEP);
            if (LOGR == 3) w4 = __ldg(tw + 4 * j * TW_STEP);
            if (INV) { w1.y = -w1.y; w2.y = -w2.y; w4.y = -w4.y; }
        }
        CP v[R];
        if (GATHER) {
            const uint4 *tp = reinterpret_cast<const uint4 *>(src) + 2 * q;
            const uint4 ta = __ldg(tp), tb = __ldg(tp + 1);
            const unsigned e[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const unsigned ir = e[m] & 0xFFFFu, ii = e[m] >> 16;
                v[m].re = ir != kNone ? x.get(ir) : make_float2(0.f, 0.f);
                v[m].im = ii != kNone ? x.get(ii) : make_float2(0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) v[m] = ld(buf, sb ^ pswz<LOGH>(m << LOGQ));
        }
        if (LOGR == 2) {
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
        if (LOGQ > 0) {     // output m holds frequency brev(m) of the radix-R DFT: times w_S^(j brev(m))
            if (LOGR == 2) {
                v[1] = cmulw(v[1], w2);
                v[2] = cmulw(v[2], w1);
                v[3] = cmulw(v[3], cmul_s(w1, w2));
            } else {
                const float2 w3 = cmul_s(w1, w2);
                v[1] = cmulw(v[1], w4);
                v[2] = cmulw(v[2], w2);
                v[3] = cmulw(v[3], cmul_s(w2, w4));
                v[4] = cmulw(v[4], w1);
                v[5] = cmulw(v[5], cmul_s(w1, w4));
                v[6] = cmulw(v[6], w3);
                v[7] = cmulw(v[7], cmul_s(w3, w4));
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) st(buf, sb ^ pswz<LOGH>(m << LOGQ), v[m]);
    }
}

// passes from sub-transform size 2^LOGS down (radix plan of sparc_dct.cuh: pass_logr); a barrier after each
template <int LOGH, int LOGS, bool INV>
__device__ __forceinline__ void passes_from(float4 *buf, const float2 *tw, int tid)
{
    if constexpr (LOGS > 0) {
        constexpr int LOGR = pass_logr(LOGS);
        pass<LOGH, LOGS, LOGR, INV, false>(buf, tw, tid);
        __syncthreads();
        passes_from<LOGH, LOGS - LOGR, INV>(buf, tw, tid);
    }
}

// Tables of one transform size for the pair kernels (device pointers).
struct PairTables {
    const float2 *tw;          // exp(-2 pi i k / H), k < H
    const float4 *pre;         // [H/2 + 1] natural k:        (0.5 pt[k], pt[k]^4)
    const float4 *post;        // [H/2 + 1] bit-reversed t:   (0.5 pt[k], pt[k]^4), k = brev_{h-1}(t), last = H/2
    const uint32_t *post_slot; // [H/2 + 1] slot of Z[k] | slot of Z[H-k] << 16 (swizzled, after the FFT)
};

// One non-zero block of the design for the pair kernels.
struct PairBlock {
    int r, c;
    float scale;                   // sqrt(2 W[r,c] / L)
    const uint16_t *fsrc;          // [H/8][16]  forward first-pass sources (index into the column block)
    const uint16_t *fgat;          // [Mr]       float2-view slot of Y[o0[i]] after post()
    const uint16_t *tsrc;          // [H/2+1][4] transposed sources: c[k], c[N-k], c[H-k], c[H+k] (index into the row block)
    const uint16_t *tgat;          // [Mc]       float2-view slot of X[o1[j]] after the inverse FFT
};

// (0.5 pt[H-k]) from (0.5 pt[k]): pt[H-k] = e^{-i pi/4} conj(pt[k])
__device__ __forceinline__ float2 half_pt_mirror(float2 hp)
{
    const float kR = 0.70710678118654752440f;
    return {kR * (hp.x - hp.y), -kR * (hp.x + hp.y)};
}

// ---- forward: buf <- DCT-II spectrum of the column block x2 (both frames); slot brev(k) ends up
// holding (Y[k], Y[N-k]).  Starts by overwriting buf: the caller makes sure nobody still reads it.
// Ends with __syncthreads().
template <int LOGH, typename Src>
__device__ __forceinline__ void forward(float4 *buf, const PairTables &tb, const PairBlock &b, const Src x, int tid)
{
    constexpr int H = 1 << LOGH, Hh = H >> 1;
    pass<LOGH, LOGH, 3, false, true, Src>(buf, tb.tw, tid, b.fsrc, x);
    __syncthreads();
    passes_from<LOGH, LOGH - 3, false>(buf, tb.tw, tid);
    // real-FFT split + DCT phase, pairs visited in bit-reversed order (consecutive lanes, consecutive slots)
    for (int t = tid; t <= Hh; t += kPairThreads) {
        if (t == 0) {
            const CP z = ld(buf, 0);
            st(buf, 0, CP{padd(z.re, z.im), pscale(psub(z.re, z.im), 0.70710678118654752440f)});
            continue;
        }
        const uint32_t sl = __ldg(tb.post_slot + t);
        const float4 tq = __ldg(tb.post + t);
        const int pa = (int)(sl & 0xFFFFu), pb = (int)(sl >> 16);
        const CP za = ld(buf, pa), zb = ld(buf, pb);
        const float2 hp = {tq.x, tq.y}, q = {tq.z, tq.w};
        // E2 = za + conj(zb), O2 = -i (za - conj(zb))  (twice the even / odd spectra; the 1/2 is in hp)
        const CP E2 = {padd(za.re, zb.re), psub(za.im, zb.im)};
        const CP O2 = {padd(za.im, zb.im), psub(zb.re, za.re)};
        const CP qO = cmulw(O2, q);
        const CP Vk = cadd(E2, qO);                              // 2 V[k]
        const CP ta = cmulw(Vk, hp);
        st(buf, pa, CP{ta.re, pneg(ta.im)});
        if (pb != pa) {
            const CP Vm = {psub(E2.re, qO.re), psub(qO.im, E2.im)};   // 2 V[H-k] = conj(E2 - qO)
            const CP t2 = cmulw(Vm, half_pt_mirror(hp));
            st(buf, pb, CP{t2.re, pneg(t2.im)});
        }
    }
    __syncthreads();
}

// ---- transposed: buf <- DCT-III of the row block y2 (shared memory, float2 per entry) scattered to
// its rows; the inverse FFT leaves X in bit-reversed slots (tgat knows them).
// Starts by overwriting buf.  Ends with __syncthreads().
template <int LOGH>
__device__ __forceinline__ void transposed(float4 *buf, const PairTables &tb, const PairBlock &b, const float2 *y2, int tid)
{
    constexpr int H = 1 << LOGH, Hh = H >> 1;
    const uint2 *src = reinterpret_cast<const uint2 *>(b.tsrc);
    for (int k = tid; k <= Hh; k += kPairThreads) {
        const uint2 e = __ldg(src + k);
        const unsigned a0 = e.x & 0xFFFFu, a1 = e.x >> 16, b0 = e.y & 0xFFFFu, b1 = e.y >> 16;
        const P zero = make_float2(0.f, 0.f);
        const CP ca = {a0 != kNone ? y2[a0] : zero, a1 != kNone ? y2[a1] : zero};     // (c[k], c[N-k])
        if (k == 0) {
            const P wH = pscale(ca.im, 0.70710678118654752440f);
            st(buf, 0, CP{padd(ca.re, wH), psub(ca.re, wH)});
            continue;
        }
        const CP cb = {b0 != kNone ? y2[b0] : zero, b1 != kNone ? y2[b1] : zero};     // (c[H-k], c[H+k])
        const float4 tq = __ldg(tb.pre + k);
        const float2 hp = {tq.x, tq.y}, q = {tq.z, tq.w};
        // W[k] = conj(pt[k]) (c[k] - i c[N-k]) / 2 = conj(hp * ca); A = W[k] + conj(W[H-k]), B = (W[k] - conj(W[H-k])) conj(q)
        const CP U = cmulw(ca, hp), V = cmulw(cb, half_pt_mirror(hp));
        const CP A = {padd(U.re, V.re), psub(V.im, U.im)};
        const CP D = {psub(U.re, V.re), pneg(padd(U.im, V.im))};
        const CP B = cmulw(D, make_float2(q.x, -q.y));
        st(buf, pswz<LOGH>(k), CP{psub(A.re, B.im), padd(A.im, B.re)});                 // A + iB
        if (k != Hh) st(buf, pswz<LOGH>(H - k), CP{padd(A.re, B.im), psub(B.re, A.im)});   // conj(A) + i conj(B)
    }
    __syncthreads();
    passes_from<LOGH, LOGH, true>(buf, tb.tw, tid);
}

}  // namespace pairfft
}  // namespace lsb_sparc

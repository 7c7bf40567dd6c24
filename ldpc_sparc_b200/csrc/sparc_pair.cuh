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
//  * real-FFT twiddles come from a (0.5 pt) table (pt^4 by two squarings), the bit-reversed pair
//    slots of the post-processing from a slot table;
//  * every phase ISSUES the global table loads of the NEXT phase before it starts its own work, so
//    their L1 / L2 latency runs under the phase and its closing barrier instead of behind it (the
//    kernel keeps only 6 warps per scheduler: ncu showed 30 % of the stall samples on table loads);
//    index tables are laid out so that one 64- / 128-bit load per thread fetches 4 / 8 entries and
//    are read with L1::no_allocate, which leaves the small L1 (228 KB minus 3 CTAs of shared memory)
//    to the twiddle and real-FFT tables that every block reuses.
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

// LSB_PAIR_FUSED (bit 0: forward, bit 1: transposed; default 3), H = 2^11 only: the real-FFT post-processing of the forward transform runs inside its
// last radix-4 pass and the pre-processing of the transposed transform inside its first pass, which is then a radix-4
// pass as well (plan 4,8,8,8 instead of 8,8,8,4).  A radix-4 butterfly of the H-point DIF FFT holds the frequencies
// b + {0,1,2,3} H/4 (last pass; b = brev(q)) resp. the inputs j + {0,1,2,3} H/4 (first pass), and the mirror
// H - k of each of them lies in the butterfly of H/4 - b resp. H/4 - j: a thread that takes both butterflies has
// every (k, H-k) pair of the split in registers.  Each fusion saves one read and one write of the whole buffer and
// one CTA barrier per transform (tools/model_pair_fused.py is the numpy model of the index conventions).
#ifndef LSB_PAIR_FUSED
#define LSB_PAIR_FUSED 3
#endif
template <int LOGH> __host__ __device__ constexpr bool fused_fwd() { return (LSB_PAIR_FUSED & 1) && LOGH == 11; }
template <int LOGH> __host__ __device__ constexpr bool fused_inv() { return (LSB_PAIR_FUSED & 2) && LOGH == 11; }
// slot swizzle of the transposed transform: its fused plan ends with a radix-8 pass of stride 1 (bits 3,4,5 -> 0,1,2)
template <int LOGH> __host__ __device__ constexpr int pswz_t(int i) { return fused_inv<LOGH>() ? (i ^ ((i >> 3) & 7)) : pswz<LOGH>(i); }
inline int pswz_t_rt(int logh, int i) { return logh == 11 ? pswz_t<11>(i) : pswz_rt(logh, i); }
inline bool fused_fwd_rt(int logh) { return logh == 11 && fused_fwd<11>(); }
inline bool fused_inv_rt(int logh) { return logh == 11 && fused_inv<11>(); }
// LSB_PAIR_STAGE (default 1), H = 2^11: the index table of the NEXT transform (32 bytes per thread) travels global ->
// shared memory with cp.async into a staging area behind the FFT buffer instead of waiting in registers: the kernel
// runs at its register limit, and prefetched table registers that live across the denoiser / the row gather were
// spilled -- the spill store waits for the load, which defeats the prefetch (ncu: long-scoreboard stalls on STL / LDL).
// A warp's 16-byte copies read consecutive addresses (an LDGSTS costs one shared-memory wavefront per 32-byte sector it
// receives: 31 per warp instruction with a 32-byte stride, 11 contiguous).  Measured alternative: one 8 KB
// cp.async.bulk per table with an mbarrier (thread 0 issues, all threads poll) -- 65.9 k against 68.8 k frames/s.
#ifndef LSB_PAIR_STAGE
#define LSB_PAIR_STAGE 1
#endif
template <int LOGH> __host__ __device__ constexpr bool stage_fwd() { return LSB_PAIR_STAGE && LOGH == 11; }
template <int LOGH> __host__ __device__ constexpr bool stage_inv() { return LSB_PAIR_STAGE && fused_inv<LOGH>(); }
// float4 slots of the staging area that follows the FFT buffer: [2][256] (thread t owns slots t and 256 + t)
template <int LOGH> __host__ __device__ constexpr int stage_slots() { return (stage_fwd<LOGH>() || stage_inv<LOGH>()) ? 2 * kPairThreads : 0; }
inline int stage_slots_rt(int logh) { return logh == 11 ? stage_slots<11>() : 0; }
template <int LOGH> __device__ __forceinline__ float4 *stage_of(float4 *buf) { return buf + (1 << LOGH); }
// thread t: entries t and 256 + t of a [2][256] uint4 table -> its two staging slots (it alone reads them back)
template <int LOGH> __device__ __forceinline__ void stage_issue(float4 *buf, const void *table, int tid)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(table);
    float4 *stage = stage_of<LOGH>(buf);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(stage + tid)), "l"(src + tid) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(stage + kPairThreads + tid)),
                 "l"(src + kPairThreads + tid)
                 : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// this thread's own 32 bytes have landed: no barrier needed
template <int LOGH> __device__ __forceinline__ void stage_wait(float4 *) { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int LOGH, bool TR> __host__ __device__ constexpr int pswz_sel(int i) { return TR ? pswz_t<LOGH>(i) : pswz<LOGH>(i); }

// ---- global loads that do not allocate in L1 (index tables: read once per block)
__device__ __forceinline__ uint4 ldg_na(const uint4 *p)
{
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_na(const uint2 *p)
{
    uint2 v;
    asm("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- pass plan: radix of sparc_dct.cuh's pass_logr; 4 passes for LOGH = 10 (8,8,4,4), 11 (8,8,8,4), 12 (8,8,8,8)
template <int LOGH> __host__ __device__ constexpr int plan_logs(int i)
{
    int s = LOGH;
    for (int k = 0; k < i; ++k) s -= pass_logr(s);
    return s;
}
template <int LOGH_, int LOGS_, int LOGR_>
struct PassX {                                                              // one DIF pass: sub-transform size 2^LOGS, radix 2^LOGR
    static constexpr int LOGH = LOGH_, LOGS = LOGS_, LOGR = LOGR_, LOGQ = LOGS - LOGR;
    static constexpr int NB = (1 << LOGH) >> LOGR;                          // butterflies of the pass
    static constexpr int NQ = NB > kPairThreads ? NB / kPairThreads : 1;    // per thread
    static constexpr int TW_STEP = (1 << LOGH) >> LOGS;
};
template <int LOGH, int I>
struct PassInfo : PassX<LOGH, plan_logs<LOGH>(I), pass_logr(plan_logs<LOGH>(I))> {};

template <int NQ> struct TwPre { float2 w1[NQ], w2[NQ], w4[NQ]; };

// issue the twiddle loads of a pass (nothing for a pass of stride 1: no twiddles)
template <typename PI, bool INV>
__device__ __forceinline__ TwPre<PI::NQ> tw_issue_p(const float2 *__restrict__ tw, int tid)
{
    TwPre<PI::NQ> t;
#pragma unroll
    for (int i = 0; i < PI::NQ; ++i) {
        t.w1[i] = t.w2[i] = t.w4[i] = make_float2(1.f, 0.f);
        const int q = tid + i * kPairThreads;
        if (PI::LOGQ > 0 && q < PI::NB) {
            const int j = q & ((1 << PI::LOGQ) - 1);
            t.w1[i] = __ldg(tw + j * PI::TW_STEP);
            t.w2[i] = __ldg(tw + 2 * j * PI::TW_STEP);
            if (PI::LOGR == 3) t.w4[i] = __ldg(tw + 4 * j * PI::TW_STEP);
            if (INV) { t.w1[i].y = -t.w1[i].y; t.w2[i].y = -t.w2[i].y; t.w4[i].y = -t.w4[i].y; }
        }
    }
    return t;
}
template <int LOGH, int I, bool INV>
__device__ __forceinline__ TwPre<PassInfo<LOGH, I>::NQ> tw_issue(const float2 *__restrict__ tw, int tid)
{
    return tw_issue_p<PassInfo<LOGH, I>, INV>(tw, tid);
}

template <int LOGR, bool INV> __device__ __forceinline__ void radix_dft(CP *v)
{
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
}
// output m of the radix-R DFT holds frequency brev(m): times w_S^(j brev(m))
template <int LOGR> __device__ __forceinline__ void radix_twiddle(CP *v, float2 w1, float2 w2, float2 w4)
{
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

// One in-place DIF pass on the pair buffer with its twiddles already in registers (TR: slots of the transposed plan).
template <typename PI, bool INV, bool TR>
__device__ __forceinline__ void pass_run_p(float4 *buf, const TwPre<PI::NQ> &t, int tid)
{
    constexpr int R = 1 << PI::LOGR, LOGH = PI::LOGH;
#pragma unroll
    for (int i = 0; i < PI::NQ; ++i) {
        const int q = tid + i * kPairThreads;
        if (PI::NB < kPairThreads && q >= PI::NB) break;
        const int j = q & ((1 << PI::LOGQ) - 1);
        const int sb = pswz_sel<LOGH, TR>(((q >> PI::LOGQ) << PI::LOGS) | j);
        CP v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = ld(buf, sb ^ pswz_sel<LOGH, TR>(m << PI::LOGQ));
        radix_dft<PI::LOGR, INV>(v);
        if (PI::LOGQ > 0) radix_twiddle<PI::LOGR>(v, t.w1[i], t.w2[i], t.w4[i]);
#pragma unroll
        for (int m = 0; m < R; ++m) st(buf, sb ^ pswz_sel<LOGH, TR>(m << PI::LOGQ), v[m]);
    }
}
template <int LOGH, int I, bool INV>
__device__ __forceinline__ void pass_run(float4 *buf, const TwPre<PassInfo<LOGH, I>::NQ> &t, int tid)
{
    pass_run_p<PassInfo<LOGH, I>, INV, false>(buf, t, tid);
}
// barrier of the 64 threads of warps 2k, 2k+1 (named barriers 1..4)
__device__ __forceinline__ void sync_warp_pair(int tid)
{
    switch (tid >> 6) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}

// Tables of one transform size for the pair kernels (device pointers).
struct PairTables {
    const float2 *tw;          // exp(-2 pi i k / H), k < H
    const float2 *pre;         // [H/2 + 1] natural k:       0.5 pt[k],  pt[k] = exp(-i pi k / (2N))
    const float2 *post;        // [H/2 + 1] bit-reversed t:  0.5 pt[k],  k = brev_{h-1}(t), last entry k = H/2
    const uint32_t *post_slot; // [H/2 + 1] slot of Z[k] | slot of Z[H-k] << 16 (swizzled, after the FFT)
    // fused_inv<LOGH>() / fused_fwd<LOGH>(): pre / post are [256][4] instead, 0.5 pt[k] of the four mirror pairs of thread t (Q = H/4):
    //   pre:  k = t, t + Q, Q - t, 2Q - t;            post: k = b, b + Q, Q - b, 2Q - b with b = brev_{h-2}(2 t);
    //   row 0 of both: k = Q/2, 3Q/2, Q, H/2 (the two self-mirrored butterflies; k = 0 needs no factor)
};

// One non-zero block of the design for the pair kernels.  "Chunked" tables hold the entries of
// element tid + 256 u (u = 4 c .. 4 c + 3) as one uint2 at [c][tid], padded with kNone.
struct PairBlock {
    int r, c;
    float scale;                   // sqrt(2 W[r,c] / L)
    int gat_chunks;                // uint4 chunks (4 slots) per thread of fgat
    const uint4 *fsrc;             // [2][H/8]     forward first-pass sources of butterfly q: elements m = 0..3 | 4..7, (re | im << 16)
    const uint4 *fgat;             // [gat_chunks][256] x 4 slots: row i | float2-view slot of Y[o0[i]] << 16 (after the
                                   //              post-processing), 0xFFFFFFFF = empty.  The rows are scheduled so that the 16
                                   //              lanes of a half-warp hit 16 different bank pairs in the buffer AND in the
                                   //              residual (edge colouring, build_pair_block): no conflict wavefronts
    const uint2 *tsrc;             // [H/2+1]      transposed sources of pair k: c[k], c[N-k], c[H-k], c[H+k] (rows of the block)
                                   //              fused_inv: [2][256] uint4 -- pairs 0,1 then pairs 2,3 of thread t, k order of PairTables::pre --, then the entry of k = 0
    const uint2 *tgat;             // [Mc] float2-view slot of X[o1[j]] after the inverse FFT, column order (thread t of
                                   //              chunk c reads columns 1024 c + 4 t .. + 3 as one uint2), padded with kNone
};
__host__ __device__ constexpr int chunk_count(int n) { return (n + 4 * kPairThreads - 1) / (4 * kPairThreads); }

// 0.5 pt[H-k] from 0.5 pt[k]: pt[H-k] = e^{-i pi/4} conj(pt[k]);  q = pt^4 = 16 (0.5 pt)^4
__device__ __forceinline__ float2 half_pt_mirror(float2 hp)
{
    const float kR = 0.70710678118654752440f;
    return {kR * (hp.x - hp.y), -kR * (hp.x + hp.y)};
}
__device__ __forceinline__ float2 q_from_half_pt(float2 hp)
{
    const float2 s1 = {hp.x * hp.x - hp.y * hp.y, 2.f * hp.x * hp.y};
    return {16.f * (s1.x * s1.x - s1.y * s1.y), 32.f * s1.x * s1.y};
}

#ifndef LSB_PAIR_PREPF
#define LSB_PAIR_PREPF 5
#endif
#ifndef LSB_PAIR_GATPF
#define LSB_PAIR_GATPF 2
#endif
constexpr int kPrePF = LSB_PAIR_PREPF;      // pairs per thread whose tables are prefetched (H/2 + 1 = 1025 pairs at H = 2048)
constexpr int kGatPF = LSB_PAIR_GATPF;      // (row, slot) chunks of the row gather (4 slots each) prefetched

// ------------------------------------------------------------------ forward
template <int LOGH> struct FwdIn {                       // what the gathering first pass needs
    typedef PassInfo<LOGH, 0> PI;
    uint4 ta[PI::NQ], tb[PI::NQ];
    TwPre<PI::NQ> tw;
};
template <int LOGH>
__device__ __forceinline__ FwdIn<LOGH> forward_issue(const PairTables &tb, const PairBlock &b, int tid, float4 *buf)
{
    typedef PassInfo<LOGH, 0> PI;
    FwdIn<LOGH> in;
    if constexpr (stage_fwd<LOGH>()) {
        static_assert(PI::NQ == 1 && PI::NB == kPairThreads, "one first-pass butterfly per thread: fsrc is [2][256] uint4");
        stage_issue<LOGH>(buf, b.fsrc, tid);
        return in;
    }
#pragma unroll
    for (int i = 0; i < PI::NQ; ++i) {
        const int q = tid + i * kPairThreads;
        if (q < PI::NB) {
            in.ta[i] = ldg_na(b.fsrc + q);
            in.tb[i] = ldg_na(b.fsrc + PI::NB + q);
        } else {
            in.ta[i] = in.tb[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        }
    }
    in.tw = tw_issue<LOGH, 0, false>(tb.tw, tid);
    return in;
}

__device__ __forceinline__ void ld4_f2(const float2 *p, float2 *o)         // four float2 as two 16-byte loads
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    o[0] = make_float2(a.x, a.y); o[1] = make_float2(a.z, a.w); o[2] = make_float2(b.x, b.y); o[3] = make_float2(b.z, b.w);
}
struct PostPre { uint32_t slot[kPrePF]; float2 hp[kPrePF]; };
template <int LOGH> __device__ __forceinline__ PostPre post_issue(const PairTables &tb, int tid)
{
    constexpr int Hh = 1 << (LOGH - 1);
    PostPre p;
    if constexpr (fused_fwd<LOGH>()) {
        ld4_f2(tb.post + 4 * tid, p.hp);
        return p;
    }
#pragma unroll
    for (int i = 0; i < kPrePF; ++i) {
        const int t = tid + i * kPairThreads;
        p.slot[i] = t <= Hh ? __ldg(tb.post_slot + t) : 0u;
        p.hp[i] = t <= Hh ? __ldg(tb.post + t) : make_float2(0.f, 0.f);
    }
    return p;
}
__device__ __forceinline__ void post_pair(float4 *buf, uint32_t sl, float2 hp)
{
    const int pa = (int)(sl & 0xFFFFu), pb = (int)(sl >> 16);
    const CP za = ld(buf, pa), zb = ld(buf, pb);
    const float2 q = q_from_half_pt(hp);
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

// the same split on a mirror pair held in registers: (Z[k], Z[H-k]) -> ((Y[k], Y[N-k]), (Y[H-k], Y[H+k]))
__device__ __forceinline__ void post_pair_reg(CP &za, CP &zb, float2 hp)
{
    const float2 q = q_from_half_pt(hp);
    const CP E2 = {padd(za.re, zb.re), psub(za.im, zb.im)};
    const CP O2 = {padd(za.im, zb.im), psub(zb.re, za.re)};
    const CP qO = cmulw(O2, q);
    const CP ta = cmulw(cadd(E2, qO), hp);
    const CP Vm = {psub(E2.re, qO.re), psub(qO.im, E2.im)};
    const CP t2 = cmulw(Vm, half_pt_mirror(hp));
    za = CP{ta.re, pneg(ta.im)};
    zb = CP{t2.re, pneg(t2.im)};
}
// Last pass of the forward transform at H = 2^11 (radix 4, stride 1) with the split fused: thread t > 0 takes the
// butterflies q1 = 2 t (frequencies b + m' Q, b = brev(q1) < Q/2) and q2 = brev(Q - b); position 4 q + m holds the
// frequency brev(q) + brev_2(m) Q.  Thread 0 takes the self-mirrored butterflies of b = 0 and b = Q/2.
__device__ __forceinline__ CP csel(bool c, CP a, CP b)
{
    return {make_float2(c ? a.re.x : b.re.x, c ? a.re.y : b.re.y), make_float2(c ? a.im.x : b.im.x, c ? a.im.y : b.im.y)};
}
template <int LOGH>
__device__ __forceinline__ void forward_last_fused(float4 *buf, const float2 *hp, int tid)
{
    constexpr int Q = 1 << (LOGH - 2);
    const bool t0 = tid == 0;
    const int q1 = 2 * tid;
    const int b = (int)(__brev((unsigned)q1) >> (34 - LOGH));
    const int q2 = t0 ? 1 : (int)(__brev((unsigned)(Q - b)) >> (34 - LOGH));
    CP v[4], u[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = ld(buf, pswz<LOGH>(4 * q1 + m));
#pragma unroll
    for (int m = 0; m < 4; ++m) u[m] = ld(buf, pswz<LOGH>(4 * q2 + m));
    radix_dft<2, false>(v);
    radix_dft<2, false>(u);
    // mirror pairs (first, second) of thread t > 0:  (v0,u3) k = b, (v2,u1) k = b + Q, (u0,v3) k = Q - b, (u2,v1) k = 2Q - b;
    // of thread 0 (table row Q/2, 3Q/2, Q, H/2):      (u0,u3),       (u2,u1),           (v2,v3),           (v1,v1);  v0 = Z[0].
    // Only the first members differ: select them (no divergent path: the warp of thread 0 would run both).
    CP a0 = csel(t0, u[0], v[0]), a1 = csel(t0, u[2], v[2]), a2 = csel(t0, v[2], u[0]), a3 = csel(t0, v[1], u[2]);
    const CP z = v[0];
    post_pair_reg(a0, u[3], hp[0]);
    post_pair_reg(a1, u[1], hp[1]);
    post_pair_reg(a2, v[3], hp[2]);
    post_pair_reg(a3, v[1], hp[3]);
    const CP z0 = {padd(z.re, z.im), pscale(psub(z.re, z.im), 0.70710678118654752440f)};   // k = 0: (Y[0], Y[H])
    v[0] = csel(t0, z0, a0);
    v[2] = csel(t0, a2, a1);
    u[0] = csel(t0, a0, a2);
    u[2] = csel(t0, a1, a3);
    v[1] = csel(t0, a3, v[1]);
#pragma unroll
    for (int m = 0; m < 4; ++m) st(buf, pswz<LOGH>(4 * q1 + m), v[m]);
#pragma unroll
    for (int m = 0; m < 4; ++m) st(buf, pswz<LOGH>(4 * q2 + m), u[m]);
}

struct GatPre { uint4 c[kGatPF]; };
__device__ __forceinline__ GatPre gather_issue(const uint4 *__restrict__ tab, int chunks, int tid)
{
    GatPre g;
#pragma unroll
    for (int c = 0; c < kGatPF; ++c)
        g.c[c] = c < chunks ? ldg_na(tab + c * kPairThreads + tid) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    return g;
}

// buf <- DCT-II spectrum of the column block x (both frames): slot brev(k) ends up holding (Y[k], Y[N-k]).
// `in` was issued by forward_issue() for this block.  Overwrites buf from the start (the caller makes sure
// nobody still reads it), ends after the barrier that closes the post-processing and returns the first
// gather ids of the block's rows (issued under the post-processing).
template <int LOGH, typename Src>
__device__ __forceinline__ GatPre forward_run(float4 *buf, const PairTables &tb, const PairBlock &b, const Src x,
                                              const FwdIn<LOGH> &in, int tid)
{
    constexpr int H = 1 << LOGH, Hh = H >> 1;
    typedef PassInfo<LOGH, 0> P0;
    static_assert(P0::LOGR == 3 && PassInfo<LOGH, 3>::LOGQ == 0, "plan: radix-8 first pass, 4 passes");
    uint4 sa[P0::NQ], sb4[P0::NQ];
    TwPre<P0::NQ> tw0;
    if constexpr (stage_fwd<LOGH>()) {
        stage_wait<LOGH>(buf);
        const uint4 *stage = reinterpret_cast<const uint4 *>(stage_of<LOGH>(buf));
        sa[0] = stage[tid];
        sb4[0] = stage[kPairThreads + tid];
        tw0 = tw_issue<LOGH, 0, false>(tb.tw, tid);
    } else {
#pragma unroll
        for (int i = 0; i < P0::NQ; ++i) { sa[i] = in.ta[i]; sb4[i] = in.tb[i]; }
        tw0 = in.tw;
    }
    const auto t1 = tw_issue<LOGH, 1, false>(tb.tw, tid);
    // pass 0 gathers its inputs from the column block: packed z[e] = v[2e] + i v[2e+1], v = Makhoul permutation
#pragma unroll
    for (int i = 0; i < P0::NQ; ++i) {
        const int q = tid + i * kPairThreads;
        if (P0::NB < kPairThreads && q >= P0::NB) break;
        const unsigned e[8] = {sa[i].x, sa[i].y, sa[i].z, sa[i].w, sb4[i].x, sb4[i].y, sb4[i].z, sb4[i].w};
        CP v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const unsigned ir = e[m] & 0xFFFFu, ii = e[m] >> 16;
            v[m].re = ir != kNone ? x.get(ir) : make_float2(0.f, 0.f);
            v[m].im = ii != kNone ? x.get(ii) : make_float2(0.f, 0.f);
        }
        radix_dft<3, false>(v);
        radix_twiddle<3>(v, tw0.w1[i], tw0.w2[i], tw0.w4[i]);
        const int sb = pswz<LOGH>(q);
#pragma unroll
        for (int m = 0; m < 8; ++m) st(buf, sb ^ pswz<LOGH>(m << P0::LOGQ), v[m]);
    }
    __syncthreads();
    const auto t2 = tw_issue<LOGH, 2, false>(tb.tw, tid);
    pass_run<LOGH, 1, false>(buf, t1, tid);
    if constexpr (fused_fwd<LOGH>()) {
        // passes 1 and 2 of warp w both work on elements [256 w, 256 w + 256): no CTA barrier between them
        __syncwarp();
        const PostPre pp = post_issue<LOGH>(tb, tid);
        pass_run<LOGH, 2, false>(buf, t2, tid);
        __syncthreads();
        const GatPre g = gather_issue(b.fgat, b.gat_chunks, tid);       // (one pass earlier: -0.2 %)
        forward_last_fused<LOGH>(buf, pp.hp, tid);
        __syncthreads();
        return g;
    }
    __syncthreads();
    pass_run<LOGH, 2, false>(buf, t2, tid);
    __syncthreads();
    const PostPre pp = post_issue<LOGH>(tb, tid);
    pass_run<LOGH, 3, false>(buf, TwPre<PassInfo<LOGH, 3>::NQ>(), tid);
    __syncthreads();
    const GatPre g = gather_issue(b.fgat, b.gat_chunks, tid);
    // real-FFT split + DCT phase, pairs visited in bit-reversed order (consecutive lanes, consecutive slots)
#pragma unroll
    for (int i = 0; i < kPrePF; ++i) {
        const int t = tid + i * kPairThreads;
        if (t > Hh) break;
        if (t == 0) {
            const CP z = ld(buf, 0);
            st(buf, 0, CP{padd(z.re, z.im), pscale(psub(z.re, z.im), 0.70710678118654752440f)});
        } else {
            post_pair(buf, pp.slot[i], pp.hp[i]);
        }
    }
    for (int t = tid + kPrePF * kPairThreads; t <= Hh; t += kPairThreads) post_pair(buf, __ldg(tb.post_slot + t), __ldg(tb.post + t));
    __syncthreads();
    return g;
}

// dst2[i] += sc * Y[id(i)] for the rows of one block (Y: float2 view of the pair buffer) in the scheduled order of
// PairBlock::fgat.  A row changes hands from block to block: the caller's barrier between two blocks orders the updates.
// NS slots at a time: all loads first, then the updates (a thread's rows are distinct; written this way the 2 NS
// shared-memory loads are in flight together -- through one slot at a time the compiler has to keep each load of the
// residual behind the previous store, whose address it cannot tell apart, and the gather runs at load latency).
template <int NS>
__device__ __forceinline__ void gather_slots(float2 *dst2, const float2 *bf2, const unsigned *e, float sc)
{
    float2 yv[NS], zv[NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        const bool on = e[u] != 0xFFFFFFFFu;
        yv[u] = on ? bf2[e[u] >> 16] : make_float2(0.f, 0.f);
        zv[u] = on ? dst2[e[u] & 0xFFFFu] : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < NS; ++u)
        if (e[u] != 0xFFFFFFFFu) dst2[e[u] & 0xFFFFu] = pfmas(yv[u], sc, zv[u]);
}
__device__ __forceinline__ void gather_add(float2 *dst2, const float4 *buf, const GatPre &g, const uint4 *__restrict__ tab,
                                           float sc, int chunks, int tid)
{
    static_assert(kGatPF == 2, "two prefetched chunks form the first group of eight slots");
    const float2 *bf2 = reinterpret_cast<const float2 *>(buf);
    const uint4 none = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    const uint4 r0 = 2 < chunks ? ldg_na(tab + 2 * kPairThreads + tid) : none;     // second group: loaded under the first
    const uint4 r1 = 3 < chunks ? ldg_na(tab + 3 * kPairThreads + tid) : none;
    {
        const unsigned e[8] = {g.c[0].x, g.c[0].y, g.c[0].z, g.c[0].w, g.c[1].x, g.c[1].y, g.c[1].z, g.c[1].w};
        gather_slots<8>(dst2, bf2, e, sc);
    }
    if (2 < chunks) {
        const unsigned e[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        gather_slots<8>(dst2, bf2, e, sc);
    }
    for (int c = 4; c < chunks; ++c) {
        const uint4 q = ldg_na(tab + c * kPairThreads + tid);
        const unsigned e[4] = {q.x, q.y, q.z, q.w};
        gather_slots<4>(dst2, bf2, e, sc);
    }
}

// ------------------------------------------------------------------ transposed
struct TrIn { uint2 src[kPrePF]; float2 hp[kPrePF]; };
template <int LOGH>
__device__ __forceinline__ TrIn transposed_issue(const PairTables &tb, const PairBlock &b, int tid, float4 *buf)
{
    constexpr int Hh = 1 << (LOGH - 1);
    TrIn in;
    if constexpr (stage_inv<LOGH>()) {
        stage_issue<LOGH>(buf, b.tsrc, tid);
        return in;
    }
    if constexpr (fused_inv<LOGH>()) {
        const uint4 a = ldg_na(reinterpret_cast<const uint4 *>(b.tsrc) + tid), c = ldg_na(reinterpret_cast<const uint4 *>(b.tsrc) + kPairThreads + tid);
        in.src[0] = make_uint2(a.x, a.y); in.src[1] = make_uint2(a.z, a.w);
        in.src[2] = make_uint2(c.x, c.y); in.src[3] = make_uint2(c.z, c.w);
        ld4_f2(tb.pre + 4 * tid, in.hp);
        return in;
    }
#pragma unroll
    for (int i = 0; i < kPrePF; ++i) {
        const int k = tid + i * kPairThreads;
        in.src[i] = k <= Hh ? ldg_na(b.tsrc + k) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        in.hp[i] = k <= Hh ? __ldg(tb.pre + k) : make_float2(0.f, 0.f);
    }
    return in;
}

template <int LOGH>
__device__ __forceinline__ void pre_pair(float4 *buf, const float2 *y2, int k, uint2 e, float2 hp)
{
    constexpr int H = 1 << LOGH, Hh = H >> 1;
    const unsigned a0 = e.x & 0xFFFFu, a1 = e.x >> 16, b0 = e.y & 0xFFFFu, b1 = e.y >> 16;
    const P zero = make_float2(0.f, 0.f);
    const CP ca = {a0 != kNone ? y2[a0] : zero, a1 != kNone ? y2[a1] : zero};     // (c[k], c[N-k])
    if (k == 0) {
        const P wH = pscale(ca.im, 0.70710678118654752440f);
        st(buf, 0, CP{padd(ca.re, wH), psub(ca.re, wH)});
        return;
    }
    const CP cb = {b0 != kNone ? y2[b0] : zero, b1 != kNone ? y2[b1] : zero};     // (c[H-k], c[H+k])
    const float2 q = q_from_half_pt(hp);
    // W[k] = conj(pt[k]) (c[k] - i c[N-k]) / 2 = conj(hp * ca); A = W[k] + conj(W[H-k]), B = (W[k] - conj(W[H-k])) conj(q)
    const CP U = cmulw(ca, hp), V = cmulw(cb, half_pt_mirror(hp));
    const CP A = {padd(U.re, V.re), psub(V.im, U.im)};
    const CP D = {psub(U.re, V.re), pneg(padd(U.im, V.im))};
    const CP B = cmulw(D, make_float2(q.x, -q.y));
    st(buf, pswz<LOGH>(k), CP{psub(A.re, B.im), padd(A.im, B.re)});                 // A + iB
    if (k != Hh) st(buf, pswz<LOGH>(H - k), CP{padd(A.re, B.im), psub(B.re, A.im)});   // conj(A) + i conj(B)
}

// the pre-processing of one mirror pair into registers: rows (c[k], c[N-k]) and (c[H-k], c[H+k]) -> Z[k], Z[H-k]
__device__ __forceinline__ void pre_pair_reg(const float2 *y2, uint2 e, float2 hp, CP &zk, CP &zm)
{
    const unsigned a0 = e.x & 0xFFFFu, a1 = e.x >> 16, b0 = e.y & 0xFFFFu, b1 = e.y >> 16;
    const P zero = make_float2(0.f, 0.f);
    const CP ca = {a0 != kNone ? y2[a0] : zero, a1 != kNone ? y2[a1] : zero};
    const CP cb = {b0 != kNone ? y2[b0] : zero, b1 != kNone ? y2[b1] : zero};
    const float2 q = q_from_half_pt(hp);
    const CP U = cmulw(ca, hp), V = cmulw(cb, half_pt_mirror(hp));
    const CP A = {padd(U.re, V.re), psub(V.im, U.im)};
    const CP D = {psub(U.re, V.re), pneg(padd(U.im, V.im))};
    const CP B = cmulw(D, make_float2(q.x, -q.y));
    zk = CP{psub(A.re, B.im), padd(A.im, B.re)};
    zm = CP{padd(A.re, B.im), psub(B.re, A.im)};
}
// First pass of the transposed transform at H = 2^11 (radix 4, stride Q = H/4) with the pre-processing fused:
// thread t > 0 builds the inputs j + m Q of butterfly j = t and jp + m Q of butterfly jp = Q - t from its four
// mirror pairs; thread 0 takes j = 0 and jp = Q/2 (self-mirrored).  `in` holds the pairs in the order of
// PairTables::pre.
template <int LOGH>
__device__ __forceinline__ void transposed_first_fused(float4 *buf, const float2 *__restrict__ tw, const float2 *__restrict__ tb_pre,
                                                       const PairBlock &b, const float2 *y2, const TrIn &in, int tid)
{
    constexpr int Q = 1 << (LOGH - 2);
    const bool t0 = tid == 0;
    const int j = tid, jp = t0 ? Q / 2 : Q - tid;
    // inverse twiddles w^j, w^2j (w = exp(2 pi i / H)); those of Q - j follow exactly: w^(Q-j) = i conj(w^j), w^(2Q-2j) = -conj(w^2j)
    float2 w1j = __ldg(tw + j), w2j = __ldg(tw + 2 * j);
    w1j.y = -w1j.y; w2j.y = -w2j.y;
    const float kR = 0.70710678118654752440f;
    const float2 w1p = t0 ? make_float2(kR, kR) : make_float2(w1j.y, w1j.x);
    const float2 w2p = t0 ? make_float2(0.f, 1.f) : make_float2(-w2j.x, w2j.y);
    // mirror pairs of thread t > 0 (rows k = j, j + Q, Q - j, 2Q - j) give (v0,u3), (v1,u2), (u0,v3), (u1,v2);
    // those of thread 0 (rows k = Q/2, 3Q/2, Q, H/2)                 give (u0,u3), (u1,u2), (v1,v3), (v2,--), v0 from k = 0.
    // Only the first members land elsewhere: select them (no divergent path: the warp of thread 0 would run both).
    uint2 src[4];
    float2 hp[4];
    if constexpr (stage_inv<LOGH>()) {
        stage_wait<LOGH>(buf);
        const uint4 *stage = reinterpret_cast<const uint4 *>(stage_of<LOGH>(buf));
        const uint4 a = stage[tid], c = stage[kPairThreads + tid];
        src[0] = make_uint2(a.x, a.y); src[1] = make_uint2(a.z, a.w); src[2] = make_uint2(c.x, c.y); src[3] = make_uint2(c.z, c.w);
        ld4_f2(tb_pre + 4 * tid, hp);
    } else {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) { src[s4] = in.src[s4]; hp[s4] = in.hp[s4]; }
    }
    CP a0, a1, a2, a3, v[4], u[4];
    pre_pair_reg(y2, src[0], hp[0], a0, u[3]);
    pre_pair_reg(y2, src[1], hp[1], a1, u[2]);
    pre_pair_reg(y2, src[2], hp[2], a2, v[3]);
    pre_pair_reg(y2, src[3], hp[3], a3, v[2]);
    CP z0 = a0;
    if (t0) {
        const uint2 e0 = ldg_na(b.tsrc + 4 * kPairThreads);      // k = 0: (c[0], c[H])
        const unsigned r0 = e0.x & 0xFFFFu, r1 = e0.x >> 16;
        const P zero = make_float2(0.f, 0.f);
        const P c0 = r0 != kNone ? y2[r0] : zero, wH = pscale(r1 != kNone ? y2[r1] : zero, kR);
        z0 = CP{padd(c0, wH), psub(c0, wH)};
    }
    v[0] = z0;
    v[1] = csel(t0, a2, a1);
    u[0] = csel(t0, a0, a2);
    u[1] = csel(t0, a1, a3);
    v[2] = csel(t0, a3, v[2]);
    radix_dft<2, true>(v);
    radix_dft<2, true>(u);
    radix_twiddle<2>(v, w1j, w2j, w2j);
    radix_twiddle<2>(u, w1p, w2p, w2p);
#pragma unroll
    for (int m = 0; m < 4; ++m) st(buf, pswz_t<LOGH>(j + m * Q), v[m]);
#pragma unroll
    for (int m = 0; m < 4; ++m) st(buf, pswz_t<LOGH>(jp + m * Q), u[m]);
}

// buf <- DCT-III of the row block y2 (shared memory, float2 per entry) scattered to its rows; the inverse
// FFT leaves X in bit-reversed slots (tgat knows them).  `in` was issued by transposed_issue() for this
// block.  Overwrites buf from the start; issue_out() is called before the last pass so the caller's loads
// for the gather that follows run under it; ends with a barrier and returns issue_out()'s result.
template <int LOGH, typename IssueOut>
__device__ __forceinline__ auto transposed_run(float4 *buf, const PairTables &tb, const PairBlock &b, const float2 *y2,
                                               const TrIn &in, int tid, IssueOut issue_out) -> decltype(issue_out())
{
    constexpr int Hh = 1 << (LOGH - 1);
    if constexpr (fused_inv<LOGH>()) {
        typedef PassX<LOGH, LOGH - 2, 3> P1;
        typedef PassX<LOGH, LOGH - 5, 3> P2;
        typedef PassX<LOGH, LOGH - 8, 3> P3;
        static_assert(P3::LOGQ == 0 && P1::NQ == 1, "plan 4,8,8,8 with one butterfly per thread");
        transposed_first_fused<LOGH>(buf, tb.tw, tb.pre, b, y2, in, tid);
        const auto t1 = tw_issue_p<P1, true>(tb.tw, tid);
        __syncthreads();
        const auto t2 = tw_issue_p<P2, true>(tb.tw, tid);
        pass_run_p<P1, true, true>(buf, t1, tid);
        sync_warp_pair(tid);                      // pass 1 of warps 2k, 2k+1 covers one 512-element sub-transform
        const auto out = issue_out();             // two passes ahead of their use (one pass: -0.4 %)
        pass_run_p<P2, true, true>(buf, t2, tid);
        __syncwarp();                             // passes 2 and 3 of warp w both work on elements [256 w, 256 w + 256)
        pass_run_p<P3, true, true>(buf, TwPre<P3::NQ>(), tid);
        __syncthreads();
        return out;
    }
    const auto t0 = tw_issue<LOGH, 0, true>(tb.tw, tid);
#pragma unroll
    for (int i = 0; i < kPrePF; ++i) {
        const int k = tid + i * kPairThreads;
        if (k > Hh) break;
        pre_pair<LOGH>(buf, y2, k, in.src[i], in.hp[i]);
    }
    for (int k = tid + kPrePF * kPairThreads; k <= Hh; k += kPairThreads) pre_pair<LOGH>(buf, y2, k, ldg_na(b.tsrc + k), __ldg(tb.pre + k));
    __syncthreads();
    const auto t1 = tw_issue<LOGH, 1, true>(tb.tw, tid);
    pass_run<LOGH, 0, true>(buf, t0, tid);
    __syncthreads();
    const auto t2 = tw_issue<LOGH, 2, true>(tb.tw, tid);
    pass_run<LOGH, 1, true>(buf, t1, tid);
    __syncthreads();
    pass_run<LOGH, 2, true>(buf, t2, tid);
    __syncthreads();
    const auto out = issue_out();
    pass_run<LOGH, 3, true>(buf, TwPre<PassInfo<LOGH, 3>::NQ>(), tid);
    __syncthreads();
    return out;
}

}  // namespace pairfft
}  // namespace lsb_sparc

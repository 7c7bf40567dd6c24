// Device code of the batched flooding BP decoder (sm_100a) + its launchers.
// Included by ldpc_bp_f32.cu / ldpc_bp_f64.cu, which instantiate one precision
// each so the two compile in parallel.
//
// What it computes (reference: ldpc_jossy/src/c_ldpc.c, paths under
// /root/reference): per frame, up to max_it iterations of
//   VN pass  (c_ldpc.c:171-178): total = ch[v] + sum of the variable's edge
//            messages in port order; every edge message <- total - message;
//            app[v] = total
//   CN pass  (c_ldpc.c:183-194 -> Lxfb :294-314 -> Lxor :234-251): every check
//            rewrites its dc messages with the extrinsic LLRs from a
//            forward/backward Lxor trellis and reports the LLR of the XOR of all
//            its inputs; the frame stops when every check reports > 0.
// The returned iteration count and app are the reference's (app comes from the
// VN pass of the last executed iteration).
//
// How it is laid out for B200:
//  * QC kernel (bp_qc_kernel): one CTA decodes one frame at a time out of a
//    device-side frame queue (atomic counter -> natural load balance under
//    early stopping).  All edge messages of the frame live in SHARED MEMORY for
//    the whole decode, laid out [protograph edge][z]: the message of protograph
//    entry e = (row r, column c, shift s) that belongs to check r*z+k sits at
//    e*z + k.  Lanes walk consecutive checks (CN pass) or consecutive variables
//    (VN pass: k = j - s mod z), so both passes are shared-memory conflict free
//    with NO index table: addresses follow from the QC structure.  Check rows
//    are visited in degree-sorted order so a warp runs one unrolled trellis
//    (register arrays, compile-time dc).  HBM traffic per frame is the
//    compulsory LLR read + APP/decision write only.
//  * generic kernel (bp_generic_kernel): any (vdeg, cdeg, intrlv) graph, messages
//    in a per-CTA global workspace (L1/L2 resident); serves the legacy
//    single-frame ABI and graphs that do not fit the QC kernel.
#pragma once
#include "ldpc_bp_host.h"

// 1: fp32 sumprod2 takes the stop-rule syndrome from the sign/min rule (same sign
// and zero as the corrected value, saves one Lxor per check)
#ifndef LSB_BP_CHEAP_SYNDROME
#define LSB_BP_CHEAP_SYNDROME 1
#endif
// 1: the fp32 sumprod2 check node runs in the exp domain (check_node_edomain); 0: Lxor form (min + 2 log terms)
#ifndef LSB_BP_EDOMAIN
#define LSB_BP_EDOMAIN 1
#endif
// resident CTAs per SM the fp32 degree<=8 kernel's register budget is tuned for
#ifndef LSB_BP_OCC8
#define LSB_BP_OCC8 5
#endif

namespace lsb_bp {

// ---------------------------------------------------------------- arithmetic
template <typename T>
struct Num;

template <>
struct Num<float> {
    static __device__ __forceinline__ float ex2(float x)
    {
        float y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    // lg2.approx.ftz: __log2f carries a denormal-input rescue (FSETP + 2 predicated ops per call) that the exp-domain
    // node does not need -- its arguments are >= 2^-120
    static __device__ __forceinline__ float lg2(float x)
    {
        float y;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        return y;
    }
    // fp32 sumprod2 messages are kept in the LOG2 DOMAIN: every LLR is scaled by
    // log2(e) when the frame is loaded (kIn) and by ln 2 when APP is stored (kOut).
    // The variable-node rule is linear and the stopping rule looks at signs only,
    // so nothing else changes, and Lxor becomes
    //   sign*min + lg2(1 + 2^-|a+b|) - lg2(1 + 2^-|a-b|)
    // with no multiplies: 4 MUFU (2 EX2 + 2 LG2) and 9 other instructions.
    // (A 3-MUFU form, lg2((1+u)/(1+v)) = 2 log2(e) atanh(w) with one MUFU.RCP and an odd series, measured
    // 1.4-5.5 % slower on B200: its extra FP32 instructions cost more issue slots than the MUFU saves.)
    // Abs. error of the correction ~2e-7; parity bound 1e-4 relative.
    static __device__ __forceinline__ float add_corr(float out, float a, float b)
    {
        const float u = ex2(-fabsf(a + b));
        const float v = ex2(-fabsf(a - b));
        return out + (__log2f(1.0f + u) - __log2f(1.0f + v));
    }
    static __device__ __forceinline__ float with_sign_of_product(float mag, float a, float b)
    {
        unsigned s = (__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u;
        return __uint_as_float(__float_as_uint(mag) | s);
    }
    static __device__ __noinline__ float tanh_half(float x) { return tanhf(0.5f * x); }
    static __device__ __noinline__ float atanh2(float x) { return 2.0f * atanhf(x); }
};

template <>
struct Num<double> {
    // Same expression and evaluation order as the reference (c_ldpc.c:246-247).
    // Not inlined: the fp64 path is the parity mode, and inlining libdevice
    // exp/log into every unrolled trellis costs minutes of compile time.
    static __device__ __noinline__ double add_corr(double out, double a, double b)
    {
        out += log(1 + exp(-fabs(a + b)));
        out -= log(1 + exp(-fabs(a - b)));
        return out;
    }
    static __device__ __forceinline__ double with_sign_of_product(double mag, double a, double b)
    {
        unsigned long long s = (__double_as_longlong(a) ^ __double_as_longlong(b)) &
                               0x8000000000000000ull;
        return __longlong_as_double(__double_as_longlong(mag) | s);
    }
    static __device__ __noinline__ double tanh_half(double x) { return tanh(x / 2.0); }
    static __device__ __noinline__ double atanh2(double x) { return 2.0 * atanh(x); }
};

// Scale applied to LLRs entering / leaving a decode (see Num<float>::add_corr).
template <typename T, int ALGO>
struct Domain {
    static __device__ __forceinline__ T in(T x) { return x; }
    static __device__ __forceinline__ T out(T x) { return x; }
};
template <>
struct Domain<float, LSB_BP_SUMPROD2> {
    static __device__ __forceinline__ float in(float x) { return x * 1.4426950408889634f; }
    static __device__ __forceinline__ float out(float x) { return x * 0.6931471805599453f; }
};

// LLR of the XOR of two bits (c_ldpc.c:234-251).  mag >= 0, so OR-ing the sign
// bit in reproduces "sign(a)sign(b) * min(|a|,|b|)" including signed zeros.
template <typename T, bool CORR>
__device__ __forceinline__ T lxor(T a, T b)
{
    T out = Num<T>::with_sign_of_product(fmin(fabs(a), fabs(b)), a, b);
    if (CORR) out = Num<T>::add_corr(out, a, b);
    return out;
}

// fp32 sumprod2 check node in the EXP DOMAIN (LSB_BP_EDOMAIN).  With E = 2^-|L| (log2-domain LLRs) the
// exact box-plus c_ldpc.c:234-251 evaluates -- sign(a)sign(b)min(|a|,|b|) + log(1+e^-|a+b|) - log(1+e^-|a-b|)
// = 2 atanh(tanh(a/2) tanh(b/2)) -- becomes
//     E(a [+] b) = (Ea + Eb) / (1 + Ea Eb),       sign(a [+] b) = sign(a) sign(b),
// a ratio of two polynomials with positive coefficients: no cancellation anywhere.  The trellis of Lxfb
// (same association order: f[k] = f[k-1] [+] L[k], b[k] = b[k+1] [+] L[k], out[k] = f[k-1] [+] b[k+1])
// is run on fractions N/D that are never divided out:
//     (Na/Da) [+] (Nb/Db) = (Na Db + Nb Da) / (Da Db + Na Nb),      (N/D) [+] E = (N + E D) / (D + E N),
// and an extrinsic magnitude is lg2(D) - lg2(N).  Per check: dc EX2 + 2 dc LG2 = 3 dc MUFU and 8 (dc - 2) FMA-pipe
// operations, against 4 (3 dc - 6) MUFU and ~13 (3 dc - 6) other instructions of the Lxor form above (dc = 7:
// 21 + 40 against 60 + 195).
// Range.  2^-|L| underflows at |L| = 126 (87 in natural units) while slowly converging frames reach LLRs of
// several hundred, so the node works on |L| - s, s = max(min|L| - 32, 0), and adds s back to its outputs: once
// every input exceeds 32 the E^2 terms are below 2^-64 and box-plus commutes with the shift to fp32 accuracy.
// Inputs more than ~120 above the shifted minimum flush to E = 0, which is exact as long as every output still
// sees one input in range; the one case that is not -- a single weak input among inputs that are all > 120
// larger -- shows up as a numerator below 2^-120 and is reported to the caller (NaN, nothing stored), who runs
// the Lxor form on that check.  D stays below 2^dc.
// The stop rule needs the sign of the XOR of all inputs and whether any input is exactly zero (then the
// reference's b[0] is 0 and "aggr <= 0" holds): the returned value carries that sign and min |L|.
constexpr float kEdomainKeep = 32.0f;
constexpr float kEdomainTiny = 7.5231638e-37f;      // 2^-120

template <int DC>
__device__ __forceinline__ float check_node_edomain(float (&L)[DC])
{
    unsigned par = __float_as_uint(L[0]);
    float mn = fabsf(L[0]);
#pragma unroll
    for (int i = 1; i + 1 < DC; i += 2) {
        par ^= __float_as_uint(L[i]) ^ __float_as_uint(L[i + 1]);
        mn = fminf(fminf(mn, fabsf(L[i])), fabsf(L[i + 1]));          // FMNMX3
    }
    if ((DC & 1) == 0) {
        par ^= __float_as_uint(L[DC - 1]);
        mn = fminf(mn, fabsf(L[DC - 1]));
    }
    const float s = fmaxf(mn - kEdomainKeep, 0.0f);
    float E[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) E[i] = Num<float>::ex2(s - fabsf(L[i]));
    float bN[DC], bD[DC];
    bN[DC - 1] = E[DC - 1];
    bD[DC - 1] = 1.0f;
#pragma unroll
    for (int i = DC - 2; i >= 1; --i) {
        bN[i] = fmaf(E[i], bD[i + 1], bN[i + 1]);
        bD[i] = fmaf(E[i], bN[i + 1], bD[i + 1]);
    }
    float oN[DC], oD[DC];
    float fN = E[0], fD = 1.0f;
    oN[0] = bN[DC > 1 ? 1 : 0];
    oD[0] = bD[DC > 1 ? 1 : 0];
#pragma unroll
    for (int i = 1; i < DC - 1; ++i) {
        oN[i] = fmaf(fN, bD[i + 1], bN[i + 1] * fD);
        oD[i] = fmaf(fD, bD[i + 1], fN * bN[i + 1]);
        const float n = fmaf(E[i], fD, fN);
        fD = fmaf(E[i], fN, fD);
        fN = n;
    }
    oN[DC - 1] = fN;
    oD[DC - 1] = fD;
    float lo = oN[0];
#pragma unroll
    for (int i = 1; i + 1 < DC; i += 2) lo = fminf(fminf(lo, oN[i]), oN[i + 1]);
    if ((DC & 1) == 0) lo = fminf(lo, oN[DC - 1]);
    if (!(lo >= kEdomainTiny)) return __int_as_float(0x7fffffff);     // out of range (or NaN input): caller takes the Lxor form
#pragma unroll
    for (int i = 0; i < DC; ++i) {
        const float mag = (Num<float>::lg2(oD[i]) - Num<float>::lg2(oN[i])) + s;
        // |mag| (rounding can leave -ulp when every other input is ~0) with the sign of the other inputs' product
        L[i] = __uint_as_float((__float_as_uint(mag) & 0x7fffffffu) | ((par ^ __float_as_uint(L[i])) & 0x80000000u));
    }
    return __uint_as_float(__float_as_uint(mn) | (par & 0x80000000u));
}

// One check node with its dc messages in registers (c_ldpc.c:294-314).
// Returns the LLR of the XOR of all inputs (b[0] in the reference).  Lxor is
// symmetric in its arguments, so only the association order matters and that
// is the reference's: f[k] = f[k-1] (+) L[k], b[k] = b[k+1] (+) L[k],
// out[k] = f[k-1] (+) b[k+1].
template <typename T, int DC, int ALGO>
__device__ __forceinline__ T check_node(T (&L)[DC], T scale)
{
    if (ALGO == LSB_BP_SUMPROD) {   // c_ldpc.c:76-102
        T t[DC];
        T prod = (T)1;
#pragma unroll
        for (int i = 0; i < DC; ++i) {
            t[i] = Num<T>::tanh_half(L[i]);
            prod *= t[i];
        }
#pragma unroll
        for (int i = 0; i < DC; ++i) L[i] = Num<T>::atanh2(prod / t[i]);
        return Num<T>::atanh2(prod);
    } else {
        constexpr bool CORR = (ALGO == LSB_BP_SUMPROD2);
        constexpr bool F32 = sizeof(T) == 4;
        if constexpr (F32 && CORR && LSB_BP_EDOMAIN && DC >= 2) return check_node_edomain<DC>(L);
        T b[DC];
        b[DC - 1] = L[DC - 1];
#pragma unroll
        for (int i = DC - 2; i >= 1; --i) b[i] = lxor<T, CORR>(b[i + 1], L[i]);
        // b[0] only feeds the stopping rule "<= 0": the uncorrected value has the
        // same sign and the same zeros (|correction| < min(|a|,|b|))
        b[0] = (F32 && LSB_BP_CHEAP_SYNDROME) ? lxor<T, false>(b[1], L[0]) : lxor<T, CORR>(b[1], L[0]);
        T f = L[0];
        L[0] = b[1];
#pragma unroll
        for (int i = 1; i < DC - 1; ++i) {
            T out = lxor<T, CORR>(f, b[i + 1]);
            f = lxor<T, CORR>(f, L[i]);
            L[i] = out;
        }
        L[DC - 1] = f;
        if (ALGO == LSB_BP_MINSUM) {
#pragma unroll
            for (int i = 0; i < DC; ++i) L[i] *= scale;
        }
        return b[0];
    }
}

// Any degree up to kMaxDc: same recursion with local-memory arrays (generic /
// compatibility path only; the QC kernel never calls it).
template <typename T, int ALGO>
__device__ __noinline__ T check_node_dyn(T *m, int stride, int dc, T scale)
{
    T L[kMaxDc], b[kMaxDc];
    for (int i = 0; i < dc; ++i) L[i] = m[i * stride];
    T syn;
    if (ALGO == LSB_BP_SUMPROD) {
        T prod = (T)1;
        for (int i = 0; i < dc; ++i) {
            b[i] = Num<T>::tanh_half(L[i]);
            prod *= b[i];
        }
        for (int i = 0; i < dc; ++i) L[i] = Num<T>::atanh2(prod / b[i]);
        syn = Num<T>::atanh2(prod);
    } else {
        constexpr bool CORR = (ALGO == LSB_BP_SUMPROD2);
        b[dc - 1] = L[dc - 1];
        for (int i = dc - 2; i >= 0; --i) b[i] = lxor<T, CORR>(b[i + 1], L[i]);
        T f = L[0];
        L[0] = b[1];
        for (int i = 1; i < dc - 1; ++i) {
            T out = lxor<T, CORR>(f, b[i + 1]);
            f = lxor<T, CORR>(f, L[i]);
            L[i] = out;
        }
        L[dc - 1] = f;
        if (ALGO == LSB_BP_MINSUM)
            for (int i = 0; i < dc; ++i) L[i] *= scale;
        syn = b[0];
    }
    for (int i = 0; i < dc; ++i) m[i * stride] = L[i];
    return syn;
}

// Loads the dc messages of a check (stride apart), runs the trellis, stores the extrinsics.  fp32 sumprod2 in the
// exp domain: a NaN return means "out of range, nothing stored" (see check_node_edomain).
template <typename T, int DC, int ALGO>
__device__ __forceinline__ T check_node_strided(T *m, int stride, T scale)
{
    T L[DC];
#pragma unroll
    for (int i = 0; i < DC; ++i) L[i] = m[i * stride];
    T syn = check_node<T, DC, ALGO>(L, scale);
    if constexpr (sizeof(T) == 4 && ALGO == LSB_BP_SUMPROD2 && LSB_BP_EDOMAIN) {
        if (syn != syn) return syn;
    }
#pragma unroll
    for (int i = 0; i < DC; ++i) m[i * stride] = L[i];
    return syn;
}

// Unrolled trellis for dc in [2, MAXDC]; MAXDC is the kernel's degree bucket.
template <typename T, int ALGO, int MAXDC>
__device__ __forceinline__ T check_node_bucket(T *m, int stride, int dc, T scale)
{
    switch (dc) {
#define LSB_CASE(D)                                                              \
    case D:                                                                      \
        if (D <= MAXDC) return check_node_strided<T, (D <= MAXDC ? D : 2), ALGO>(m, stride, scale); \
        break;
        LSB_CASE(2) LSB_CASE(3) LSB_CASE(4) LSB_CASE(5) LSB_CASE(6) LSB_CASE(7) LSB_CASE(8)
        LSB_CASE(9) LSB_CASE(10) LSB_CASE(11) LSB_CASE(12) LSB_CASE(13) LSB_CASE(14)
        LSB_CASE(15) LSB_CASE(16) LSB_CASE(17) LSB_CASE(18) LSB_CASE(19) LSB_CASE(20)
        LSB_CASE(21) LSB_CASE(22) LSB_CASE(23) LSB_CASE(24)
#undef LSB_CASE
        default:
            break;
    }
    return (T)0;   // unreachable for a graph validated against the bucket
}

// One variable node of a QC column (c_ldpc.c:171-178): total = ch + sum of its
// dv edge messages in port order (= increasing protograph row); every message is
// replaced by total - message.  The message of variable j on protograph entry
// (edge, shift) sits at edge*z + (j - shift mod z).
// entry[t] = (32-bit SHARED-WINDOW address of message edge*z - shift, shift); j_b = j * sizeof(T), z_b = z * sizeof(T):
// the address of a message is three integer instructions (compare, select, 3-input add) -- no scaling, no base
// pointer -- and the access is an explicit ld/st.shared on it.
__device__ __forceinline__ float lds(unsigned a, float)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds(unsigned a, double)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// x + j_b, plus z_b when j < shift: compare, add, predicated add
__device__ __forceinline__ unsigned wrap_addr(int x, int shift, int j, int j_b, int z_b)
{
    unsigned a;
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.lt.s32 p, %1, %2;\n\t"
        "add.s32 %0, %3, %4;\n\t"
        "@p add.s32 %0, %0, %5;\n\t"
        "}"
        : "=r"(a)
        : "r"(j), "r"(shift), "r"(x), "r"(j_b), "r"(z_b));
    return a;
}

template <typename T, int DV>
__device__ __forceinline__ T variable_node(const int2 *entry, int j, int j_b, int z_b, T total)
{
    unsigned a[DV];
    T m[DV];
#pragma unroll
    for (int i = 0; i < DV; ++i) {
        const int2 e = entry[i];
        a[i] = wrap_addr(e.x, e.y, j, j_b, z_b);
        m[i] = lds(a[i], T());
    }
#pragma unroll
    for (int i = 0; i < DV; ++i) total += m[i];
#pragma unroll
    for (int i = 0; i < DV; ++i) sts(a[i], total - m[i]);
    return total;
}

template <typename T>
__device__ __forceinline__ T variable_node_any(const int2 *entry, int dv, int j, int z, T total)
{
    int j_b = j * (int)sizeof(T);
    asm volatile("" : "+r"(j_b));          // keep j_b in a register: folded back into j * 4 + x it costs an IMAD per edge
    const int z_b = z * (int)sizeof(T);
    switch (dv) {
#define LSB_CASE(D) \
    case D:         \
        return variable_node<T, D>(entry, j, j_b, z_b, total);
        LSB_CASE(1) LSB_CASE(2) LSB_CASE(3) LSB_CASE(4) LSB_CASE(5) LSB_CASE(6) LSB_CASE(7) LSB_CASE(8)
        LSB_CASE(9) LSB_CASE(10) LSB_CASE(11) LSB_CASE(12)
#undef LSB_CASE
        default:
            break;
    }
    T sum = total;
    for (int t = 0; t < dv; ++t) {
        const int2 e = entry[t];
        sum += lds(wrap_addr(e.x, e.y, j, j_b, z_b), T());
    }
    for (int t = 0; t < dv; ++t) {
        const int2 e = entry[t];
        const unsigned a = wrap_addr(e.x, e.y, j, j_b, z_b);
        sts(a, sum - lds(a, T()));
    }
    return sum;
}

// ------------------------------------------------------------------ Monte-Carlo front end
// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) -> 4 x 32 random bits.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// two standard normals from two 32-bit words (Box-Muller; u1 in (0, 1], |n| <= 6.66)
__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b)
{
    const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;        // (a + 1) / 2^32; a = 2^32 - 1 rounds to 1.0
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf((float)b * 1.4629180792671596e-9f, &sn, &cs);               // 2 pi b / 2^32
    return make_float2(r * cs, r * sn);
}
// the 4 noise samples of code bits 4 q .. 4 q + 3 of the frame with global index gf
__device__ __forceinline__ float4 mc_noise4(unsigned long long seed, unsigned long long gf, unsigned q)
{
    const uint4 r = philox4x32_10(make_uint4(q, 0u, (unsigned)gf, (unsigned)(gf >> 32)),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------ QC kernel
template <int MAXDC, int ELEM>
struct QcOccupancy {   // resident CTAs per SM the register budget is tuned for
    static constexpr int value = ELEM == 8 ? 2 : (MAXDC <= 8 ? LSB_BP_OCC8 : (MAXDC <= 16 ? 4 : 3));
};

// hard decisions of one frame as packed bits: one warp ballot per 32 decisions (uniform trip count).  Not inlined: the
// epilogue of the BITS variant then leaves the register allocation of the decode loop as it is without it.
template <typename T>
__device__ __noinline__ void pack_decisions(const T *ap, uint32_t *o, int nv, int tid)
{
    for (int v0 = 0; v0 < nv; v0 += kThreads) {
        const int v = v0 + tid;
        const unsigned m = __ballot_sync(0xffffffffu, v < nv && ap[v] < (T)0);
        if ((tid & 31) == 0 && v < nv) o[v >> 5] = m;
    }
}

// BITS: the epilogue also writes the hard decisions as packed bits (McParams::hard_bits).  A template parameter --
// as a run-time branch the extra epilogue changed the register allocation of the main loop (19.25 -> 19.43 ms).
template <typename T, int ALGO, int MAXDC, bool BITS = false>
__global__ void __launch_bounds__(kThreads, QcOccupancy<MAXDC, sizeof(T)>::value)
bp_qc_kernel(const QcTables *__restrict__ g_tabs, const T *__restrict__ llr, T *__restrict__ app,
             uint8_t *__restrict__ hard, int32_t *__restrict__ iters, long long n_frames,
             int max_it, T scale, unsigned long long *__restrict__ queue, McParams mc, int z)
{
    // z arrives as a kernel parameter (not only in the tables) so that the compiler keeps it, and every multiple of it
    // in the strided check-node addressing, in uniform registers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ QcTables tabs;
    __shared__ long long s_frame;
    __shared__ int s_err[kThreads / 32];
    constexpr bool kEdomain = sizeof(T) == 4 && ALGO == LSB_BP_SUMPROD2 && LSB_BP_EDOMAIN;
    constexpr int kExcCap = 128;                       // checks deferred to the Lxor form per iteration (overflow: inline)
    __shared__ int s_exc_n;
    __shared__ unsigned short s_exc[kExcCap];
    if (threadIdx.x == 0) s_exc_n = 0;
    unsigned long long tot_bit = 0, tot_frame = 0, tot_n = 0, tot_it = 0;     // thread 0: this CTA's Monte-Carlo counters

    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(sizeof(QcTables) / sizeof(int)); i += kThreads)
        reinterpret_cast<int *>(&tabs)[i] = reinterpret_cast<const int *>(g_tabs)[i];
    __syncthreads();
    {   // column entries -> shared-window byte addresses of the messages (see variable_node)
        const unsigned msg_base = (unsigned)__cvta_generic_to_shared(smem_raw);
        for (int i = tid; i < tabs.col_ptr[tabs.cols]; i += kThreads)
            tabs.col_entry[i].x = (int)(msg_base + (unsigned)tabs.col_entry[i].x * (unsigned)sizeof(T));
    }
    __syncthreads();

    const int nv = tabs.nv, nc = tabs.nc;
    const int n_msg = tabs.n_edges * z;
    const unsigned magic = tabs.magic;
    T *msg = reinterpret_cast<T *>(smem_raw);   // [n_edges][z]
    T *ch = msg + n_msg;                        // [nv]
    T *ap = ch + nv;                            // [nv]

    for (;;) {
        if (tid == 0) s_frame = (long long)atomicAdd(queue, 1ull);
        __syncthreads();
        const long long f = s_frame;
        if (f >= n_frames) break;

        const uint8_t *frame_cw = nullptr;
        if (mc.cw) {
            // generate the frame instead of reading it: LLR = 2/sigma^2 ((1 - 2 x) + sigma n)  (ldpc_awgn.py:45-56)
            const unsigned long long gf = mc.frame0 + (unsigned long long)f;
            frame_cw = mc.cw + (long long)(gf % (unsigned long long)mc.n_cw) * nv;
            for (int q = tid; 4 * q < nv; q += kThreads) {
                const float4 nz = mc_noise4(mc.seed, gf, (unsigned)q);
                const float nn[4] = {nz.x, nz.y, nz.z, nz.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = 4 * q + u;
                    if (v < nv) {
                        const float l = mc.llr_scale * ((1.0f - 2.0f * (float)frame_cw[v]) + mc.sigma * nn[u]);
                        if (mc.llr_out) mc.llr_out[f * nv + v] = l;
                        ch[v] = Domain<T, ALGO>::in((T)l);
                        ap[v] = (T)0;
                    }
                }
            }
        } else {
            const T *frame_llr = llr + f * nv;
            for (int v = tid; v < nv; v += kThreads) {
                ch[v] = Domain<T, ALGO>::in(frame_llr[v]);
                ap[v] = (T)0;
            }
        }
        for (int i = tid; i < n_msg; i += kThreads) msg[i] = (T)0;
        __syncthreads();

        int it = 0;
        for (; it < max_it; ++it) {
            // ---- variable nodes: lane = consecutive variables of a column
            for (int v = tid; v < nv; v += kThreads) {
                const int c = (int)__umulhi((unsigned)v, magic);
                const int j = v - c * z;
                const int t0 = tabs.col_ptr[c];
                const T total = variable_node_any<T>(tabs.col_entry + t0, tabs.col_ptr[c + 1] - t0, j, z, ch[v]);
                ap[v] = total;
            }
            __syncthreads();
            // ---- check nodes: lane = consecutive checks of a (degree-sorted) row
            int unsatisfied = 0, deferred = 0;
            for (int p = tid; p < nc; p += kThreads) {
                const int ri = (int)__umulhi((unsigned)p, magic);
                const int k = p - ri * z;
                const int r = tabs.row_order[ri];
                const int e0 = tabs.row_ptr[r];
                const int dc = tabs.row_ptr[r + 1] - e0;
                T syn = check_node_bucket<T, ALGO, MAXDC>(msg + e0 * z + k, z, dc, scale);
                if (kEdomain && syn != syn) {
                    // out of the exp-domain node's range (rare: see check_node_edomain): queue the check for the
                    // Lxor form, run by consecutive threads after the pass instead of divergently inside it
                    const int slot = atomicAdd(&s_exc_n, 1);
                    if (slot < kExcCap) {
                        s_exc[slot] = (unsigned short)p;
                        deferred = 1;
                        continue;
                    }
                    syn = check_node_dyn<T, ALGO>(msg + e0 * z + k, z, dc, scale);
                }
                unsatisfied |= (syn <= (T)0);
            }
            int any = __syncthreads_or(unsatisfied | deferred);
            if (kEdomain) {
                const int n_exc = min(s_exc_n, kExcCap);          // CTA-uniform: every push precedes the barrier
                if (n_exc > 0) {
                    for (int q = tid; q < n_exc; q += kThreads) {
                        const int p = s_exc[q];
                        const int ri = (int)__umulhi((unsigned)p, magic);
                        const int k = p - ri * z;
                        const int r = tabs.row_order[ri];
                        const int e0 = tabs.row_ptr[r];
                        const T syn = check_node_dyn<T, ALGO>(msg + e0 * z + k, z, tabs.row_ptr[r + 1] - e0, scale);
                        unsatisfied |= (syn <= (T)0);
                    }
                    any = __syncthreads_or(unsatisfied);
                    if (tid == 0) s_exc_n = 0;                      // next push: after the next VN-pass barrier
                }
            }
            if (!any) break;
        }

        if (tid == 0 && iters) iters[f] = it;
        if (app) {
            T *o = app + f * nv;
            for (int v = tid; v < nv; v += kThreads) o[v] = Domain<T, ALGO>::out(ap[v]);
        }
        if (hard) {
            uint8_t *o = hard + f * nv;
            for (int v = tid; v < nv; v += kThreads) o[v] = ap[v] < (T)0;
        }
        if (BITS) pack_decisions<T>(ap, mc.hard_bits + f * ((nv + 31) >> 5), nv, tid);
        if (mc.cw) {
            // wrong hard decisions over all N code bits (ldpc_awgn.py:97-103), summed in a fixed order
            int wrong = 0;
            for (int v = tid; v < nv; v += kThreads) wrong += (int)(ap[v] < (T)0) != (int)frame_cw[v];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wrong += __shfl_xor_sync(0xffffffffu, wrong, o);
            if ((tid & 31) == 0) s_err[tid >> 5] = wrong;
            __syncthreads();
            if (tid == 0) {
                int total = 0;
                for (int w = 0; w < kThreads / 32; ++w) total += s_err[w];
                if (mc.bit_errors) mc.bit_errors[f] = total;
                tot_bit += (unsigned long long)total;
                tot_frame += total != 0;
                tot_n += 1;
                tot_it += (unsigned long long)it;
            }
        }
    }
    if (mc.cw && mc.counters && tid == 0 && tot_n) {
        atomicAdd(mc.counters + 0, tot_bit);
        atomicAdd(mc.counters + 1, tot_frame);
        atomicAdd(mc.counters + 2, tot_n);
        atomicAdd(mc.counters + 3, tot_it);
    }
}

// ------------------------------------------------------------- generic kernel
// Messages in the reference's own numbering (check-ordered array addressed
// through intrlv on the variable side).  ws: per-CTA [nmsg + nv] scratch.
template <typename T, int ALGO>
__global__ void __launch_bounds__(kThreads)
bp_generic_kernel(const int *__restrict__ voff, const int *__restrict__ coff,
                  const int *__restrict__ intrlv, int nv, int nc, int nmsg,
                  const T *__restrict__ llr, T *__restrict__ app, uint8_t *__restrict__ hard,
                  int32_t *__restrict__ iters, long long n_frames, int max_it, T scale,
                  T *__restrict__ ws, unsigned long long *__restrict__ queue)
{
    __shared__ long long s_frame;
    const int tid = threadIdx.x;
    T *msg = ws + (size_t)blockIdx.x * (size_t)(nmsg + nv);
    T *ap = msg + nmsg;

    for (;;) {
        if (tid == 0) s_frame = (long long)atomicAdd(queue, 1ull);
        __syncthreads();
        const long long f = s_frame;
        if (f >= n_frames) break;
        const T *ch = llr + f * nv;
        for (int i = tid; i < nmsg; i += kThreads) msg[i] = (T)0;
        for (int v = tid; v < nv; v += kThreads) ap[v] = (T)0;
        __syncthreads();

        int it = 0;
        for (; it < max_it; ++it) {
            for (int v = tid; v < nv; v += kThreads) {
                const int e0 = voff[v], e1 = voff[v + 1];
                T total = Domain<T, ALGO>::in(ch[v]);
                for (int e = e0; e < e1; ++e) total += msg[intrlv[e]];
                for (int e = e0; e < e1; ++e) {
                    const int s = intrlv[e];
                    msg[s] = total - msg[s];
                }
                ap[v] = total;
            }
            __syncthreads();
            int unsatisfied = 0;
            for (int c = tid; c < nc; c += kThreads) {
                const int b0 = coff[c];
                T syn = check_node_dyn<T, ALGO>(msg + b0, 1, coff[c + 1] - b0, scale);
                unsatisfied |= (syn <= (T)0);
            }
            if (!__syncthreads_or(unsatisfied)) break;
        }

        if (tid == 0) iters[f] = it;
        if (app) {
            T *o = app + f * nv;
            for (int v = tid; v < nv; v += kThreads) o[v] = Domain<T, ALGO>::out(ap[v]);
        }
        if (hard) {
            uint8_t *o = hard + f * nv;
            for (int v = tid; v < nv; v += kThreads) o[v] = ap[v] < (T)0;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ launchers
template <typename T, int ALGO, int MAXDC>
int launch_qc_bucket(const QcLaunch<T> &a)
{
    constexpr bool kBitsVariant = sizeof(T) == 4 && ALGO == LSB_BP_SUMPROD2;     // the throughput mode only
    if (a.mc.hard_bits && !kBitsVariant) {
        lsb::set_error("packed decisions are built for the fp32 sumprod2 kernel only");
        return LSB_ERR_INVALID;
    }
    auto kern = (kBitsVariant && a.mc.hard_bits) ? bp_qc_kernel<T, ALGO, MAXDC, kBitsVariant> : bp_qc_kernel<T, ALGO, MAXDC, false>;
    const size_t smem = a.smem;
    LSB_CUDA(lsb::ensure_dynamic_smem(kern, smem));       // per (kernel, device), raise-only: other codes share kern
    if (*a.cached_grid == 0) {
        int per_sm = 0;
        LSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        if (per_sm < 1) per_sm = 1;
        *a.cached_grid = per_sm * lsb::sm_count();
    }
    long long grid = *a.cached_grid;
    if (grid > a.n_frames) grid = a.n_frames;
    kern<<<(unsigned)grid, kThreads, smem, a.stream>>>(a.tabs, a.llr, a.app, a.hard, a.iters, a.n_frames,
                                                      a.max_it, a.scale, a.queue, a.mc, a.z);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

template <typename T, int ALGO>
int launch_qc_algo(const QcLaunch<T> &a)
{
    if (a.max_dc <= 8) return launch_qc_bucket<T, ALGO, 8>(a);
    if (a.max_dc <= 16) return launch_qc_bucket<T, ALGO, 16>(a);
    return launch_qc_bucket<T, ALGO, 24>(a);
}

template <typename T>
int launch_qc(const QcLaunch<T> &a)
{
    switch (a.algo) {
        case LSB_BP_SUMPROD2: return launch_qc_algo<T, LSB_BP_SUMPROD2>(a);
        case LSB_BP_SUMPROD: return launch_qc_algo<T, LSB_BP_SUMPROD>(a);
        default: return launch_qc_algo<T, LSB_BP_MINSUM>(a);
    }
}

template <typename T, int ALGO>
int launch_generic_algo(const GenericLaunch<T> &a)
{
    auto kern = bp_generic_kernel<T, ALGO>;
    kern<<<(unsigned)a.grid, kThreads, 0, a.stream>>>(a.voff, a.coff, a.intrlv, a.nv, a.nc, a.nmsg,
                                                     a.llr, a.app, a.hard, a.iters, a.n_frames,
                                                     a.max_it, a.scale, a.ws, a.queue);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

template <typename T>
int launch_generic(const GenericLaunch<T> &a)
{
    switch (a.algo) {
        case LSB_BP_SUMPROD2: return launch_generic_algo<T, LSB_BP_SUMPROD2>(a);
        case LSB_BP_SUMPROD: return launch_generic_algo<T, LSB_BP_SUMPROD>(a);
        default: return launch_generic_algo<T, LSB_BP_MINSUM>(a);
    }
}

template <typename T>
int generic_blocks_per_sm()
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, bp_generic_kernel<T, LSB_BP_SUMPROD2>, kThreads, 0) != cudaSuccess)
        per_sm = 1;
    return per_sm < 1 ? 1 : per_sm;
}

}  // namespace lsb_bp

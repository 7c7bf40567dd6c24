// tcgen05 / TMEM / TMA building blocks shared by the tensor-core products (sm_100a):
// dense_tc.cu (dense Gaussian design, sparc_sophie) and sparc_band_tc.cu (structured DCT design as a block-banded
// matrix).  Everything here has internal linkage (anonymous namespace): each including .cu gets its own copy.
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

// Element type of the split operands (x = hi + lo, three MMAs per product: hi hi + lo hi + hi lo).
//   1: fp16 pairs  -- 11 + 11 mantissa bits (lo enters the subnormal range below |x| ~ 0.1, leaving ~ 20 bits there):
//      products accurate to ~ 5e-7, i.e. fp32-grade; range +-65504 (AMP operands are O(1) .. O(10); the split saturates);
//   0: fp16 pairs  -- 8 + 8 mantissa bits, ~ 8e-6 per product, fp32 range (round 1).
// Same MMA rate (kind::f16) and bytes either way.
#ifndef LSB_TC_HALF
#define LSB_TC_HALF 1
#endif

#include <mutex>

namespace {

#if LSB_TC_HALF
typedef __half tc_t;
constexpr uint32_t kIdescAB = 0u;                                  // A and B formats: F16
constexpr CUtensorMapDataType kTcMapType = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
__device__ __forceinline__ tc_t to_tc(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f)); }
__device__ __forceinline__ float from_tc(tc_t v) { return __half2float(v); }
__device__ __forceinline__ unsigned short tc_bits(tc_t v) { return __half_as_ushort(v); }
#else
typedef __nv_bfloat16 tc_t;
constexpr uint32_t kIdescAB = (1u << 7) | (1u << 10);              // A and B formats: BF16
constexpr CUtensorMapDataType kTcMapType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
__device__ __forceinline__ tc_t to_tc(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float from_tc(tc_t v) { return __bfloat162float(v); }
__device__ __forceinline__ unsigned short tc_bits(tc_t v) { return __bfloat16_as_ushort(v); }
#endif

constexpr int kStages = 3;
constexpr int kTileRows = 128;   // UMMA M: rows of W per tile
constexpr int kTileN = 128;      // UMMA N: frames per tile
constexpr int kBlockK = 64;      // fp16 elements per 128-byte swizzle row
constexpr int kTileBytes = 128 * 128;                 // one operand tile: 128 rows x 128 B
constexpr int kStageBytes = 4 * kTileBytes;           // W_hi, W_lo, X_hi, X_lo
constexpr int kSmemBytes = kStages * kStageBytes + 1024;   // + slack for 1024 B alignment
constexpr int kThreadsTc = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (lane = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K blocks accumulated in TMEM before the epilogue takes the accumulator over into registers.  The tensor core
// rounds its fp32 accumulator toward zero at every MMA, a bias that grows linearly with the K length (measured:
// ~2e-4 relative after the 6144 MMAs of a K = 32768 item, which the AMP transient amplified to 7e-4 in tau);
// chunks of 32 K blocks (384 MMAs) summed in registers with round-to-nearest keep it near 1e-5.
constexpr int kChunkKb = 8;
// the dense-design kernels: their products feed a better-conditioned recursion (no decoding wave) and their denoiser
// epilogue needs the slack; 32 K blocks (384 MMAs, ~1e-5) measured +4 % at config 5 against 8
constexpr int kChunkKbDense = 32;
// TMEM accumulators in the ring (4 x 128 fp32 columns = all 512 columns; one CTA per SM): the MMA warp may run three
// chunks ahead of the epilogue, which needs that slack at item boundaries (stores / the denoiser of a finished tile)
constexpr int kAccs = 4;

// Persistent: CTA b works on items b, b + gridDim.x, ...; an item is one 128 x 128 output
// tile restricted to one of `splits` contiguous K ranges (partial s goes to C + s * split_stride;
// the caller sums the partials in a fixed order, so results do not depend on scheduling).
// Two TMEM accumulators: the epilogue of item j overlaps the MMAs of item j + 1.
__global__ void __launch_bounds__(kThreadsTc, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
               const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
               float *__restrict__ C, int rows_total, int F, int K, int n_row_tiles, int n_f_tiles, int splits,
               long long split_stride, float out_scale)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full_bar[kAccs], acc_empty_bar[kAccs];
    __shared__ uint32_t tmem_base_slot;

    uint8_t *tiles = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + kBlockK - 1) / kBlockK;
    const int n_tiles = n_row_tiles * n_f_tiles;
    const int n_items = n_tiles * splits;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < kAccs; ++a) {
            mbar_init(&acc_full_bar[a], 1);
            mbar_init(&acc_empty_bar[a], 4);     // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "n"(kAccs * kTileN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    // item -> (W row tile, frame tile, K range).  Frame tiles vary fastest: the CTAs running
    // at the same time share a few W panels (one DRAM read each, the rest L2 hits) and the
    // frame operand stays L2-resident.  With W row tiles fastest, A^T z (W = 403 MB of fp16
    // hi/lo, 8 frame tiles) re-read W from DRAM once per frame tile: 578 us vs 467 us.
#define LSB_TC_DECODE_ITEM(item)                                               \
    const int split = (item) / n_tiles, tile = (item) - split * n_tiles;      \
    const int row0 = (tile / n_f_tiles) * kTileRows;                           \
    const int f0 = (tile % n_f_tiles) * kTileN;                                \
    const int kb0 = (int)(((long long)split * num_kb) / splits);               \
    const int kb1 = (int)(((long long)(split + 1) * num_kb) / splits);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                LSB_TC_DECODE_ITEM(item)
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t phase = (it / kStages) & 1;
                    mbar_wait(&empty_bar[s], phase ^ 1);
                    uint8_t *st = tiles + s * kStageBytes;
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    tma_load_2d(st, &mapW_hi, &full_bar[s], kb * kBlockK, row0);
                    tma_load_2d(st + kTileBytes, &mapW_lo, &full_bar[s], kb * kBlockK, row0);
                    tma_load_2d(st + 2 * kTileBytes, &mapX_hi, &full_bar[s], kb * kBlockK, f0);
                    tma_load_2d(st + 3 * kTileBytes, &mapX_lo, &full_bar[s], kb * kBlockK, f0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | kIdescAB | ((uint32_t)(kTileN >> 3) << 17) |
                                   ((uint32_t)(kTileRows >> 4) << 24);
            uint32_t it = 0, j = 0;                          // j counts K chunks (one TMEM accumulator each)
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                LSB_TC_DECODE_ITEM(item)
                (void)row0;
                (void)f0;
                for (int kc = kb0; kc < kb1; kc += kChunkKbDense, ++j) {
                    const int kc1 = kc + kChunkKbDense < kb1 ? kc + kChunkKbDense : kb1;
                    const uint32_t a = j % kAccs;
                    mbar_wait(&acc_empty_bar[a], ((j / kAccs) & 1) ^ 1);     // epilogue has drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + a * kTileN;
                    for (int kb = kc; kb < kc1; ++kb, ++it) {
                        const int s = it % kStages;
                        const uint32_t phase = (it / kStages) & 1;
                        mbar_wait(&full_bar[s], phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t base = smem_u32(tiles + s * kStageBytes);
                        const uint64_t dW_hi = umma_desc(base), dW_lo = umma_desc(base + kTileBytes);
                        const uint64_t dX_hi = umma_desc(base + 2 * kTileBytes), dX_lo = umma_desc(base + 3 * kTileBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);   // 32 bytes per K step inside the swizzle row
                            umma_f16(tmem_d, dW_hi + adv, dX_hi + adv, idesc, (kb != kc) || (k != 0));
                            umma_f16(tmem_d, dW_lo + adv, dX_hi + adv, idesc, 1);
                            umma_f16(tmem_d, dW_hi + adv, dX_lo + adv, idesc, 1);
                        }
                        umma_commit(&empty_bar[s]);      // stage reusable once these MMAs have read it
                    }
                    umma_commit(&acc_full_bar[a]);       // chunk complete
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                   // TMEM lane quarter this warp may read
        uint32_t j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            LSB_TC_DECODE_ITEM(item)
            float *Cs = C + (size_t)split * (size_t)split_stride;
            float acc[kTileN];                    // this thread's row of the tile, summed over the K chunks (RN)
            for (int kc = kb0; kc < kb1; kc += kChunkKbDense, ++j) {
                const uint32_t a = j % kAccs;
                mbar_wait(&acc_full_bar[a], (j / kAccs) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < kTileN; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + a * kTileN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (kc == kb0) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] = v[jj];
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] += v[jj];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty_bar[a]);
            }
            const int r = row0 + q * 32 + lane;
            if (r < rows_total && kb1 > kb0) {
#pragma unroll
                for (int jj = 0; jj < kTileN; ++jj) {
                    const int f = f0 + jj;
                    if (f < F) Cs[(size_t)f * rows_total + r] = acc[jj] * out_scale;
                }
            }
        }
    }
#undef LSB_TC_DECODE_ITEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccs * kTileN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------
// Item-list variant of the kernel above: the host cuts a product (or several block products that share the
// two tensors) into work items -- one 128 x 128 output tile over one contiguous K range of any length -- and
// deals them to the CTAs by longest-processing-time-first, so uneven K ranges (band edges of a spatially
// coupled design) still fill the SMs without K splits and partial sums.  Coordinates of an item are free:
// the W tile may start at any row / K element of the W tensor, the frame operand at any K element.
struct TcItem {
    int w_row0;          // first row of the W tile in the W tensor (the tile is always 128 rows of the tensor)
    int w_k0;            // first K element in the W tensor
    int x_k0;            // first K element in the frame tensor
    int f0;              // first frame
    int kb_count;        // number of 64-element K blocks (W is zero-padded to whole blocks)
    int rows_valid;      // rows of the tile that exist in the output (<= 128)
    long long out_off;   // C[out_off + f * ldc + r_local]
};

__global__ void __launch_bounds__(kThreadsTc, 1)
tc_items_kernel(const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
                float *__restrict__ C, long long ldc, int F, const TcItem *__restrict__ items,
                const int *__restrict__ cta_first)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full_bar[kAccs], acc_empty_bar[kAccs];
    __shared__ uint32_t tmem_base_slot;

    uint8_t *tiles = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = cta_first[blockIdx.x], i1 = cta_first[blockIdx.x + 1];

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < kAccs; ++a) {
            mbar_init(&acc_full_bar[a], 1);
            mbar_init(&acc_empty_bar[a], 4);     // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "n"(kAccs * kTileN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int i = i0; i < i1; ++i) {
                const TcItem im = items[i];
                for (int kb = 0; kb < im.kb_count; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t phase = (it / kStages) & 1;
                    mbar_wait(&empty_bar[s], phase ^ 1);
                    uint8_t *st = tiles + s * kStageBytes;
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    tma_load_2d(st, &mapW_hi, &full_bar[s], im.w_k0 + kb * kBlockK, im.w_row0);
                    tma_load_2d(st + kTileBytes, &mapW_lo, &full_bar[s], im.w_k0 + kb * kBlockK, im.w_row0);
                    tma_load_2d(st + 2 * kTileBytes, &mapX_hi, &full_bar[s], im.x_k0 + kb * kBlockK, im.f0);
                    tma_load_2d(st + 3 * kTileBytes, &mapX_lo, &full_bar[s], im.x_k0 + kb * kBlockK, im.f0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | kIdescAB | ((uint32_t)(kTileN >> 3) << 17) |
                                   ((uint32_t)(kTileRows >> 4) << 24);
            uint32_t it = 0, j = 0;                          // j counts K chunks (one TMEM accumulator each)
            for (int i = i0; i < i1; ++i) {
                const int kb_count = items[i].kb_count;
                for (int kb0 = 0; kb0 < kb_count; kb0 += kChunkKb, ++j) {
                    const int kb1 = kb0 + kChunkKb < kb_count ? kb0 + kChunkKb : kb_count;
                    const uint32_t a = j % kAccs;
                    mbar_wait(&acc_empty_bar[a], ((j / kAccs) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + a * kTileN;
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int s = it % kStages;
                        const uint32_t phase = (it / kStages) & 1;
                        mbar_wait(&full_bar[s], phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t base = smem_u32(tiles + s * kStageBytes);
                        const uint64_t dW_hi = umma_desc(base), dW_lo = umma_desc(base + kTileBytes);
                        const uint64_t dX_hi = umma_desc(base + 2 * kTileBytes), dX_lo = umma_desc(base + 3 * kTileBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                            umma_f16(tmem_d, dW_hi + adv, dX_hi + adv, idesc, (kb != kb0) || (k != 0));
                            umma_f16(tmem_d, dW_lo + adv, dX_hi + adv, idesc, 1);
                            umma_f16(tmem_d, dW_hi + adv, dX_lo + adv, idesc, 1);
                        }
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&acc_full_bar[a]);
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t j = 0;
        for (int i = i0; i < i1; ++i) {
            const TcItem im = items[i];
            float acc[kTileN];                       // this thread's row of the output tile, summed over the K chunks (RN)
            for (int kb0 = 0; kb0 < im.kb_count; kb0 += kChunkKb, ++j) {
                const uint32_t a = j % kAccs;
                mbar_wait(&acc_full_bar[a], (j / kAccs) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < kTileN; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + a * kTileN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (kb0 == 0) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] = v[jj];
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] += v[jj];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty_bar[a]);
            }
            const int r = q * 32 + lane;
            if (r < im.rows_valid) {
                float *Cr = C + im.out_off + r;
#pragma unroll
                for (int jj = 0; jj < kTileN; ++jj) {
                    const int f = im.f0 + jj;
                    if (f < F) Cr[(long long)f * ldc] = acc[jj];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccs * kTileN) : "memory");
    }
}

__device__ __forceinline__ uint32_t pack_pair(tc_t a, tc_t b)
{
    return (uint32_t)tc_bits(a) | ((uint32_t)tc_bits(b) << 16);
}

// ---------------------------------------------------------------------------------------------------
// A^T z product with the AMP denoiser as its epilogue (dense design, sparc_new.py:907-908 + 1040-1066):
//     s = beta + A^T z,   beta <- sqrt_nPl * softmax_section(sqrt_nPl * s / tau^2)
// Same pipeline as tc_gemm_kernel with the operand roles swapped: the frame tile is the MMA's M operand, so a
// TMEM lane is a FRAME and the 128 accumulator columns are 128 consecutive entries of its beta -- whole sections
// (M <= 128 divides 128), each owned by one thread.  The epilogue thread reads its accumulator row, adds the old
// beta, runs the section softmax in registers and writes the new beta as fp32 and as the fp16 hi / lo pair that
// the next A beta product consumes, so the separate denoiser, the fp32 -> fp16 split and the [F, ML] round trip
// of A^T z are gone.  |beta|^2 goes out as one fp32 partial per (row tile, frame); the residual kernel adds the
// partials of a frame in index order.  No K split (the epilogue needs the whole sum).
template <int M>
__global__ void __launch_bounds__(kThreadsTc, 1)
tc_gemm_denoise_kernel(const __grid_constant__ CUtensorMap mapW_hi, const __grid_constant__ CUtensorMap mapW_lo,
                       const __grid_constant__ CUtensorMap mapX_hi, const __grid_constant__ CUtensorMap mapX_lo,
                       float *beta, int has_beta, tc_t *__restrict__ beta_hi, tc_t *__restrict__ beta_lo,
                       float *__restrict__ s_out, const double *__restrict__ tau_sqr, float *__restrict__ bsq_part,
                       float sqrt_nPl, int rows_total, int F, int K, int n_row_tiles, int n_f_tiles, float out_scale,
                       int32_t *__restrict__ idx_out)
{
    constexpr int SEC = M < 32 ? 32 : M;          // accumulator columns handled at a time (whole sections)
    static_assert(M >= 2 && M <= 128 && (M & (M - 1)) == 0, "section size must be a power of two in [2, 128]");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_full_bar[kAccs], acc_empty_bar[kAccs];
    __shared__ uint32_t tmem_base_slot;

    uint8_t *tiles = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + kBlockK - 1) / kBlockK;
    const int n_items = n_row_tiles * n_f_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < kAccs; ++a) {
            mbar_init(&acc_full_bar[a], 1);
            mbar_init(&acc_empty_bar[a], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "n"(kAccs * kTileN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    // frame tiles fastest (see tc_gemm_kernel)
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int row0 = (item / n_f_tiles) * kTileRows, f0 = (item % n_f_tiles) * kTileN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t phase = (it / kStages) & 1;
                    mbar_wait(&empty_bar[s], phase ^ 1);
                    uint8_t *st = tiles + s * kStageBytes;
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    tma_load_2d(st, &mapW_hi, &full_bar[s], kb * kBlockK, row0);
                    tma_load_2d(st + kTileBytes, &mapW_lo, &full_bar[s], kb * kBlockK, row0);
                    tma_load_2d(st + 2 * kTileBytes, &mapX_hi, &full_bar[s], kb * kBlockK, f0);
                    tma_load_2d(st + 3 * kTileBytes, &mapX_lo, &full_bar[s], kb * kBlockK, f0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | kIdescAB | ((uint32_t)(kTileRows >> 3) << 17) |
                                   ((uint32_t)(kTileN >> 4) << 24);
            uint32_t it = 0, j = 0;                          // j counts K chunks (one TMEM accumulator each)
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int kc = 0; kc < num_kb; kc += kChunkKbDense, ++j) {
                    const int kc1 = kc + kChunkKbDense < num_kb ? kc + kChunkKbDense : num_kb;
                    const uint32_t a = j % kAccs;
                    mbar_wait(&acc_empty_bar[a], ((j / kAccs) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + a * kTileRows;
                    for (int kb = kc; kb < kc1; ++kb, ++it) {
                        const int s = it % kStages;
                        const uint32_t phase = (it / kStages) & 1;
                        mbar_wait(&full_bar[s], phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t base = smem_u32(tiles + s * kStageBytes);
                        const uint64_t dW_hi = umma_desc(base), dW_lo = umma_desc(base + kTileBytes);
                        const uint64_t dX_hi = umma_desc(base + 2 * kTileBytes), dX_lo = umma_desc(base + 3 * kTileBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                            umma_f16(tmem_d, dX_hi + adv, dW_hi + adv, idesc, (kb != kc) || (k != 0));     // lanes = frames
                            umma_f16(tmem_d, dX_hi + adv, dW_lo + adv, idesc, 1);
                            umma_f16(tmem_d, dX_lo + adv, dW_hi + adv, idesc, 1);
                        }
                        umma_commit(&empty_bar[s]);
                    }
                    umma_commit(&acc_full_bar[a]);
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t j = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int row_tile = item / n_f_tiles;
            const int row0 = row_tile * kTileRows, f0 = (item % n_f_tiles) * kTileN;
            const int f = f0 + q * 32 + lane;
            const bool live = f < F;
            const float scale = live ? (float)((double)sqrt_nPl / tau_sqr[f]) : 0.0f;
            float sq = 0.0f;
            float acc[kTileRows];                    // this frame's 128 entries of A^T z, summed over the K chunks (RN)
            for (int kc = 0; kc < num_kb; kc += kChunkKbDense, ++j) {
                const uint32_t a = j % kAccs;
                mbar_wait(&acc_full_bar[a], (j / kAccs) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < kTileRows; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + a * kTileRows + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (kc == 0) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] = v[jj];
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) acc[c0 + jj] += v[jj];
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty_bar[a]);
            }
#pragma unroll
            for (int c0 = 0; c0 < kTileRows; c0 += SEC) {
                float *sv = acc + c0;
                if (live && row0 + c0 < rows_total) {
                    const size_t base = (size_t)f * (size_t)rows_total + (size_t)(row0 + c0);
                    if (has_beta) {
#pragma unroll
                        for (int k = 0; k < SEC / 4; ++k) {
                            const float4 b = *reinterpret_cast<const float4 *>(beta + base + 4 * k);
                            sv[4 * k] = fmaf(sv[4 * k], out_scale, b.x);
                            sv[4 * k + 1] = fmaf(sv[4 * k + 1], out_scale, b.y);
                            sv[4 * k + 2] = fmaf(sv[4 * k + 2], out_scale, b.z);
                            sv[4 * k + 3] = fmaf(sv[4 * k + 3], out_scale, b.w);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < SEC; ++k) sv[k] *= out_scale;
                    }
                    if (s_out) {
#pragma unroll
                        for (int k = 0; k < SEC / 4; ++k)
                            *reinterpret_cast<float4 *>(s_out + base + 4 * k) =
                                make_float4(sv[4 * k], sv[4 * k + 1], sv[4 * k + 2], sv[4 * k + 3]);
                    }
#pragma unroll
                    for (int s0 = 0; s0 < SEC; s0 += M) {
                        float mx = sv[s0];
#pragma unroll
                        for (int i = 1; i < M; ++i) mx = fmaxf(mx, sv[s0 + i]);
                        if (idx_out) {                       // MAP decision on s (sparc_new.py:1099-1116): first maximum
                            int arg = 0;
#pragma unroll
                            for (int i = M - 1; i >= 0; --i) arg = sv[s0 + i] == mx ? i : arg;
                            idx_out[(size_t)f * (size_t)(rows_total / M) + (size_t)((row0 + c0 + s0) / M)] = arg;
                        }
                        float sum = 0.0f;
#pragma unroll
                        for (int i = 0; i < M; ++i) {
                            sv[s0 + i] = expf((sv[s0 + i] - mx) * scale);
                            sum += sv[s0 + i];
                        }
                        const float inv = sqrt_nPl / sum;
#pragma unroll
                        for (int i = 0; i < M; ++i) {
                            sv[s0 + i] *= inv;
                            sq = fmaf(sv[s0 + i], sv[s0 + i], sq);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < SEC / 4; ++k)
                        *reinterpret_cast<float4 *>(beta + base + 4 * k) =
                            make_float4(sv[4 * k], sv[4 * k + 1], sv[4 * k + 2], sv[4 * k + 3]);
#pragma unroll
                    for (int k = 0; k < SEC / 8; ++k) {
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float v0 = sv[8 * k + 2 * u], v1 = sv[8 * k + 2 * u + 1];
                            const tc_t h0 = to_tc(v0), h1 = to_tc(v1);
                            h[u] = pack_pair(h0, h1);
                            l[u] = pack_pair(to_tc(v0 - from_tc(h0)),
                                             to_tc(v1 - from_tc(h1)));
                        }
                        *reinterpret_cast<uint4 *>(beta_hi + base + 8 * k) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(beta_lo + base + 8 * k) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                }
            }
            if (live) bsq_part[(size_t)row_tile * (size_t)F + (size_t)f] = sq;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccs * kTileN) : "memory");
    }
}

// deterministic block sum (fixed tree), result valid in thread 0
__device__ __forceinline__ double tc_block_sum(double v, double *red /* [kThreadsTc / 32] shared */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreadsTc / 32; ++w) tot += red[w];
    __syncthreads();
    return tot;
}

// Residual step of the dense-design AMP on the tensor-core path (sparc_new.py:898,904-905), one CTA per frame:
//     z <- y - sum_s part_s + z (P - |beta|^2 / n) / tau^2        (first: z <- y),     tau^2 <- |z|^2 / n
// reading the K-split partials of A beta directly (no reduce pass) and the per-row-tile partials of |beta|^2, and
// writing z as fp32 plus the fp16 hi / lo pair the A^T z product consumes (no split pass).  n % 8 == 0.
__global__ void __launch_bounds__(kThreadsTc)
tc_residual_kernel(const float *__restrict__ y, const float *__restrict__ part, int splits, long long split_stride,
                   float *__restrict__ z, tc_t *__restrict__ z_hi, tc_t *__restrict__ z_lo,
                   const float *__restrict__ bsq_part, int n_bsq, double *__restrict__ tau_sqr, double *__restrict__ tau_trace,
                   int F, int n, double P, int first)
{
    __shared__ double red[kThreadsTc / 32];
    __shared__ float s_coef;
    const size_t f = blockIdx.x;
    if (!first) {
        double b = 0.0;
        for (int i = threadIdx.x; i < n_bsq; i += kThreadsTc) b += (double)bsq_part[(size_t)i * F + f];
        const double bsq = tc_block_sum(b, red);
        if (threadIdx.x == 0) s_coef = (float)((P - bsq / n) / tau_sqr[f]);
        __syncthreads();
    }
    const float coef = first ? 0.0f : s_coef;
    double acc = 0.0;
    for (int i8 = threadIdx.x; i8 < n / 8; i8 += kThreadsTc) {
        const size_t p = f * (size_t)n + (size_t)i8 * 8;
        float v[8];
        {
            const float4 a = *reinterpret_cast<const float4 *>(y + p), b = *reinterpret_cast<const float4 *>(y + p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
        if (!first) {
            float ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int s = 0; s < splits; ++s) {
                const float *ps = part + (size_t)s * (size_t)split_stride + p;
                const float4 a = *reinterpret_cast<const float4 *>(ps), b = *reinterpret_cast<const float4 *>(ps + 4);
                ab[0] += a.x; ab[1] += a.y; ab[2] += a.z; ab[3] += a.w; ab[4] += b.x; ab[5] += b.y; ab[6] += b.z; ab[7] += b.w;
            }
            const float4 a = *reinterpret_cast<const float4 *>(z + p), b = *reinterpret_cast<const float4 *>(z + p + 4);
            const float zo[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (v[k] - ab[k]) + zo[k] * coef;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const tc_t h0 = to_tc(v[2 * u]), h1 = to_tc(v[2 * u + 1]);
            h[u] = pack_pair(h0, h1);
            l[u] = pack_pair(to_tc(v[2 * u] - from_tc(h0)),
                             to_tc(v[2 * u + 1] - from_tc(h1)));
        }
        *reinterpret_cast<float4 *>(z + p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(z + p + 4) = make_float4(v[4], v[5], v[6], v[7]);
        *reinterpret_cast<uint4 *>(z_hi + p) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(z_lo + p) = make_uint4(l[0], l[1], l[2], l[3]);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += (double)v[k] * (double)v[k];
    }
    const double tot = tc_block_sum(acc, red);
    if (threadIdx.x == 0) {
        tau_sqr[f] = tot / n;
        if (tau_trace) tau_trace[f] = tot / n;
    }
}

// out = sum of `splits` partial products, added in index order (fixed result)
__global__ void reduce_splits_kernel(const float4 *__restrict__ part, float4 *__restrict__ out, long long n4, int splits)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n4; i += step) {
        float4 a = part[i];
        for (int s = 1; s < splits; ++s) {
            const float4 b = part[(long long)s * n4 + i];
            a.x += b.x;
            a.y += b.y;
            a.z += b.z;
            a.w += b.w;
        }
        out[i] = a;
    }
}

// x (fp32 or fp64) -> hi = fp16(x), lo = fp16(x - hi)
template <typename T>
__global__ void split_kernel(const T *__restrict__ x, tc_t *__restrict__ hi, tc_t *__restrict__ lo,
                             long long n, double pre)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        const double xv = (double)x[i] * pre;          // pre: power of two (exact)
        const tc_t h = to_tc((float)xv);
        hi[i] = h;
        lo[i] = to_tc((float)(xv - (double)from_tc(h)));
    }
}

// fp32 frame operand, 8 elements per thread: two 16-byte loads, one 16-byte store per output
// (the scalar kernel's 2-byte stores held it at 4.1 TB/s).  x - hi is exact in fp32 (hi is x rounded to
// 8 mantissa bits), so the residual needs no fp64.
__global__ void split8_kernel(const float4 *__restrict__ x, uint4 *__restrict__ hi, uint4 *__restrict__ lo, long long n8)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n8; i += step) {
        const float4 a = x[2 * i], b = x[2 * i + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const tc_t h0 = to_tc(v[2 * k]), h1 = to_tc(v[2 * k + 1]);
            const tc_t l0 = to_tc(v[2 * k] - from_tc(h0));
            const tc_t l1 = to_tc(v[2 * k + 1] - from_tc(h1));
            h[k] = (uint32_t)tc_bits(h0) | ((uint32_t)tc_bits(h1) << 16);
            l[k] = (uint32_t)tc_bits(l0) | ((uint32_t)tc_bits(l1) << 16);
        }
        hi[i] = make_uint4(h[0], h[1], h[2], h[3]);
        lo[i] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// out[c, r] = in[r, c]  (tiled through shared memory), in is [rows, cols]
__global__ void transpose_kernel(const double *__restrict__ in, double *__restrict__ out, int rows, int cols)
{
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 2-D fp16 tensor [rows, K] (K contiguous), box = 64 x 128, 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t K)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        lsb::set_error("cuTensorMapEncodeTiled entry point not available");
        return LSB_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {K * sizeof(tc_t)};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, kTcMapType, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        lsb::set_error("cuTensorMapEncodeTiled failed (%d) for [%llu x %llu]", (int)r, (unsigned long long)rows,
                       (unsigned long long)K);
        return LSB_ERR_CUDA;
    }
    return LSB_OK;
}

}  // namespace

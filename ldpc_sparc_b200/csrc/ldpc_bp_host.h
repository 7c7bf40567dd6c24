// Types shared between the BP kernels' translation units and the host API.
#pragma once
#include "common.cuh"

namespace lsb_bp {

constexpr int kMaxRows = 32;
constexpr int kMaxCols = 64;
constexpr int kMaxEdges = 160;
constexpr int kMaxDcQc = 24;   // largest check degree the QC kernel unrolls
constexpr int kMaxDc = 64;     // hard limit (generic kernel, local-memory trellis)
constexpr int kThreads = 256;

// Protograph tables of a QC code, copied to shared memory by every CTA.
struct QcTables {
    int z, rows, cols, n_edges, nv, nc;
    unsigned magic;                  // ceil(2^32 / z): p / z == umulhi(p, magic) for p < 65536
    short row_ptr[kMaxRows + 1];     // protograph edges of row r: [row_ptr[r], row_ptr[r+1])
    short row_order[kMaxRows];       // rows sorted by degree (stable)
    short col_ptr[kMaxCols + 1];
    int2 col_entry[kMaxEdges];       // per column-list entry: (edge*z - shift, shift mod z)
};

// Fused Monte-Carlo front end of the QC kernel (reference loop: ldpc_jossy/py/ldpc_awgn.py:88-107).  With
// cw != nullptr the kernel does not read LLRs: frame f carries codeword (frame0 + f) mod n_cw, BPSK-modulated
// (1 - 2 x, :55-56), plus sigma * N(0, 1) noise from Philox4x32-10 keyed by (seed, frame0 + f) -- the noise of a
// frame depends on its GLOBAL index only, not on the launch geometry or the number of GPUs -- scaled to LLRs by
// 2 / sigma^2 (:51-53); after the decode it counts the wrong hard decisions over all N code bits (:99).
struct McParams {
    const uint8_t *cw = nullptr;       // [n_cw, N] code bits (0 / 1), device
    long long n_cw = 0;
    float llr_scale = 0.f, sigma = 0.f;
    unsigned long long seed = 0, frame0 = 0;
    int32_t *bit_errors = nullptr;     // [F] wrong bits of the frame (optional)
    float *llr_out = nullptr;          // [F, N] the channel LLRs that were generated (optional; tests)
    unsigned long long *counters = nullptr;   // [4] += bit errors, frame errors, frames, iterations (optional)
    // independent of cw: hard decisions as packed bits, [F][ceil(N / 32)] words, bit v % 32 of word v / 32 (optional)
    uint32_t *hard_bits = nullptr;
};

template <typename T>
struct QcLaunch {
    const QcTables *tabs;
    const T *llr;
    T *app;
    uint8_t *hard;
    int32_t *iters;
    long long n_frames;
    int max_it, algo, max_dc, z;
    T scale;
    unsigned long long *queue;
    size_t smem;        // dynamic shared memory of one CTA
    int *cached_grid;   // per (handle, precision, algo) launch geometry cache
    cudaStream_t stream;
    McParams mc;
};

template <typename T>
struct GenericLaunch {
    const int *voff, *coff, *intrlv;
    int nv, nc, nmsg;
    const T *llr;
    T *app;
    uint8_t *hard;
    int32_t *iters;
    long long n_frames;
    int max_it, algo;
    T scale;
    T *ws;
    unsigned long long *queue;
    long long grid;
    cudaStream_t stream;
};

// defined in ldpc_bp_f32.cu / ldpc_bp_f64.cu
template <typename T> int launch_qc(const QcLaunch<T> &a);
template <typename T> int launch_generic(const GenericLaunch<T> &a);
template <typename T> int generic_blocks_per_sm();
// scalar helpers behind the legacy Lxor / Lxfb symbols (ldpc_bp_f64.cu)
void launch_lxor(double a, double b, int corr, double *d_out);
void launch_lxfb(double *d_L, int dc, int corr, double *d_out);

}  // namespace lsb_bp

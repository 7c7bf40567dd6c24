/*
 * ldpc_sparc_b200.h -- C-ABI of libldpcsparc_b200.so (sm_100a CUDA inside).
 *
 * This is the drop-in boundary for the iterative-decoding hot path of
 * SophieLangdon27/LDPC_SPARC.  Two groups of entry points:
 *
 *  (1) LEGACY symbols: exactly what the reference's only FFI binds
 *      (ldpc_jossy/py/ldpc.py:467-503 does ct.CDLL(...).sumprod2(...) etc. on
 *      ldpc_jossy/src/c_ldpc.c).  Host pointers, doubles, one frame per call,
 *      arbitrary Tanner graph as (vdeg, cdeg, intrlv).  The computation runs on
 *      the GPU in fp64 (there is no CPU code path in this library).
 *
 *  (2) BATCHED symbols (lsb_*): the B200-native interface.  Fixed-width types,
 *      opaque device-resident handles, many independent frames per call,
 *      fp32 (throughput) or fp64 (parity) arithmetic, device pointers (+stream)
 *      or host pointers (copies pipelined inside the call).
 *
 * Conventions for lsb_* functions: return 0 on success, a negative lsb_status
 * otherwise; lsb_last_error() gives a thread-local message.  Nothing throws
 * across this boundary.  All device work is issued on the CURRENT CUDA device
 * (the host side sets it, e.g. torch.cuda.set_device(local_rank)); a handle
 * must be used on the device it was created on.  An empty batch (zero frames)
 * is valid with any pointer values and returns 0 without touching the device.
 *
 * LLR sign convention (as the reference): positive = bit 0 more likely.
 */
#ifndef LDPC_SPARC_B200_H
#define LDPC_SPARC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum lsb_status {
    LSB_OK = 0,
    LSB_ERR_INVALID = -1,     /* bad argument (null pointer, size, degree...) */
    LSB_ERR_CUDA = -2,        /* a CUDA runtime call failed                   */
    LSB_ERR_NO_DEVICE = -3,   /* no sm_100 device visible                     */
    LSB_ERR_UNSUPPORTED = -4, /* valid request this build does not implement  */
    LSB_ERR_NOMEM = -5        /* host or device allocation failed             */
};

/* check-node rule selector for the batched decoders */
enum lsb_bp_algo {
    LSB_BP_SUMPROD2 = 0, /* forward/backward Lxor trellis, exact correction (c_ldpc.c:138-206) */
    LSB_BP_SUMPROD = 1,  /* tanh-product rule                               (c_ldpc.c:32-113)  */
    LSB_BP_MINSUM = 2    /* Lxor without correction, scaled                 (c_ldpc.c:339-381) */
};

int lsb_version(void);              /* ABI version, currently 1               */
const char *lsb_last_error(void);   /* thread-local, never NULL               */
int lsb_device_count(void);         /* number of visible CUDA devices (>= 0)  */
/* Number of kernels this library has launched since load / last reset
 * (bench.py reports it as gpu_launches). */
int64_t lsb_launch_count(void);
void lsb_launch_count_reset(void);

/* ------------------------------------------------------- (1) legacy symbols
 * Replaces ldpc_jossy/src/c_ldpc.c.  `long` is the platform C long, as in the
 * reference prototype (c_ldpc.c:32-33,138-139,339-340); on LP64 Linux that is
 * 64 bits.  Return value: number of iterations in [0, max_itcount], or -1 on
 * allocation / device failure (the reference returns -1 when calloc fails,
 * c_ldpc.c:41-42,165-166).  ch, graph arrays: read-only; app: overwritten.
 */
int sumprod(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc,
            int Nmsg, double *app, int max_itcount);           /* c_ldpc.c:32  */
int sumprod2(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc,
             int Nmsg, double *app, int max_itcount);          /* c_ldpc.c:138 */
int minsum(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc,
           int Nmsg, double *app, double correction_factor,
           int max_itcount);                                   /* c_ldpc.c:339 */
double Lxor(double L1, double L2, int corr_flag);              /* c_ldpc.c:234 */
double Lxfb(double *L, long dc, int corr_flag);                /* c_ldpc.c:294 */

/* ------------------------------------------------ (2) batched LDPC BP decode
 * Graph handles.  A QC graph is built straight from the protograph that
 * ldpc.py:24-272 (assign_proto) holds: proto[r*cols + c] = -1 for an all-zero
 * block, otherwise the right cyclic shift of the z x z identity (taken mod z,
 * ldpc.py:376).  Its message numbering is the reference's: the message array
 * is check-ordered with ports in increasing protograph column, a variable's
 * ports are in increasing check index (ldpc.py:303-396), so summation orders
 * and hence results are those of the reference.  A generic graph takes the
 * reference's (vdeg, cdeg, intrlv) triple as is.
 */
typedef struct lsb_ldpc_graph lsb_ldpc_graph;

int lsb_ldpc_graph_create_qc(const int32_t *proto, int rows, int cols, int z,
                             lsb_ldpc_graph **out);
int lsb_ldpc_graph_create(const int64_t *vdeg, const int64_t *cdeg,
                          const int64_t *intrlv, int64_t nv, int64_t nc,
                          int64_t nmsg, lsb_ldpc_graph **out);
void lsb_ldpc_graph_destroy(lsb_ldpc_graph *g);
/* nv, nc, nmsg, is_qc, z */
int lsb_ldpc_graph_info(const lsb_ldpc_graph *g, int64_t info[5]);

/* Systematic encoder on the device (QC graphs only), replaces code.encode
 * (ldpc.py:400-460) for Monte-Carlo frame generation: info [n_frames, K] bytes
 * (0/1) -> codewords [n_frames, N] bytes (0/1); asynchronous on `stream`. */
int lsb_ldpc_encode_dev(const lsb_ldpc_graph *g, const uint8_t *info, uint8_t *codewords,
                        int64_t n_frames, void *stream);

/* Device-pointer decode, asynchronous on `stream` (a cudaStream_t; NULL = the
 * legacy default stream).  Replaces a loop of `code.decode(ch[f], max_it)`
 * (ldpc.py:463-490) over n_frames independent frames.
 *   llr   [n_frames, nv]  in   channel LLRs
 *   app   [n_frames, nv]  out  a-posteriori LLRs of the last executed VN pass
 *                              (may be NULL if only hard decisions are wanted)
 *   hard  [n_frames, nv]  out  uint8 app < 0           (may be NULL)
 *   iters [n_frames]      out  iteration count as the reference returns it
 *   scale                 in   min-sum correction factor (ignored otherwise)
 */
int lsb_ldpc_decode_dev_f32(const lsb_ldpc_graph *g, const float *llr,
                            float *app, uint8_t *hard, int32_t *iters,
                            int64_t n_frames, int max_it, int algo,
                            double scale, void *stream);
int lsb_ldpc_decode_dev_f64(const lsb_ldpc_graph *g, const double *llr,
                            double *app, uint8_t *hard, int32_t *iters,
                            int64_t n_frames, int max_it, int algo,
                            double scale, void *stream);

/* Fused Monte-Carlo step (replaces the body of the reference's simulation loop,
 * ldpc_jossy/py/ldpc_awgn.py:88-107: bpsk -> awgn -> ch2llr -> decode -> count) for QC
 * graphs.  Frame f of the call carries codeword (frame0 + f) mod n_codewords of the device
 * array codewords [n_codewords, N] (0 / 1, e.g. from lsb_ldpc_encode_dev), BPSK-modulated,
 * plus N(0, sigma2) noise from Philox4x32-10 keyed by (seed, frame0 + f): the noise of a
 * frame depends on its global index only, so results do not change with the batch split
 * or the number of GPUs.  Nothing but the codeword pool is read from HBM and nothing but
 * the requested outputs is written: bit_errors [F] wrong hard decisions over all N code
 * bits (:99), iters [F], llr_out [F, N] the generated channel LLRs (tests), counters [4]
 * += {bit errors, frame errors, frames, iterations} (each optional, device pointers). */
int lsb_ldpc_mc_dev_f32(const lsb_ldpc_graph *g, const uint8_t *codewords, int64_t n_codewords,
                        double sigma2, uint64_t seed, uint64_t frame0, int64_t n_frames,
                        int max_it, int algo, double scale, int32_t *bit_errors, int32_t *iters,
                        float *llr_out, uint64_t *counters, void *stream);

/* Host-pointer decode: the call copies llr host->device, decodes and copies
 * the requested outputs back, in chunks pipelined over internal streams, and
 * returns when the outputs are complete.  Buffers may be pageable or pinned
 * (pinned overlaps copies with compute).  compute_f64 = 0 computes in fp32,
 * 1 in fp64 (the I/O type is that of the function name).
 */
int lsb_ldpc_decode_host_f32(const lsb_ldpc_graph *g, const float *llr,
                             float *app, uint8_t *hard, int32_t *iters,
                             int64_t n_frames, int max_it, int algo,
                             double scale);
/* The same with bit-packed hard decisions as the only output besides the
 * iteration counts: bits[f * ceil(N / 32) + v / 32] >> (v % 32) & 1 = (app[f, v] < 0),
 * what the reference's Monte-Carlo loop keeps of a decode (ldpc_awgn.py:97);
 * 1/8 of the device->host bytes of `hard`, 1/32 of `app`. */
int lsb_ldpc_decode_host_bits_f32(const lsb_ldpc_graph *g, const float *llr,
                                  uint32_t *bits, int32_t *iters,
                                  int64_t n_frames, int max_it, int algo,
                                  double scale);
int lsb_ldpc_decode_host_f64(const lsb_ldpc_graph *g, const double *llr,
                             double *app, uint8_t *hard, int32_t *iters,
                             int64_t n_frames, int max_it, int algo,
                             double scale, int compute_f64);

/* Pinned host memory helpers (so a ctypes/cgo host can stage buffers that
 * overlap with compute).  */
void *lsb_host_alloc(uint64_t bytes);
void lsb_host_free(void *p);

/* ------------------------------------- (3) SPARC, sub-sampled DCT design matrix
 * Replaces the closures sparc_transforms() returns (sparc_public/sparc.py:703-880:
 * Ab(x) = A x, Az(y) = A^T y) and the AMP loop sparc_amp() (sparc.py:883-999) for
 * real, unmodulated SPARCs (K = 1), for many frames that share one design.
 *
 * A design handle holds, on the device, the base matrix W [Lr x Lc] (a scalar
 * W is 1x1, a power-allocation vector 1xB, sparc.py:570-589) and per non-zero
 * block the row / column orderings that generate_ordering() draws
 * (sparc.py:735-775): order0 [Lr, Lc, Mr] and order1 [Lr, Lc, Mc], entries in
 * [1, w), Mr = n/Lr, Mc = L*M/Lc, w = 2^ceil(log2(max(Mr, Mc) + 1)); entries of
 * zero blocks are ignored.  Block (r,c) of A is
 *   sqrt(W[r,c]/L) * sqrt(2) * cos(pi * order0[i] * (2*order1[j] + 1) / (2w)).
 */
typedef struct lsb_sparc_design lsb_sparc_design;

int lsb_sparc_design_create(const double *W, int Lr, int Lc, int L, int M, int n,
                            const uint32_t *order0, const uint32_t *order1,
                            lsb_sparc_design **out);
void lsb_sparc_design_destroy(lsb_sparc_design *g);
/* L, M, n, Lr, Lc, w, non-zero blocks, shared-memory bytes of the fp32 AMP kernel */
int lsb_sparc_design_info(const lsb_sparc_design *g, int64_t info[8]);

/* out[f, n] = A x[f, L*M]  and  out[f, L*M] = A^T y[f, n]; device pointers,
 * asynchronous on `stream`. */
int lsb_sparc_ab_dev_f32(const lsb_sparc_design *g, const float *x, float *out,
                         int64_t n_frames, void *stream);
int lsb_sparc_ab_dev_f64(const lsb_sparc_design *g, const double *x, double *out,
                         int64_t n_frames, void *stream);
int lsb_sparc_az_dev_f32(const lsb_sparc_design *g, const float *y, float *out,
                         int64_t n_frames, void *stream);
int lsb_sparc_az_dev_f64(const lsb_sparc_design *g, const double *y, double *out,
                         int64_t n_frames, void *stream);

/* decode_params of the reference (sparc.py:151-170) + the channel noise variance */
typedef struct lsb_sparc_amp_params {
    double awgn_var;
    double rtol;           /* early-stop relative tolerance on psi (default 1e-6)   */
    double atol;           /* absolute tolerance; the reference uses 2e-15 (:916)   */
    int32_t t_max;         /* the loop runs at most t_max - 1 iterations (:931)     */
    int32_t phi_est_method;/* 1: awgn_var + gamma, 2: mean(z^2)  (:949-955)         */
    int32_t modulation;    /* K of the reference: 0 or 1 = unmodulated, 2 = real    *
                            * 2-PSK values +1 / -1 (:433-441, 488-492); complex     *
                            * designs (lsb_csparc_*) take any power of two <= 64    */
    int32_t reserved;      /* must be 0                                             */
} lsb_sparc_amp_params;

/* Outputs of one batched AMP decode (device pointers; optional ones may be NULL).
 * beta_soft / s_last have the element type of the entry point. */
typedef struct lsb_sparc_amp_out {
    int32_t *idx;      /* [F, L]          MAP decision: argmax of s per section (:997);   *
                        *                  modulation 2: argmax|s| + M if that s < 0   */
    int32_t *t_final;  /* [F]             iterations run, as the reference returns it  */
    double *psi;       /* [F, Lc]         final psi                                    */
    float *nmse;       /* [F, t_max, Lc]  NMSE per iteration; needs idx0   (optional)  */
    void *beta_soft;   /* [F, L*M]        beta of the last iteration       (optional)  */
    void *s_last;      /* [F, L*M]        s of the last iteration          (optional)  */
    double *tau_trace; /* [F, t_max, Lc]  tau per iteration                (optional)  */
    double *psi_trace; /* [F, t_max, Lc]  psi per iteration                (optional)  */
} lsb_sparc_amp_out;

/* y [F, n] received words; idx0 [F, L] true section indices (NULL unless nmse is
 * wanted -- the reference takes beta0 for the same purpose, sparc.py:889), coded like
 * out->idx (location + M for a -1 value when modulation = 2). */
int lsb_sparc_amp_dev_f32(lsb_sparc_design *g, const float *y, int64_t n_frames,
                          const lsb_sparc_amp_params *p, const int32_t *idx0,
                          const lsb_sparc_amp_out *out, void *stream);
int lsb_sparc_amp_dev_f64(lsb_sparc_design *g, const double *y, int64_t n_frames,
                          const lsb_sparc_amp_params *p, const int32_t *idx0,
                          const lsb_sparc_amp_out *out, void *stream);

/* Host-pointer form of the batched decode: what a loop over the reference's
 * sparc_decode (sparc_public/sparc.py:55-74, called per frame by sparc_sim.py:21-24)
 * takes and returns.  y [F, n] received words in host memory (pinned memory makes the
 * copies asynchronous); out: idx [F, L] MAP section decisions, t_final [F], psi
 * [F, Lc] final psi (optional, NULL to skip; expect_err of the reference is
 * mean(psi) >= 1e-3).  Returns when the outputs are in host memory. */
int lsb_sparc_amp_host_f32(lsb_sparc_design *g, const float *y, int64_t n_frames,
                           const lsb_sparc_amp_params *p, int32_t *idx, int32_t *t_final,
                           double *psi);

/* ---- (3b) the same design stored explicitly, products on the tensor cores (csrc/sparc_band_tc.cu)
 * For designs whose blocks keep few of the w transform outputs (spatially coupled: BASELINE config 4 keeps 332 of
 * 65536 per block) the blocks of A are generated once on the device from the closed form above and kept as fp16
 * hi / lo pairs in both orientations; A beta and A^T z for a frame batch are block-banded tcgen05 GEMMs (three
 * fp16 products per operand pair, fp32 accumulation) and the rest of sparc_amp (sparc.py:931-999) runs in two
 * memory-bound kernels per iteration.  Same arguments and outputs as lsb_sparc_design_create /
 * lsb_sparc_ab_dev_f32 / lsb_sparc_az_dev_f32 / lsb_sparc_amp_dev_f32 (modulation <= 1 only; M <= 1024; L*M a
 * multiple of 8).  Device memory: lsb_band_design_bytes().  lsb_band_amp_dev_f32 enqueues on `stream` but also
 * synchronises it every fourth iteration (it reads back how many frames have stopped), so it cannot be captured
 * into a CUDA graph. */
typedef struct lsb_band_design lsb_band_design;
int lsb_band_design_create(const double *W, int Lr, int Lc, int L, int M, int n,
                           const uint32_t *order0, const uint32_t *order1, lsb_band_design **out);
void lsb_band_design_destroy(lsb_band_design *g);
int64_t lsb_band_design_bytes(const lsb_band_design *g);
int lsb_band_ab_dev_f32(lsb_band_design *g, const float *x, float *out, int64_t n_frames, void *stream);
int lsb_band_az_dev_f32(lsb_band_design *g, const float *y, float *out, int64_t n_frames, void *stream);
int lsb_band_amp_dev_f32(lsb_band_design *g, const float *y, int64_t n_frames,
                         const lsb_sparc_amp_params *p, const int32_t *idx0,
                         const lsb_sparc_amp_out *out, void *stream);
/* Diagnostic (host only, no GPU call): the order in which the fp32 AMP kernel adds one design block's rows into the
 * residual.  slot_ids[i] = shared-memory slot (8-byte units) row i reads, row0 = first residual row of the block.
 * words[(chunk * 256 + thread) * 4 + s] = row | slot << 16 or 0xFFFFFFFF; within every half-warp (16 consecutive
 * threads) of a slot position the rows' residual bank pairs ((row0 + row) % 16) and the slots' bank pairs (slot % 16)
 * are all different, and every row appears once.  max_words >= 4096 * ceil(degree / 64) (16384 suffices for 4095 rows). */
int lsb_sparc_schedule_row_gather(const uint16_t *slot_ids, int n_rows, int row0,
                                  uint32_t *words, int64_t max_words, int *chunks);

/* host-pointer form, as lsb_sparc_amp_host_f32 */
int lsb_band_amp_host_f32(lsb_band_design *g, const float *y, int64_t n_frames,
                          const lsb_sparc_amp_params *p, int32_t *idx, int32_t *t_final, double *psi);

/* ----------------------------- (4) SPARC, dense Gaussian design (sparc_sophie)
 * Replaces, for batches of frames that share one design matrix A [n x ML]
 * (sparc_new.py:1284-1294), the NumPy steps of sparc_amp / sparc_amp_single_it
 * (sparc_new.py:885-912, 975-990) and the pure-Python section <-> bit glue around
 * the BP decoder (sparc_new.py:1118-1138, 1162-1193, 1260-1279, 824-883).  The
 * host side (ldpc_sparc_b200/sparc_new.py) chains these with lsb_ldpc_decode_dev_*
 * into sparc_ldpc_decode / naively_integrated_decoder / integrated_decoder.
 * All pointers are device pointers; per-frame scalars (tau_sqr, beta_sq,
 * onsager_sum) are double[F]; calls are asynchronous on `stream`.
 */
typedef struct lsb_dense_design lsb_dense_design;

int lsb_dense_design_create(const double *A /* host, row-major [n, ML] */, int64_t n, int64_t ML,
                            int want_f32, int want_f64, lsb_dense_design **out);
void lsb_dense_design_destroy(lsb_dense_design *g);

/* out[F, n] = x[F, ML] A^T  (np.dot(A, beta), :903)  /  out[F, ML] = z[F, n] A  (np.dot(AT, z), :907) */
int lsb_dense_ab_dev_f32(const lsb_dense_design *g, const float *x, float *out, int64_t F, void *stream);
int lsb_dense_ab_dev_f64(const lsb_dense_design *g, const double *x, double *out, int64_t F, void *stream);
int lsb_dense_az_dev_f32(const lsb_dense_design *g, const float *z, float *out, int64_t F, void *stream);
int lsb_dense_az_dev_f64(const lsb_dense_design *g, const double *z, double *out, int64_t F, void *stream);

/* z <- y - ab + z*coef, tau_sqr <- |z|^2/n.  first != 0: z <- y.  mode 0: coef =
 * (P - beta_sq/n)/tau_sqr (:904); mode 1: coef = onsager_sum/n (:492). */
int lsb_amp_residual_dev_f32(const float *y, const float *ab, float *z, const double *beta_sq,
                             const double *onsager_sum, double *tau_sqr, int64_t F, int n, double P,
                             int mode, int first, void *stream);
int lsb_amp_residual_dev_f64(const double *y, const double *ab, double *z, const double *beta_sq,
                             const double *onsager_sum, double *tau_sqr, int64_t F, int n, double P,
                             int mode, int first, void *stream);
/* s = atz + beta_in (beta_in may be NULL = 0, and may alias beta_out);
 * beta_out <- sqrt_nPl softmax_section(sqrt_nPl s / tau_sqr) (:1040-1066);
 * beta_sq <- |beta_out|^2; s_out optional. */
int lsb_amp_denoise_dev_f32(const float *beta_in, const float *atz, float *beta_out, float *s_out,
                            const double *tau_sqr, double *beta_sq, int64_t F, int L, int M,
                            double sqrt_nPl, void *stream);
int lsb_amp_denoise_dev_f64(const double *beta_in, const double *atz, double *beta_out, double *s_out,
                            const double *tau_sqr, double *beta_sq, int64_t F, int L, int M,
                            double sqrt_nPl, void *stream);
/* beta [n_sections, M] -> probs (P(bit = 0), optional) and llr (optional), each [n_sections, logM] (:1118-1138,1166-1169) */
int lsb_section_to_bit_dev_f32(const float *beta, float *probs, float *llr, int64_t n_sections, int M,
                               double sqrt_nPl, void *stream);
int lsb_section_to_bit_dev_f64(const double *beta, double *probs, double *llr, int64_t n_sections, int M,
                               double sqrt_nPl, void *stream);
/* app [F, L, logM] -> probs = sigmoid(app) (optional), beta [F, L, M] (:1190, 1260-1279) and
 * beta_sq[f] = |beta_f|^2 (optional) */
int lsb_bit_to_section_dev_f32(const float *app, float *probs, float *beta, double *beta_sq, int64_t F,
                               int L, int M, double sqrt_nPl, void *stream);
int lsb_bit_to_section_dev_f64(const double *app, double *probs, double *beta, double *beta_sq, int64_t F,
                               int L, int M, double sqrt_nPl, void *stream);
/* bits[blk, k] = app[blk, k] < 0 for the K systematic positions of each N-bit block (:1184-1186) */
int lsb_hard_info_bits_dev_f32(const float *app, uint8_t *bits, int64_t n_blocks, int N, int K, void *stream);
int lsb_hard_info_bits_dev_f64(const double *app, uint8_t *bits, int64_t n_blocks, int N, int K, void *stream);
/* out[f] = sum_i beta_i * main_term_i of differentiated_eta_calc (:824-841, 871-883) */
int lsb_integrated_onsager_dev_f32(const float *beta, const float *vk, const float *vk0,
                                   const float *alpha_weighted, const double *tau_sqr, double *out,
                                   int64_t F, int L, int M, double sqrt_nPl, void *stream);
int lsb_integrated_onsager_dev_f64(const double *beta, const double *vk, const double *vk0,
                                   const double *alpha_weighted, const double *tau_sqr, double *out,
                                   int64_t F, int L, int M, double sqrt_nPl, void *stream);

/* Tensor-core variant of the two dense products (tcgen05 + TMEM + TMA, csrc/dense_tc.cu):
 * operands split into two fp16 terms, three MMAs per tile into one fp32 accumulator, so
 * results carry ~16 mantissa bits (about 2e-5 relative), enough for the 1e-4 bound on
 * tau^2 / beta.  fp32 in / fp32 out; n and ML must be multiples of 8.  The handle holds
 * fp16 hi/lo copies of A and of A^T (8 bytes per entry in total). */
typedef struct lsb_tc_design lsb_tc_design;
int lsb_tc_design_create(const double *A /* host, row-major [n, ML] */, int64_t n, int64_t ML,
                         lsb_tc_design **out);
void lsb_tc_design_destroy(lsb_tc_design *g);
int lsb_tc_ab_dev_f32(lsb_tc_design *g, const float *x /* [F, ML] */, float *out /* [F, n] */, int64_t F,
                      void *stream);
int lsb_tc_az_dev_f32(lsb_tc_design *g, const float *z /* [F, n] */, float *out /* [F, ML] */, int64_t F,
                      void *stream);
/* Whole dense-design AMP decode on the tensor-core path (replaces the loop of sparc_sophie/sparc_new.py:885-912
 * for a batch): per iteration  A beta (tcgen05 GEMM) -> residual + tau^2 -> A^T z with the section-wise softmax
 * denoiser (sparc_new.py:1040-1066) as the GEMM epilogue.  y [F, n] device fp32; beta_out [F, ML] receives the soft
 * estimate after t_max iterations; s_out (may be NULL) the last s = beta + A^T z; tau_trace (may be NULL) [t_max, F]
 * the tau^2 of every iteration.  M must be a power of two in [2, 128]; pointers 16-byte aligned. */
int lsb_tc_amp_dev_f32(lsb_tc_design *g, const float *y, int64_t F, int L, int M, double P, double sqrt_nPl, int t_max,
                       float *beta_out, float *s_out, double *tau_trace, int32_t *idx_out /* [F, L] MAP decisions, may be NULL */,
                       void *stream);
/* host-pointer form: y [F, n] host -> idx [F, L] host (argmax of the last s per section, sparc_new.py:79-80) */
int lsb_tc_amp_host_f32(lsb_tc_design *g, const float *y, int64_t F, int L, int M, double P, double sqrt_nPl, int t_max,
                        int32_t *idx);

/* ---------------------------------------------------------------------------
 * Complex SPARCs: sub-sampled DFT design (sparc.py sub_fft :593-646, sparc_transforms
 * :703-880 with csparc=True) and the AMP decoder on complex observations with K-PSK
 * modulated sections (sparc_amp :883-999; estimators :402-512, all K branches).
 * Complex vectors are interleaved (re, im) pairs of the entry point's element type.
 * order0 / order1 as for lsb_sparc_design_create, entries in [0, w) with
 * w = 2^ceil(log2(max(Mr + 2, Mc + 2))) (:747).  A section symbol (idx0, out->idx) is
 * location + M * k, k the index into the K-PSK constellation psk_constel(K) (:225-239);
 * p->modulation = K (0 or 1: unmodulated). */
typedef struct lsb_csparc_design lsb_csparc_design;
int lsb_csparc_design_create(const double *W, int Lr, int Lc, int L, int M, int n, const uint32_t *order0,
                             const uint32_t *order1, lsb_csparc_design **out);
void lsb_csparc_design_destroy(lsb_csparc_design *g);
int lsb_csparc_ab_dev_f32(lsb_csparc_design *g, const void *x /* [F, LM] complex */, void *out /* [F, n] */,
                          int64_t n_frames, void *stream);
int lsb_csparc_ab_dev_f64(lsb_csparc_design *g, const void *x, void *out, int64_t n_frames, void *stream);
int lsb_csparc_az_dev_f32(lsb_csparc_design *g, const void *y /* [F, n] complex */, void *out /* [F, LM] */,
                          int64_t n_frames, void *stream);
int lsb_csparc_az_dev_f64(lsb_csparc_design *g, const void *y, void *out, int64_t n_frames, void *stream);
int lsb_csparc_amp_dev_f32(lsb_csparc_design *g, const void *y /* [F, n] complex */, int64_t n_frames,
                           const lsb_sparc_amp_params *p, const int32_t *idx0, const lsb_sparc_amp_out *out,
                           void *stream);
int lsb_csparc_amp_dev_f64(lsb_csparc_design *g, const void *y, int64_t n_frames, const lsb_sparc_amp_params *p,
                           const int32_t *idx0, const lsb_sparc_amp_out *out, void *stream);

/* ---------------------------------------------------------------------------
 * State evolution (sparc_public/sparc_se.py).  Replaces sparc_se_E (:82-118):
 *   E[t] = mean over the mc_samples rows of u [mc_samples, M] (device, standard
 *   normals, row-major) of A / (A + sum_{j>=1} C_j) at noise level tau[t], K = 1
 *   (exp) or K = 2 (sinh / cosh).  tau and E are HOST arrays of n_tau
 *   entries (the column blocks of one iteration); the call returns after the
 *   stream has finished.  sparc_se's outer recursion (:120-183) stays on the host. */
int lsb_sparc_se_expect_f64(const double *u_dev, int64_t mc_samples, int M, int K, const double *tau_host,
                            int n_tau, double *E_host, void *stream);
int lsb_sparc_se_expect_f32(const float *u_dev, int64_t mc_samples, int M, int K, const double *tau_host,
                            int n_tau, double *E_host, void *stream);
/* complex samples u [mc_samples, M] of interleaved (re, im), K-PSK with K = 4 .. 64 (:100-113) */
int lsb_sparc_se_expect_c128(const double *u_dev, int64_t mc_samples, int M, int K, const double *tau_host,
                             int n_tau, double *E_host, void *stream);
int lsb_sparc_se_expect_c64(const float *u_dev, int64_t mc_samples, int M, int K, const double *tau_host,
                            int n_tau, double *E_host, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_SPARC_B200_H */

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
 * must be used on the device it was created on.
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
int lsb_ldpc_decode_host_f64(const lsb_ldpc_graph *g, const double *llr,
                             double *app, uint8_t *hard, int32_t *iters,
                             int64_t n_frames, int max_it, int algo,
                             double scale, int compute_f64);

/* Pinned host memory helpers (so a ctypes/cgo host can stage buffers that
 * overlap with compute).  */
void *lsb_host_alloc(uint64_t bytes);
void lsb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_SPARC_B200_H */

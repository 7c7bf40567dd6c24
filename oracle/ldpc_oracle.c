/*
 * ldpc_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the flooding belief-propagation decoder of the
 * reference, written from its algorithmic description.  It exists only so
 * that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs can check (and time) the CUDA path against it.
 * Nothing under ldpc_sparc_b200/ may link, import or call this file.
 *
 * Reference followed (paths relative to /root/reference):
 *   orc_lxor      <- ldpc_jossy/src/c_ldpc.c:234-251   (Lxor)
 *   orc_lxfb      <- ldpc_jossy/src/c_ldpc.c:294-314   (Lxfb)
 *   orc_sumprod2  <- ldpc_jossy/src/c_ldpc.c:138-206   (sumprod2, the default)
 *   orc_sumprod   <- ldpc_jossy/src/c_ldpc.c:32-113    (sumprod, tanh rule)
 *   orc_minsum    <- ldpc_jossy/src/c_ldpc.c:339-381   (minsum, with the loop
 *                    increment bug at :364 FIXED -- the reference README:89
 *                    documents minsum as broken; this is the intended rule)
 *
 * Parity pinning: tests/test_oracle_ldpc.py checks this file bit-for-bit
 * against oracle/_ref/c_ldpc_ref.so (the reference's own c_ldpc.c compiled
 * in place by oracle/Makefile) on random noisy frames, and against the
 * committed golden vectors tests/golden/ldpc_*.npz that were produced by
 * that reference build (oracle/make_golden.py).
 *
 * Graph description (same as the reference's): messages live in an array
 * ordered by check node (check c owns cdeg[c] consecutive slots); variable
 * node v owns vdeg[v] consecutive entries of `intrlv`, each the slot index
 * of one of its edges.  All integer arrays are int64 here (fixed width; the
 * reference's `long` is 32-bit on its MinGW build, 64-bit on LP64).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_DC 64

double orc_lxor(double a, double b, int corrected)
{
    double mag = fmin(fabs(a), fabs(b));
    double out = (signbit(a) == signbit(b)) ? mag : -mag;
    if (corrected) {
        out += log(1 + exp(-fabs(a + b)));
        out -= log(1 + exp(-fabs(a - b)));
    }
    return out;
}

/* In-place extrinsic update of one check node's dc incoming LLRs.
 * Returns the LLR of the XOR of all dc inputs (the check's soft syndrome). */
double orc_lxfb(double *m, int64_t dc, int corrected)
{
    double fwd[ORC_MAX_DC], bwd[ORC_MAX_DC];
    int64_t i;
    fwd[0] = m[0];
    bwd[dc - 1] = m[dc - 1];
    /* the reference interleaves the two recursions in one loop; they are
     * independent, so the order between them does not change any value */
    for (i = 1; i < dc; ++i) {
        fwd[i] = orc_lxor(fwd[i - 1], m[i], corrected);
        bwd[dc - 1 - i] = orc_lxor(bwd[dc - i], m[dc - 1 - i], corrected);
    }
    m[0] = bwd[1];
    m[dc - 1] = fwd[dc - 2];
    for (i = 1; i < dc - 1; ++i)
        m[i] = orc_lxor(fwd[i - 1], bwd[i + 1], corrected);
    return bwd[0];
}

/* variable-node half-iteration shared by all three decoders */
static void orc_vn_pass(const double *ch, const int64_t *vdeg,
                        const int64_t *intrlv, int64_t nv, double *msg,
                        double *app)
{
    int64_t v, e = 0, k;
    for (v = 0; v < nv; ++v) {
        double total = ch[v];
        for (k = 0; k < vdeg[v]; ++k)
            total += msg[intrlv[e + k]];
        for (k = 0; k < vdeg[v]; ++k)
            msg[intrlv[e + k]] = total - msg[intrlv[e + k]];
        app[v] = total;
        e += vdeg[v];
    }
}

int orc_sumprod2(const double *ch, const int64_t *vdeg, const int64_t *cdeg,
                 const int64_t *intrlv, int64_t nv, int64_t nc, int64_t nmsg,
                 double *app, int max_it)
{
    double *msg = (double *)calloc((size_t)nmsg, sizeof(double));
    int it;
    if (!msg) return -1;
    for (it = 0; it < max_it; ++it) {
        int64_t c, base = 0;
        int unsatisfied = 0;
        orc_vn_pass(ch, vdeg, intrlv, nv, msg, app);
        for (c = 0; c < nc; ++c) {
            double syn = orc_lxfb(msg + base, cdeg[c], 1);
            if (syn <= 0.0) unsatisfied = 1;
            base += cdeg[c];
        }
        if (!unsatisfied) break;
    }
    free(msg);
    return it;
}

int orc_minsum(const double *ch, const int64_t *vdeg, const int64_t *cdeg,
               const int64_t *intrlv, int64_t nv, int64_t nc, int64_t nmsg,
               double *app, double scale, int max_it)
{
    double *msg = (double *)calloc((size_t)nmsg, sizeof(double));
    int it;
    if (!msg) return -1;
    for (it = 0; it < max_it; ++it) {
        int64_t c, base = 0, k;
        int unsatisfied = 0;
        orc_vn_pass(ch, vdeg, intrlv, nv, msg, app);
        for (c = 0; c < nc; ++c) {
            double syn = orc_lxfb(msg + base, cdeg[c], 0);
            if (syn <= 0.0) unsatisfied = 1;
            for (k = 0; k < cdeg[c]; ++k) msg[base + k] *= scale;
            base += cdeg[c];
        }
        if (!unsatisfied) break;
    }
    free(msg);
    return it;
}

int orc_sumprod(const double *ch, const int64_t *vdeg, const int64_t *cdeg,
                const int64_t *intrlv, int64_t nv, int64_t nc, int64_t nmsg,
                double *app, int max_it)
{
    double *msg = (double *)calloc((size_t)nmsg, sizeof(double));
    int it;
    if (!msg) return -1;
    for (it = 0; it < max_it; ++it) {
        int64_t c, base = 0, k;
        int unsatisfied = 0;
        orc_vn_pass(ch, vdeg, intrlv, nv, msg, app);
        for (c = 0; c < nc; ++c) {
            double prod = 1.0;
            for (k = 0; k < cdeg[c]; ++k) {
                msg[base + k] = tanh(msg[base + k] / 2.0);
                prod *= msg[base + k];
            }
            /* the reference only evaluates atanh until the first failing
             * check (short-circuit); the flag value is the same either way */
            if (!unsatisfied && 2.0 * atanh(prod) <= 0.0) unsatisfied = 1;
            for (k = 0; k < cdeg[c]; ++k)
                msg[base + k] = 2.0 * atanh(prod / msg[base + k]);
            base += cdeg[c];
        }
        if (!unsatisfied) break;
    }
    free(msg);
    return it;
}

/* Batched convenience wrapper used by the tests and the CPU baseline:
 * frames are rows of ch[f*nv .. f*nv+nv), same graph for every frame. */
void orc_sumprod2_batch(const double *ch, const int64_t *vdeg,
                        const int64_t *cdeg, const int64_t *intrlv, int64_t nv,
                        int64_t nc, int64_t nmsg, double *app, int32_t *iters,
                        int64_t n_frames, int max_it)
{
    int64_t f;
    for (f = 0; f < n_frames; ++f)
        iters[f] = orc_sumprod2(ch + f * nv, vdeg, cdeg, intrlv, nv, nc, nmsg,
                                app + f * nv, max_it);
}

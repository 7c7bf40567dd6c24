// Host side of the batched LDPC BP decoder: graph handles, the batched
// device/host entry points and the legacy single-frame symbols.  The kernels are
// in ldpc_bp_kernels.cuh (instantiated by ldpc_bp_f32.cu / ldpc_bp_f64.cu).
#include "ldpc_bp_host.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

using namespace lsb_bp;

namespace {
constexpr int kCounters = 64;

template <typename A, typename B>
__global__ void convert_kernel(const A *__restrict__ in, B *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = (B)in[i];
}

// Systematic QC encoder (reference ldpc.py:400-460), one CTA per codeword, x held in
// shared memory as [cols][z] bytes.  np.roll(v, -s)[i] = v[(i + s) mod z].
__global__ void qc_encode_kernel(const int *__restrict__ proto, int rows, int cols, int z, int toff,
                                 const uint8_t *__restrict__ info, uint8_t *__restrict__ cw, long long n_frames)
{
    extern __shared__ unsigned char enc_smem[];
    const int Kp = cols - rows;
    uint8_t *x = enc_smem;                 // [cols][z]
    uint8_t *p = x + cols * z;             // [rows][z]
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        for (int i = threadIdx.x; i < Kp * z; i += blockDim.x) x[i] = info[f * (long long)Kp * z + i] & 1;
        __syncthreads();
        for (int e = threadIdx.x; e < rows * z; e += blockDim.x) {       // p_j = sum_k roll(x_k, -proto[j,k])
            const int j = e / z, i = e - j * z;
            unsigned acc = 0;
            for (int k = 0; k < Kp; ++k) {
                const int s = proto[j * cols + k];
                if (s != -1) acc ^= x[k * z + (i + s) % z];
            }
            p[e] = (uint8_t)acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < z; i += blockDim.x) {              // first parity block: roll(sum_j p_j, toff)
            const int src = (i - toff + z) % z;
            unsigned acc = 0;
            for (int j = 0; j < rows; ++j) acc ^= p[j * z + src];
            x[Kp * z + i] = (uint8_t)acc;
        }
        __syncthreads();
        for (int j = 0; j + 1 < rows; ++j) {                            // remaining parity blocks, one row at a time
            const int myk = Kp + j + 1;
            for (int i = threadIdx.x; i < z; i += blockDim.x) {
                unsigned acc = p[j * z + i];
                for (int k = Kp; k < myk; ++k) {
                    const int s = proto[j * cols + k];
                    if (s != -1) acc ^= x[k * z + (i + s) % z];
                }
                x[myk * z + i] = (uint8_t)acc;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < cols * z; i += blockDim.x) cw[f * (long long)cols * z + i] = x[i];
        __syncthreads();
    }
}
}  // namespace

// ------------------------------------------------------------------ host side
struct lsb_ldpc_graph {
    int device = 0;
    int64_t nv = 0, nc = 0, nmsg = 0;
    bool is_qc = false;
    int z = 0, max_dc = 0;
    // generic tables
    int *d_voff = nullptr, *d_coff = nullptr, *d_intrlv = nullptr;
    // QC tables
    QcTables *d_tabs = nullptr;
    int qc_edges = 0;
    // systematic encoder (QC graphs): protograph on the device, inverse offset of the first parity block
    int *d_proto = nullptr;
    int rows = 0, cols = 0, enc_toff = -1;
    // frame-queue counters (ring, one per launch in flight)
    unsigned long long *d_queue = nullptr;
    std::atomic<unsigned> next_queue{0};
    cudaEvent_t queue_ev[kCounters] = {};      // recorded behind the launch that used the slot: its next user waits on it
    // generic-kernel message workspace: one buffer per handle, so generic
    // launches are chained through ws_event (they may sit on different streams)
    std::mutex ws_mu;
    void *d_ws = nullptr;
    size_t ws_bytes = 0;
    cudaEvent_t ws_event = nullptr;
    // cached launch geometry per [fp64?][algo]
    int qc_grid[2][3] = {{0, 0, 0}, {0, 0, 0}};
    int gen_grid[2][3] = {{0, 0, 0}, {0, 0, 0}};
    // host-API staging (lazily sized), one host call at a time per handle
    std::mutex stage_mu;
    struct Stage {
        cudaStream_t stream = nullptr;
        void *d_in = nullptr, *d_app = nullptr, *d_in32 = nullptr, *d_app32 = nullptr;
        uint8_t *d_hard = nullptr;
        uint32_t *d_bits = nullptr;
        int32_t *d_iters = nullptr;
        size_t frames = 0;
        int elem = 0;
    } stage[3];
};

namespace {

int check_device(const lsb_ldpc_graph *g)
{
    int dev = -1;
    LSB_CUDA(cudaGetDevice(&dev));
    LSB_REQUIRE(dev == g->device, "graph was created on device %d, current device is %d",
                g->device, dev);
    return LSB_OK;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    LSB_CUDA(cudaMalloc((void **)dst, src.size() * sizeof(T)));
    LSB_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return LSB_OK;
}

int finish_graph(lsb_ldpc_graph *g, const std::vector<int> &voff, const std::vector<int> &coff,
                 const std::vector<int> &intrlv)
{
    int rc;
    if ((rc = upload(&g->d_voff, voff))) return rc;
    if ((rc = upload(&g->d_coff, coff))) return rc;
    if ((rc = upload(&g->d_intrlv, intrlv))) return rc;
    LSB_CUDA(cudaMalloc((void **)&g->d_queue, kCounters * sizeof(unsigned long long)));
    LSB_CUDA(cudaMemset(g->d_queue, 0, kCounters * sizeof(unsigned long long)));
    return LSB_OK;
}

size_t qc_smem_bytes(const lsb_ldpc_graph *g, size_t elem)
{
    return ((size_t)g->qc_edges * g->z + 2 * (size_t)g->nv) * elem;
}

// the graph runs on the shared-memory QC kernel (else: generic kernel)
bool qc_path(const lsb_ldpc_graph *g, size_t elem)
{
    return g->is_qc && g->max_dc <= kMaxDcQc &&
           qc_smem_bytes(g, elem) + sizeof(QcTables) + 64 <= (size_t)lsb::max_smem_optin();
}

template <typename T>
int decode_dev(const lsb_ldpc_graph *cg, const T *llr, T *app, uint8_t *hard, int32_t *iters,
               int64_t n_frames, int max_it, int algo, double scale, cudaStream_t st, const McParams *mc = nullptr)
{
    lsb_ldpc_graph *g = const_cast<lsb_ldpc_graph *>(cg);
    if (n_frames == 0) return LSB_OK;                  // an empty batch is valid with any pointers
    LSB_REQUIRE(g && (mc || (llr && iters)), "null graph / llr / iters pointer");
    LSB_REQUIRE(n_frames >= 0 && max_it >= 0, "negative n_frames or max_it");
    LSB_REQUIRE(algo >= 0 && algo <= 2, "unknown algo %d", algo);
    int rc = check_device(g);
    if (rc) return rc;

    // Frame-queue counter: a ring of kCounters slots.  The slot's previous user (64 launches ago, possibly on another
    // stream and still running) must be done before the counter is zeroed again.
    const unsigned slot = g->next_queue.fetch_add(1) % kCounters;
    unsigned long long *queue = g->d_queue + slot;
    if (g->queue_ev[slot]) LSB_CUDA(cudaStreamWaitEvent(st, g->queue_ev[slot], 0));
    else LSB_CUDA(cudaEventCreateWithFlags(&g->queue_ev[slot], cudaEventDisableTiming));
    LSB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));

    if (qc_path(g, sizeof(T))) {
        QcLaunch<T> a;
        a.tabs = g->d_tabs;
        a.llr = llr;
        a.app = app;
        a.hard = hard;
        a.iters = iters;
        a.n_frames = n_frames;
        a.max_it = max_it;
        a.algo = algo;
        a.max_dc = g->max_dc;
        a.z = g->z;
        a.scale = (T)scale;
        a.queue = queue;
        a.smem = qc_smem_bytes(g, sizeof(T));
        a.cached_grid = &g->qc_grid[sizeof(T) == 8][algo];
        a.stream = st;
        if (mc) a.mc = *mc;
        rc = launch_qc<T>(a);
        if (!rc) LSB_CUDA(cudaEventRecord(g->queue_ev[slot], st));
        return rc;
    }
    LSB_REQUIRE(!mc, "the fused Monte-Carlo front end needs a QC graph within the QC kernel's limits");
    // generic path: one message workspace per handle, so launches are chained
    // through ws_event (they may sit on different streams)
    int &cached = g->gen_grid[sizeof(T) == 8][0];
    if (cached == 0) cached = generic_blocks_per_sm<T>() * lsb::sm_count();
    GenericLaunch<T> a;
    a.voff = g->d_voff;
    a.coff = g->d_coff;
    a.intrlv = g->d_intrlv;
    a.nv = (int)g->nv;
    a.nc = (int)g->nc;
    a.nmsg = (int)g->nmsg;
    a.llr = llr;
    a.app = app;
    a.hard = hard;
    a.iters = iters;
    a.n_frames = n_frames;
    a.max_it = max_it;
    a.algo = algo;
    a.scale = (T)scale;
    a.queue = queue;
    a.grid = cached < n_frames ? cached : n_frames;
    a.stream = st;
    const size_t need = (size_t)cached * (size_t)(g->nmsg + g->nv) * sizeof(double);
    std::lock_guard<std::mutex> lock(g->ws_mu);
    if (g->ws_bytes < need) {
        LSB_CUDA(cudaDeviceSynchronize());
        if (g->d_ws) cudaFree(g->d_ws);
        g->d_ws = nullptr;
        g->ws_bytes = 0;
        LSB_CUDA(cudaMalloc(&g->d_ws, need));
        g->ws_bytes = need;
    }
    a.ws = (T *)g->d_ws;
    if (!g->ws_event)
        LSB_CUDA(cudaEventCreateWithFlags(&g->ws_event, cudaEventDisableTiming));
    else
        LSB_CUDA(cudaStreamWaitEvent(st, g->ws_event, 0));
    rc = launch_generic<T>(a);
    if (rc) return rc;
    LSB_CUDA(cudaEventRecord(g->ws_event, st));
    LSB_CUDA(cudaEventRecord(g->queue_ev[slot], st));
    return LSB_OK;
}

void free_stage(lsb_ldpc_graph::Stage &s)
{
    cudaFree(s.d_in);
    cudaFree(s.d_app);
    cudaFree(s.d_in32);
    cudaFree(s.d_app32);
    cudaFree(s.d_hard);
    cudaFree(s.d_bits);
    cudaFree(s.d_iters);
    s.d_in = s.d_app = s.d_in32 = s.d_app32 = nullptr;
    s.d_hard = nullptr;
    s.d_bits = nullptr;
    s.d_iters = nullptr;
    s.frames = 0;
}

int ensure_stage(lsb_ldpc_graph *g, lsb_ldpc_graph::Stage &s, size_t frames, int elem,
                 bool need32)
{
    if (!s.stream) LSB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (s.frames >= frames && s.elem >= elem && (!need32 || s.d_in32)) return LSB_OK;
    LSB_CUDA(cudaStreamSynchronize(s.stream));
    free_stage(s);
    const size_t n = frames * (size_t)g->nv;
    LSB_CUDA(cudaMalloc(&s.d_in, n * elem));
    LSB_CUDA(cudaMalloc(&s.d_app, n * elem));
    if (need32) {
        LSB_CUDA(cudaMalloc(&s.d_in32, n * sizeof(float)));
        LSB_CUDA(cudaMalloc(&s.d_app32, n * sizeof(float)));
    }
    LSB_CUDA(cudaMalloc((void **)&s.d_hard, n));
    LSB_CUDA(cudaMalloc((void **)&s.d_bits, frames * (size_t)((g->nv + 31) / 32) * sizeof(uint32_t)));
    LSB_CUDA(cudaMalloc((void **)&s.d_iters, frames * sizeof(int32_t)));
    s.frames = frames;
    s.elem = elem;
    return LSB_OK;
}

// hard decisions [frames][nv] bytes -> [frames][ceil(nv / 32)] words, bit v of a frame = word v / 32, bit v % 32.
// Only behind the generic kernel; the QC kernel packs in its epilogue (McParams::hard_bits) -- a second kernel on the
// chunk's stream queues behind the persistent decode grids of the other two chunks (measured: e2e 2.79 instead of
// 3.19 Gbit/s).
__global__ void pack_bits_kernel(const uint8_t *__restrict__ hard, uint32_t *__restrict__ bits, long long n_frames, int nv)
{
    const int pitch = (nv + 31) / 32, lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_frames * pitch; w += warps) {
        const long long f = w / pitch;
        const int v = (int)(w - f * pitch) * 32 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, v < nv && hard[f * nv + v] != 0);
        if (lane == 0) bits[w] = m;
    }
}

// Host-pointer decode: chunks of frames rotate over three streams so the H2D
// copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1 overlap.
template <typename T>
int decode_host(const lsb_ldpc_graph *cg, const T *llr, T *app, uint8_t *hard, int32_t *iters,
                int64_t n_frames, int max_it, int algo, double scale, bool compute_f32_from_f64,
                uint32_t *bits = nullptr)
{
    lsb_ldpc_graph *g = const_cast<lsb_ldpc_graph *>(cg);
    if (n_frames == 0) return LSB_OK;                  // an empty batch is valid with any pointers
    LSB_REQUIRE(g && llr && iters, "null graph / llr / iters pointer");
    LSB_REQUIRE(n_frames >= 0 && max_it >= 0, "negative n_frames or max_it");
    int rc = check_device(g);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g->stage_mu);   // staging buffers are per handle

    const size_t nv = (size_t)g->nv;
    // chunk: ~32 MiB of LLRs, at least 1 frame, at most n_frames/3 so three stages overlap
    size_t chunk = (32u << 20) / (nv * sizeof(T));
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)(n_frames + 2) / 3) chunk = (size_t)(n_frames + 2) / 3;
    if (chunk < 1) chunk = 1;

    // The first chunks are short (1/8, 1/4, 1/2 of a chunk): the decoder starts after a 4 MB copy instead of a 32 MB one
    // (the pipeline's fill time is all of the difference between this call and the resident-input number).
    int si = 0, ci = 0;
    size_t nf = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += (int64_t)nf, si = (si + 1) % 3, ++ci) {
        lsb_ldpc_graph::Stage &s = g->stage[si];
        size_t want = ci < 3 ? chunk >> (3 - ci) : chunk;
        if (want < 1) want = 1;
        nf = (size_t)((n_frames - f0 < (int64_t)want) ? n_frames - f0 : want);
        if ((rc = ensure_stage(g, s, chunk, sizeof(T), compute_f32_from_f64))) return rc;
        const size_t n = nf * nv;
        const bool bits_in_kernel = bits && algo == LSB_BP_SUMPROD2 && (compute_f32_from_f64 || sizeof(T) == 4) &&
                                    qc_path(g, sizeof(float));
        McParams mcb;
        mcb.hard_bits = s.d_bits;
        LSB_CUDA(cudaMemcpyAsync(s.d_in, llr + (size_t)f0 * nv, n * sizeof(T),
                                 cudaMemcpyHostToDevice, s.stream));
        if (compute_f32_from_f64) {
            const int blocks = lsb::sm_count() * 8;
            convert_kernel<T, float><<<blocks, 256, 0, s.stream>>>((const T *)s.d_in,
                                                                   (float *)s.d_in32, (long long)n);
            lsb::count_launch();
            rc = decode_dev<float>(g, (const float *)s.d_in32, app ? (float *)s.d_app32 : nullptr,
                                   (hard || (bits && !bits_in_kernel)) ? s.d_hard : nullptr, s.d_iters, (int64_t)nf,
                                   max_it, algo, scale, s.stream, bits_in_kernel ? &mcb : nullptr);
            if (!rc && app) {
                convert_kernel<float, T><<<blocks, 256, 0, s.stream>>>((const float *)s.d_app32,
                                                                       (T *)s.d_app, (long long)n);
                lsb::count_launch();
            }
        } else {
            rc = decode_dev<T>(g, (const T *)s.d_in, app ? (T *)s.d_app : nullptr,
                               (hard || (bits && !bits_in_kernel)) ? s.d_hard : nullptr, s.d_iters, (int64_t)nf, max_it,
                               algo, scale, s.stream, bits_in_kernel ? &mcb : nullptr);
        }
        if (rc) return rc;
        if (app)
            LSB_CUDA(cudaMemcpyAsync(app + (size_t)f0 * nv, s.d_app, n * sizeof(T),
                                     cudaMemcpyDeviceToHost, s.stream));
        if (hard)
            LSB_CUDA(cudaMemcpyAsync(hard + (size_t)f0 * nv, s.d_hard, n, cudaMemcpyDeviceToHost,
                                     s.stream));
        if (bits) {
            const size_t pitch = (nv + 31) / 32;
            if (!bits_in_kernel) {
                pack_bits_kernel<<<lsb::sm_count() * 8, 256, 0, s.stream>>>(s.d_hard, s.d_bits, (long long)nf, (int)nv);
                lsb::count_launch();
                LSB_CUDA(cudaGetLastError());
            }
            LSB_CUDA(cudaMemcpyAsync(bits + (size_t)f0 * pitch, s.d_bits, nf * pitch * sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, s.stream));
        }
        LSB_CUDA(cudaMemcpyAsync(iters + f0, s.d_iters, nf * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, s.stream));
    }
    for (int i = 0; i < 3; ++i)
        if (g->stage[i].stream) LSB_CUDA(cudaStreamSynchronize(g->stage[i].stream));
    return LSB_OK;
}

}  // namespace

extern "C" {

int lsb_ldpc_mc_dev_f32(const lsb_ldpc_graph *g, const uint8_t *codewords, int64_t n_codewords, double sigma2,
                        uint64_t seed, uint64_t frame0, int64_t n_frames, int max_it, int algo, double scale,
                        int32_t *bit_errors, int32_t *iters, float *llr_out, uint64_t *counters, void *stream)
{
    if (n_frames == 0) return LSB_OK;
    LSB_REQUIRE(g && codewords && n_codewords > 0, "null graph / codewords");
    LSB_REQUIRE(sigma2 > 0.0, "sigma2 must be positive");
    LSB_REQUIRE(n_frames >= 0 && max_it >= 0, "negative n_frames or max_it");
    LSB_REQUIRE(algo >= 0 && algo <= 2, "unknown algo %d", algo);
    McParams mc;
    mc.cw = codewords;
    mc.n_cw = n_codewords;
    mc.llr_scale = (float)(2.0 / sigma2);
    mc.sigma = (float)std::sqrt(sigma2);
    mc.seed = seed;
    mc.frame0 = frame0;
    mc.bit_errors = bit_errors;
    mc.llr_out = llr_out;
    mc.counters = (unsigned long long *)counters;
    return decode_dev<float>(g, nullptr, nullptr, nullptr, iters, n_frames, max_it, algo, scale, (cudaStream_t)stream, &mc);
}

int lsb_ldpc_graph_create(const int64_t *vdeg, const int64_t *cdeg, const int64_t *intrlv,
                          int64_t nv, int64_t nc, int64_t nmsg, lsb_ldpc_graph **out)
{
    LSB_REQUIRE(vdeg && cdeg && intrlv && out, "null pointer");
    LSB_REQUIRE(nv > 0 && nc > 0 && nmsg > 0 && nmsg < (1ll << 30), "bad graph sizes");
    std::vector<int> voff(nv + 1, 0), coff(nc + 1, 0), il(nmsg);
    int max_dc = 0;
    for (int64_t v = 0; v < nv; ++v) {
        LSB_REQUIRE(vdeg[v] >= 0, "negative variable degree");
        voff[v + 1] = voff[v] + (int)vdeg[v];
    }
    for (int64_t c = 0; c < nc; ++c) {
        LSB_REQUIRE(cdeg[c] >= 2 && cdeg[c] <= kMaxDc, "check degree %lld outside [2, %d]",
                    (long long)cdeg[c], kMaxDc);
        coff[c + 1] = coff[c] + (int)cdeg[c];
        if (cdeg[c] > max_dc) max_dc = (int)cdeg[c];
    }
    LSB_REQUIRE(voff[nv] == nmsg && coff[nc] == nmsg, "sum(vdeg)=%d sum(cdeg)=%d Nmsg=%lld differ",
                voff[nv], coff[nc], (long long)nmsg);
    for (int64_t e = 0; e < nmsg; ++e) {
        LSB_REQUIRE(intrlv[e] >= 0 && intrlv[e] < nmsg, "intrlv[%lld] out of range", (long long)e);
        il[e] = (int)intrlv[e];
    }
    lsb_ldpc_graph *g = new (std::nothrow) lsb_ldpc_graph();
    if (!g) return LSB_ERR_NOMEM;
    g->nv = nv;
    g->nc = nc;
    g->nmsg = nmsg;
    g->max_dc = max_dc;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    int rc = finish_graph(g, voff, coff, il);
    if (rc) {
        lsb_ldpc_graph_destroy(g);
        return rc;
    }
    *out = g;
    return LSB_OK;
}

int lsb_ldpc_graph_create_qc(const int32_t *proto, int rows, int cols, int z,
                             lsb_ldpc_graph **out)
{
    LSB_REQUIRE(proto && out, "null pointer");
    LSB_REQUIRE(rows > 0 && rows <= kMaxRows && cols > rows && cols <= kMaxCols,
                "protograph %dx%d outside supported range", rows, cols);
    LSB_REQUIRE(z > 0 && (int64_t)cols * z < 65536, "expansion factor %d out of range", z);

    // Expand exactly as ldpc.py:303-396 does: check (r,k) <-> variable c*z+(k+shift)%z,
    // check ports in increasing column, variable ports in increasing row.
    QcTables t;
    memset(&t, 0, sizeof(t));
    t.z = z;
    t.rows = rows;
    t.cols = cols;
    t.nv = cols * z;
    t.nc = rows * z;
    t.magic = (unsigned)(((1ull << 32) + (unsigned)z - 1) / (unsigned)z);
    std::vector<int> edge_id(rows * cols, -1);
    int ne = 0;
    for (int r = 0; r < rows; ++r) {
        t.row_ptr[r] = (short)ne;
        for (int c = 0; c < cols; ++c)
            if (proto[r * cols + c] != -1) {
                LSB_REQUIRE(proto[r * cols + c] >= 0, "negative shift in protograph");
                LSB_REQUIRE(ne < kMaxEdges, "more than %d protograph entries", kMaxEdges);
                edge_id[r * cols + c] = ne++;
            }
        const int dc = ne - t.row_ptr[r];
        LSB_REQUIRE(dc >= 2 && dc <= kMaxDc, "row %d has degree %d outside [2, %d]", r, dc, kMaxDc);
    }
    t.row_ptr[rows] = (short)ne;
    t.n_edges = ne;
    // stable sort of rows by degree
    {
        std::vector<int> order(rows);
        for (int r = 0; r < rows; ++r) order[r] = r;
        for (int i = 1; i < rows; ++i) {   // insertion sort keeps equal degrees in order
            int r = order[i], d = t.row_ptr[r + 1] - t.row_ptr[r], j = i - 1;
            while (j >= 0 && (t.row_ptr[order[j] + 1] - t.row_ptr[order[j]]) > d) {
                order[j + 1] = order[j];
                --j;
            }
            order[j + 1] = r;
        }
        for (int r = 0; r < rows; ++r) t.row_order[r] = (short)order[r];
    }
    int nt = 0, max_dc = 0;
    for (int c = 0; c < cols; ++c) {
        t.col_ptr[c] = (short)nt;
        for (int r = 0; r < rows; ++r) {
            const int e = edge_id[r * cols + c];
            if (e < 0) continue;
            const int s = proto[r * cols + c] % z;
            t.col_entry[nt] = make_int2(e * z - s, s);
            ++nt;
        }
    }
    t.col_ptr[cols] = (short)nt;

    // generic tables in the reference's numbering (also what the façade exposes)
    const int64_t nv = t.nv, nc = t.nc, nmsg = (int64_t)ne * z;
    std::vector<int> voff(nv + 1, 0), coff(nc + 1, 0), il(nmsg);
    for (int r = 0; r < rows; ++r) {
        const int dc = t.row_ptr[r + 1] - t.row_ptr[r];
        if (dc > max_dc) max_dc = dc;
        for (int k = 0; k < z; ++k) coff[r * z + k + 1] = dc;
    }
    for (int64_t c = 0; c < nc; ++c) coff[c + 1] += coff[c];
    for (int c = 0; c < cols; ++c)
        for (int j = 0; j < z; ++j) voff[c * z + j + 1] = t.col_ptr[c + 1] - t.col_ptr[c];
    for (int64_t v = 0; v < nv; ++v) voff[v + 1] += voff[v];
    for (int c = 0; c < cols; ++c) {
        int port = 0;
        for (int r = 0; r < rows; ++r) {
            const int e = edge_id[r * cols + c];
            if (e < 0) continue;
            const int s = proto[r * cols + c] % z;
            const int cport = e - t.row_ptr[r];
            for (int k = 0; k < z; ++k) {
                const int v = c * z + (k + s) % z;
                il[voff[v] + port] = coff[r * z + k] + cport;
            }
            ++port;
        }
    }

    lsb_ldpc_graph *g = new (std::nothrow) lsb_ldpc_graph();
    if (!g) return LSB_ERR_NOMEM;
    g->nv = nv;
    g->nc = nc;
    g->nmsg = nmsg;
    g->is_qc = true;
    g->z = z;
    g->qc_edges = ne;
    g->rows = rows;
    g->cols = cols;
    {   // encoder: the offsets in column Kp must add to a single offset (ldpc.py:435-446)
        const int Kp = cols - rows;
        std::vector<int> cnt(z, 0);
        for (int r = 0; r < rows; ++r)
            if (proto[r * cols + Kp] != -1) cnt[proto[r * cols + Kp] % z] ^= 1;
        int n_odd = 0, off = -1;
        for (int i = 0; i < z; ++i)
            if (cnt[i]) { ++n_odd; off = i; }
        g->enc_toff = n_odd == 1 ? off : -1;
    }
    g->max_dc = max_dc;
    if (cudaGetDevice(&g->device) != cudaSuccess) {
        lsb::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        delete g;
        return LSB_ERR_NO_DEVICE;
    }
    int rc = finish_graph(g, voff, coff, il);
    if (!rc) {
        cudaError_t e = cudaMalloc((void **)&g->d_tabs, sizeof(QcTables));
        if (e == cudaSuccess)
            e = cudaMemcpy(g->d_tabs, &t, sizeof(QcTables), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc((void **)&g->d_proto, sizeof(int) * rows * cols);
        if (e == cudaSuccess) e = cudaMemcpy(g->d_proto, proto, sizeof(int) * rows * cols, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            lsb::set_error("uploading QC tables: %s", cudaGetErrorString(e));
            rc = LSB_ERR_CUDA;
        }
    }
    if (rc) {
        lsb_ldpc_graph_destroy(g);
        return rc;
    }
    *out = g;
    return LSB_OK;
}

void lsb_ldpc_graph_destroy(lsb_ldpc_graph *g)
{
    if (!g) return;
    for (auto &s : g->stage) {
        if (s.stream) {
            cudaStreamSynchronize(s.stream);
            free_stage(s);
            cudaStreamDestroy(s.stream);
        }
    }
    cudaFree(g->d_voff);
    cudaFree(g->d_coff);
    cudaFree(g->d_intrlv);
    cudaFree(g->d_tabs);
    cudaFree(g->d_proto);
    cudaFree(g->d_queue);
    cudaFree(g->d_ws);
    if (g->ws_event) cudaEventDestroy(g->ws_event);
    for (cudaEvent_t e : g->queue_ev)
        if (e) cudaEventDestroy(e);
    delete g;
}

int lsb_ldpc_encode_dev(const lsb_ldpc_graph *g, const uint8_t *info, uint8_t *codewords, int64_t n_frames,
                        void *stream)
{
    if (n_frames == 0) return LSB_OK;                  // an empty batch is valid with any pointers
    LSB_REQUIRE(g && info && codewords && n_frames > 0, "null pointer or negative frame count");
    LSB_REQUIRE(g->is_qc && g->d_proto, "encoding needs a QC graph (lsb_ldpc_graph_create_qc)");
    LSB_REQUIRE(g->enc_toff >= 0, "The offsets in colum Kp+1 of proto do not add to a single offset");
    int rc = check_device(g);
    if (rc) return rc;
    const size_t smem = (size_t)(g->cols + g->rows) * g->z;
    LSB_REQUIRE(smem <= 48 * 1024, "code too large for the encoder kernel");
    long long grid = (long long)lsb::sm_count() * 8;
    if (grid > n_frames) grid = n_frames;
    qc_encode_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(g->d_proto, g->rows, g->cols, g->z,
                                                                        g->enc_toff, info, codewords,
                                                                        (long long)n_frames);
    lsb::count_launch();
    LSB_CUDA(cudaGetLastError());
    return LSB_OK;
}

int lsb_ldpc_graph_info(const lsb_ldpc_graph *g, int64_t info[5])
{
    LSB_REQUIRE(g && info, "null pointer");
    info[0] = g->nv;
    info[1] = g->nc;
    info[2] = g->nmsg;
    info[3] = g->is_qc ? 1 : 0;
    info[4] = g->z;
    return LSB_OK;
}

int lsb_ldpc_decode_dev_f32(const lsb_ldpc_graph *g, const float *llr, float *app, uint8_t *hard,
                            int32_t *iters, int64_t n_frames, int max_it, int algo, double scale,
                            void *stream)
{
    return decode_dev<float>(g, llr, app, hard, iters, n_frames, max_it, algo, scale,
                             (cudaStream_t)stream);
}

int lsb_ldpc_decode_dev_f64(const lsb_ldpc_graph *g, const double *llr, double *app,
                            uint8_t *hard, int32_t *iters, int64_t n_frames, int max_it, int algo,
                            double scale, void *stream)
{
    return decode_dev<double>(g, llr, app, hard, iters, n_frames, max_it, algo, scale,
                              (cudaStream_t)stream);
}

int lsb_ldpc_decode_host_bits_f32(const lsb_ldpc_graph *g, const float *llr, uint32_t *bits, int32_t *iters,
                                  int64_t n_frames, int max_it, int algo, double scale)
{
    LSB_REQUIRE(bits || n_frames == 0, "null bits pointer");
    return decode_host<float>(g, llr, nullptr, nullptr, iters, n_frames, max_it, algo, scale, false, bits);
}

int lsb_ldpc_decode_host_f32(const lsb_ldpc_graph *g, const float *llr, float *app, uint8_t *hard,
                             int32_t *iters, int64_t n_frames, int max_it, int algo, double scale)
{
    return decode_host<float>(g, llr, app, hard, iters, n_frames, max_it, algo, scale, false);
}

int lsb_ldpc_decode_host_f64(const lsb_ldpc_graph *g, const double *llr, double *app,
                             uint8_t *hard, int32_t *iters, int64_t n_frames, int max_it, int algo,
                             double scale, int compute_f64)
{
    return decode_host<double>(g, llr, app, hard, iters, n_frames, max_it, algo, scale,
                               !compute_f64);
}

// ------------------------------------------------------------ legacy symbols
// One frame, host pointers, arbitrary graph (c_ldpc.c:32,138,339).  The graph
// arrays are hashed and the device copy cached, since the reference's loader
// passes the same arrays on every call (ldpc.py:464-485).
namespace {

struct LegacyCache {
    std::mutex mu;
    std::unordered_map<uint64_t, lsb_ldpc_graph *> graphs;
};
LegacyCache &legacy_cache()
{
    static LegacyCache *c = new LegacyCache();   // leaked on purpose: outlives atexit order
    return *c;
}

uint64_t fnv(uint64_t h, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}

int legacy_decode(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg,
                  double *app, int max_it, int algo, double scale)
{
    if (!ch || !vdeg || !cdeg || !intrlv || !app || Nv <= 0 || Nc <= 0 || Nmsg <= 0) {
        lsb::set_error("legacy decode: bad arguments");
        return -1;
    }
    if (max_it < 0) max_it = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        lsb::set_error("legacy decode: no CUDA device");
        return -1;
    }
    uint64_t h = fnv(1469598103934665603ull, &dev, sizeof(dev));
    h = fnv(h, vdeg, sizeof(long) * Nv);
    h = fnv(h, cdeg, sizeof(long) * Nc);
    h = fnv(h, intrlv, sizeof(long) * Nmsg);
    lsb_ldpc_graph *g = nullptr;
    {
        LegacyCache &c = legacy_cache();
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.graphs.find(h);
        if (it != c.graphs.end()) {
            g = it->second;
        } else {
            std::vector<int64_t> v(vdeg, vdeg + Nv), cd(cdeg, cdeg + Nc), il(intrlv, intrlv + Nmsg);
            if (lsb_ldpc_graph_create(v.data(), cd.data(), il.data(), Nv, Nc, Nmsg, &g) != LSB_OK)
                return -1;
            c.graphs[h] = g;
        }
    }
    int32_t iters = -1;
    if (lsb_ldpc_decode_host_f64(g, ch, app, nullptr, &iters, 1, max_it, algo, scale, 1) != LSB_OK)
        return -1;
    if (max_it == 0) memset(app, 0, sizeof(double) * Nv);   // reference leaves app untouched
    return iters;
}

}  // namespace

int sumprod(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg,
            double *app, int max_itcount)
{
    return legacy_decode(ch, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, app, max_itcount, LSB_BP_SUMPROD,
                         1.0);
}

int sumprod2(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg,
             double *app, int max_itcount)
{
    return legacy_decode(ch, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, app, max_itcount, LSB_BP_SUMPROD2,
                         1.0);
}

int minsum(double *ch, long *vdeg, long *cdeg, long *intrlv, int Nv, int Nc, int Nmsg,
           double *app, double correction_factor, int max_itcount)
{
    return legacy_decode(ch, vdeg, cdeg, intrlv, Nv, Nc, Nmsg, app, max_itcount, LSB_BP_MINSUM,
                         correction_factor);
}

double Lxor(double L1, double L2, int corr_flag)
{
    double *d = nullptr, out = NAN;
    if (cudaMalloc((void **)&d, sizeof(double)) != cudaSuccess) return NAN;
    launch_lxor(L1, L2, corr_flag, d);
    cudaMemcpy(&out, d, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return out;
}

double Lxfb(double *L, long dc, int corr_flag)
{
    if (!L || dc < 2 || dc > kMaxDc) {
        lsb::set_error("Lxfb: dc=%ld outside [2, %d]", dc, kMaxDc);
        return NAN;
    }
    double *d = nullptr, out = NAN;
    if (cudaMalloc((void **)&d, sizeof(double) * (dc + 1)) != cudaSuccess) return NAN;
    cudaMemcpy(d, L, sizeof(double) * dc, cudaMemcpyHostToDevice);
    launch_lxfb(d, (int)dc, corr_flag, d + dc);
    cudaMemcpy(L, d, sizeof(double) * dc, cudaMemcpyDeviceToHost);
    cudaMemcpy(&out, d + dc, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return out;
}

}  // extern "C"

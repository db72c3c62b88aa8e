// ofb_api.cu -- context object and the extern "C" surface declared in include/ofb200.h.
#include "host_extract.hpp"
#include "ofb_common.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace ofb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace ofb

#define OFB_LANES 6 /* streams available to the batched host entry point; it uses c->lanes of them */

// One growable device workspace per context; carved by offset for each call.  Growing it frees the
// old block (cudaFree synchronises the device), so steady-state calls allocate nothing -- unlike
// the reference, which does 30 cudaMalloc/cudaFree pairs per level (OptFlowGpu.cu:1105-1124 etc.).
struct ofb_ctx {
    int device = 0;
    int sm_count = 0;
    int lanes = 3;      // sub-batches in flight in ofb_flow_pairs_host (developer override: OFB_E2E_LANES)
    int sub_pairs = 0;  // pairs per sub-batch there; 0 = chosen from the batch size (developer override: OFB_E2E_SUB)
    int solve_fast = 0; // OFB_SOLVE_*: which 2x2 solve the fused level kernel runs (ofb_ctx_set_solve)
    cudaStream_t stream = nullptr; // used by the synchronous host-pointer entry points
    cudaStream_t lane_stream[OFB_LANES] = {}; // the batched host entry point pipelines sub-batches over these
    uint8_t *ws = nullptr;
    size_t ws_bytes = 0;
    // The workspace serves ONE stream at a time: every asynchronous use records ws_event on its stream, and a later user
    // on another stream waits for it first (ws_acquire / ws_release).
    cudaEvent_t ws_event = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_busy = false;
    unsigned long long launches = 0;
    // optional in-situ timing: CUDA events around each fused-LK launch (tag = level) and around the
    // pyramid build (tag = OFB_PROFILE_PYRAMID), recorded on the stream the work is launched on
    struct ProfRec {
        int tag;
        cudaEvent_t a, b;
    };
    double *bil_lut = nullptr; // range-weight table of the bilateral pre-filter (device)
    // 3-channel host images: channel 0 extracted by host threads into pinned planar staging buffers, one pair of buffers
    // per lane (ofb_ctx_set_host_threads; 0 = all three channels are uploaded and the device drops two)
    int host_threads = 0;
    ofb::HostPool *pool = nullptr;
    uint8_t *stage = nullptr;
    size_t stage_bytes = 0;
    cudaEvent_t stage_ev[OFB_LANES] = {};
    bool prof_on = false;
    std::vector<ProfRec> prof;
};

namespace ofb {

// accessors for the other translation units (strips.cu): ofb_ctx stays private to this file
int ctx_device(const ofb_ctx *c) { return c->device; }
int ctx_sm_count(const ofb_ctx *c) { return c->sm_count; }
int ctx_solve_fast(const ofb_ctx *c) { return c->solve_fast; }
unsigned long long *ctx_launch_counter(ofb_ctx *c) { return &c->launches; }

// Growing frees the old block after a device synchronisation: CUDA graphs captured from earlier calls hold pointers
// into it and must be re-captured (ofb_ctx_reserve_pairs sizes the workspace up front so that this never happens).
static int ws_reserve(ofb_ctx *c, size_t bytes)
{
    if (bytes <= c->ws_bytes) return OFB_OK;
    if (c->ws) {
        OFB_CUDA_TRY(cudaDeviceSynchronize());
        OFB_CUDA_TRY(cudaFree(c->ws));
        c->ws = nullptr;
        c->ws_bytes = 0;
    }
    const size_t want = align_up(bytes + bytes / 8, 1 << 20);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->ws), want);
    if (e != cudaSuccess) {
        set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return OFB_ERR_NOMEM;
    }
    c->ws_bytes = want;
    return OFB_OK;
}

// Orders the uses of the shared workspace across streams (see ofb_ctx::ws_event).
// A stream that is being captured into a CUDA graph takes no part: a capture cannot depend on eager work of another
// stream, and an event recorded inside a capture belongs to the graph.  Replays of such a graph are ordered against the
// context's other users by their caller (one context per concurrently used stream is the simple rule).
static bool ws_capturing(cudaStream_t st)
{
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone;
}
static int ws_acquire(ofb_ctx *c, cudaStream_t st)
{
    if (c->ws_busy && c->ws_stream != st && !ws_capturing(st)) OFB_CUDA_TRY(cudaStreamWaitEvent(st, c->ws_event, 0));
    return OFB_OK;
}
static int ws_release(ofb_ctx *c, cudaStream_t st)
{
    if (ws_capturing(st)) return OFB_OK;
    OFB_CUDA_TRY(cudaEventRecord(c->ws_event, st));
    c->ws_stream = st;
    c->ws_busy = true;
    return OFB_OK;
}

struct Carver {
    size_t off = 0;
    size_t take(size_t bytes)
    {
        const size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

static int check_params(const ofb_params *p)
{
    if (!p) {
        set_error("params is NULL");
        return OFB_ERR_INVALID;
    }
    if (p->levels < 1 || p->levels > OFB_MAX_LEVELS) {
        set_error("levels %d outside 1..%d", p->levels, OFB_MAX_LEVELS);
        return OFB_ERR_INVALID;
    }
    if (p->w < 1 || p->h < 1 || (p->w >> (p->levels - 1)) < 2 || (p->h >> (p->levels - 1)) < 2) {
        set_error("frame %dx%d too small for %d levels (coarsest level must be at least 2x2)", p->w, p->h, p->levels);
        return OFB_ERR_INVALID;
    }
    if (p->win < 3 || p->win > OFB_MAX_WINDOW || !(p->win & 1)) {
        set_error("window %d not supported (odd 3..%d)", p->win, OFB_MAX_WINDOW);
        return OFB_ERR_UNSUPPORTED;
    }
    if (p->warp_mode < OFB_WARP_AS_WRITTEN || p->warp_mode > OFB_WARP_BILINEAR) {
        set_error("unknown warp mode %d", p->warp_mode);
        return OFB_ERR_INVALID;
    }
    if (p->n_pairs < 1) {
        set_error("n_pairs %d < 1", p->n_pairs);
        return OFB_ERR_INVALID;
    }
    return OFB_OK;
}

// Pyramids of both frames + coarse-to-fine LK, everything device resident.  `base` is the
// context workspace region reserved for this call.
struct PairPlan {
    int L = 0;
    int w[OFB_MAX_LEVELS], h[OFB_MAX_LEVELS];
    size_t pitch[OFB_MAX_LEVELS], istride[OFB_MAX_LEVELS];
    size_t off_prev[OFB_MAX_LEVELS], off_next[OFB_MAX_LEVELS], off_cum[OFB_MAX_LEVELS];
    size_t bytes = 0;
};

// Two pyramid steps per launch (pyramid.cu: pyr_roll_kernel<true>); OFB_PYR_FUSE=0 builds level by level (A/B measurements).
static bool pyr_fuse_enabled()
{
    static const bool on = [] {
        const char *e = getenv("OFB_PYR_FUSE");
        return !(e && e[0] == '0');
    }();
    return on;
}

static void plan_pairs(const ofb_params *p, PairPlan *pl, Carver *cv)
{
    pl->L = p->levels;
    for (int k = 0; k < p->levels; k++) {
        pl->w[k] = p->w >> k;
        pl->h[k] = p->h >> k;
        pl->pitch[k] = align_up((size_t)pl->w[k], 64);
        pl->istride[k] = pl->pitch[k] * (size_t)pl->h[k];
        pl->off_prev[k] = pl->off_next[k] = pl->off_cum[k] = (size_t)-1;
        if (k >= 1) {
            pl->off_prev[k] = cv->take(pl->istride[k] * p->n_pairs);
            pl->off_next[k] = cv->take(pl->istride[k] * p->n_pairs);
        }
        if (k >= 1 && k <= p->levels - 2) pl->off_cum[k] = cv->take((size_t)pl->w[k] * pl->h[k] * 8 * p->n_pairs);
    }
    pl->bytes = cv->off;
}

static int prof_begin(ofb_ctx *c, int tag, cudaStream_t st)
{
    if (!c->prof_on) return OFB_OK;
    ofb_ctx::ProfRec r{tag, nullptr, nullptr};
    OFB_CUDA_TRY(cudaEventCreate(&r.a));
    OFB_CUDA_TRY(cudaEventCreate(&r.b));
    OFB_CUDA_TRY(cudaEventRecord(r.a, st));
    c->prof.push_back(r);
    return OFB_OK;
}
static int prof_end(ofb_ctx *c, cudaStream_t st)
{
    if (!c->prof_on || c->prof.empty()) return OFB_OK;
    OFB_CUDA_TRY(cudaEventRecord(c->prof.back().b, st));
    return OFB_OK;
}

static int run_pairs_device(ofb_ctx *c, const ofb_params *p, const PairPlan &pl, uint8_t *base, const uint8_t *prev0,
                            const uint8_t *next0, size_t pitch0, size_t istride0, float *const *flow_levels,
                            float *total_flow, cudaStream_t st)
{
    const int L = p->levels, n = p->n_pairs;
    const uint8_t *pp[OFB_MAX_LEVELS], *pn[OFB_MAX_LEVELS];
    size_t pitch[OFB_MAX_LEVELS], istr[OFB_MAX_LEVELS];
    pp[0] = prev0;
    pn[0] = next0;
    pitch[0] = pitch0;
    istr[0] = istride0;
    if (L > 1) {
        int rc = prof_begin(c, OFB_PROFILE_PYRAMID, st);
        if (rc) return rc;
    }
    for (int k = 1; k < L; k++) {
        pp[k] = base + pl.off_prev[k];
        pn[k] = base + pl.off_next[k];
        pitch[k] = pl.pitch[k];
        istr[k] = pl.istride[k];
    }
    const bool one = 2 * n <= 65535; // grid.z limit: both frames of every pair in one launch
    for (int k = 1; k < L;) {
        // two pyramid steps per launch where two are left (level k is then never read back), else one
        const bool two = pyr_fuse_enabled() && k + 1 < L;
        for (int f = 0; f < (one ? 1 : 2); f++) {
            const uint8_t *const *src = f ? pn : pp;
            int rc = launch_pyr_roll(src[k - 1], pitch[k - 1], istr[k - 1], pl.w[k - 1], pl.h[k - 1], const_cast<uint8_t *>(src[k]),
                                     pitch[k], istr[k], two ? const_cast<uint8_t *>(src[k + 1]) : nullptr, two ? pitch[k + 1] : 0,
                                     two ? istr[k + 1] : 0, n, st, &c->launches, one ? pn[k - 1] : nullptr,
                                     one ? const_cast<uint8_t *>(pn[k]) : nullptr,
                                     one && two ? const_cast<uint8_t *>(pn[k + 1]) : nullptr, c->sm_count);
            if (rc) return rc;
        }
        k += two ? 2 : 1;
    }
    if (L > 1) {
        int rc = prof_end(c, st);
        if (rc) return rc;
    }
    // Which warped levels materialise their cumulative flow.  Level 1 need not when level 0 composes it on the fly from
    // flow_1 and cum_2 (lk_level_kernel<..., COMP>): 8 bytes per pixel less to write at level 1, whose kernel is bound by
    // exactly those stores.  Composition needs even sizes (no clamped coarse indices), TMA-able rows and a per-pixel warp
    // mode.  Only level 0 composes, and only when it does not itself write the total flow: a reader that also writes its
    // cumulative flow pays more for the composition than the (four times smaller) coarser level saves -- measured on the
    // 4-level 4K configuration: level 1 +8 %, level 2 -20 %, a net loss.
    bool mat[OFB_MAX_LEVELS] = {}, comp[OFB_MAX_LEVELS] = {};
    for (int k = 1; k < L; k++) mat[k] = true;
    static const bool allow_compose = [] {
        const char *e = getenv("OFB_NO_COMPOSE"); // developer switch for A/B measurements
        return !(e && e[0] == '1');
    }();
    if (L >= 3 && !total_flow && allow_compose && p->warp_mode != OFB_WARP_AS_WRITTEN) {
        const bool even = !((pl.w[0] | pl.h[0] | pl.w[1] | pl.h[1] | pl.w[2]) & 1);
        const float *c2 = (2 == L - 1) ? flow_levels[2] : reinterpret_cast<const float *>(base + pl.off_cum[2]);
        const bool aligned = !((reinterpret_cast<uintptr_t>(flow_levels[1]) | reinterpret_cast<uintptr_t>(c2)) & 15);
        if (even && aligned) {
            comp[0] = true;
            mat[1] = false;
        }
    }
    for (int k = L - 1; k >= 0; k--) {
        LkLevelArgs a{};
        a.prev = pp[k];
        a.next = pn[k];
        a.pitch = pitch[k];
        a.image_stride = istr[k];
        a.w = pl.w[k];
        a.h_local = pl.h[k];
        a.y_off = 0;
        a.h_global = pl.h[k];
        a.out_y0 = 0;
        a.out_y1 = pl.h[k];
        a.n_pairs = n;
        a.win = p->win;
        a.warp_mode = p->warp_mode;
        a.flow_scale = p->flow_scale;
        a.flow_out = flow_levels[k];
        a.flow_pair_stride = (size_t)pl.w[k] * pl.h[k];
        a.sm_count = c->sm_count;
        a.solve_fast = c->solve_fast;
        if (k < L - 1) {
            // cum_{k+1}: the coarsest level's cumulative flow is its residual flow; a level that did not materialise its
            // own hands over its residual flow together with its parent's cumulative flow
            a.cum_in = (k + 1 == L - 1 || comp[k]) ? flow_levels[k + 1] : reinterpret_cast<const float *>(base + pl.off_cum[k + 1]);
            a.cum_w = pl.w[k + 1];
            a.cum_h_global = pl.h[k + 1];
            a.cum_y_off = 0;
            a.cum_h_local = pl.h[k + 1];
            a.cum_pair_stride = (size_t)pl.w[k + 1] * pl.h[k + 1];
            if (comp[k]) {
                a.cum2_in = (k + 2 == L - 1) ? flow_levels[k + 2] : reinterpret_cast<const float *>(base + pl.off_cum[k + 2]);
                a.cum2_pair_stride = (size_t)pl.w[k + 2] * pl.h[k + 2];
            }
        }
        if (k == 0) a.cum_out = total_flow;
        else if (k <= L - 2 && mat[k]) a.cum_out = reinterpret_cast<float *>(base + pl.off_cum[k]);
        int rc = prof_begin(c, k, st);
        if (rc) return rc;
        rc = launch_lk_level(a, st, &c->launches);
        if (rc) return rc;
        rc = prof_end(c, st);
        if (rc) return rc;
    }
    return OFB_OK;
}

} // namespace ofb

using namespace ofb;

#define OFB_CHECK_CTX(c)                     \
    do {                                     \
        if (!(c)) {                          \
            set_error("context is NULL");    \
            return OFB_ERR_INVALID;          \
        }                                    \
    } while (0)
#define OFB_GUARD(c)            \
    DeviceGuard _guard((c)->device); \
    if (!_guard.ok) return OFB_ERR_CUDA

extern "C" {

const char *ofb_last_error(void) { return g_err; }
int ofb_version(void) { return 100; }

int ofb_ctx_create(int device, ofb_ctx **out)
{
    if (!out) {
        set_error("out is NULL");
        return OFB_ERR_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return OFB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d outside 0..%d", device, ndev - 1);
        return OFB_ERR_INVALID;
    }
    DeviceGuard g(device);
    if (!g.ok) return OFB_ERR_CUDA;
    cudaDeviceProp prop;
    OFB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return OFB_ERR_UNSUPPORTED;
    }
    ofb_ctx *c = new (std::nothrow) ofb_ctx();
    if (!c) {
        set_error("out of host memory");
        return OFB_ERR_NOMEM;
    }
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("OFB_E2E_LANES")) { // developer overrides for experiments
        const int v = atoi(e);
        if (v >= 1 && v <= OFB_LANES) c->lanes = v;
    }
    if (const char *e = getenv("OFB_E2E_SUB")) c->sub_pairs = atoi(e) > 0 ? atoi(e) : 0;
    if (const char *e = getenv("OFB_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 64) c->host_threads = v;
    }
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
        delete c;
        return OFB_ERR_CUDA;
    }
    e = cudaEventCreateWithFlags(&c->ws_event, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        set_error("cudaEventCreate failed: %s", cudaGetErrorString(e));
        cudaStreamDestroy(c->stream);
        delete c;
        return OFB_ERR_CUDA;
    }
    for (int l = 0; l < OFB_LANES; l++) {
        e = cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
            delete c;
            return OFB_ERR_CUDA;
        }
    }
    *out = c;
    return OFB_OK;
}

int ofb_ctx_destroy(ofb_ctx *c)
{
    OFB_CHECK_CTX(c);
    {
        DeviceGuard g(c->device);
        if (g.ok) {
            cudaDeviceSynchronize();
            for (auto &r : c->prof) {
                cudaEventDestroy(r.a);
                cudaEventDestroy(r.b);
            }
            if (c->ws) cudaFree(c->ws);
            if (c->bil_lut) cudaFree(c->bil_lut);
            if (c->ws_event) cudaEventDestroy(c->ws_event);
            if (c->stream) cudaStreamDestroy(c->stream);
            for (int l = 0; l < OFB_LANES; l++) {
                if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]);
                if (c->stage_ev[l]) cudaEventDestroy(c->stage_ev[l]);
            }
            if (c->stage) cudaFreeHost(c->stage);
        }
    }
    delete c->pool;
    delete c;
    return OFB_OK;
}

int ofb_ctx_set_host_threads(ofb_ctx *c, int threads)
{
    OFB_CHECK_CTX(c);
    if (threads < 0 || threads > 64) {
        set_error("host threads %d outside 0..64", threads);
        return OFB_ERR_INVALID;
    }
    if (c->pool && c->pool->threads() != threads) {
        delete c->pool;
        c->pool = nullptr;
    }
    c->host_threads = threads;
    return OFB_OK;
}

int ofb_ctx_get_host_threads(const ofb_ctx *c, int *threads)
{
    OFB_CHECK_CTX(c);
    if (threads) *threads = c->host_threads;
    return OFB_OK;
}

int ofb_c3_extract_host(const unsigned char *src_c3, unsigned char *dst_planar, size_t n_pixels, int threads)
{
    if (!src_c3 || !dst_planar || threads < 1 || threads > 64) {
        set_error("c3_extract_host: bad arguments");
        return OFB_ERR_INVALID;
    }
    try {
        ofb::HostPool pool(threads);
        // (the two-set helper with one set: the second set is empty)
        const size_t half = n_pixels / 2;
        ofb::c3_extract_pair_sets(&pool, src_c3, src_c3 + 3 * half, dst_planar, dst_planar + half, half);
        ofb::c3_extract_channel0(src_c3 + 3 * 2 * half, dst_planar + 2 * half, n_pixels - 2 * half);
    } catch (...) {
        set_error("c3_extract_host: could not start the host threads");
        return OFB_ERR_NOMEM;
    }
    return OFB_OK;
}

int ofb_ctx_device(const ofb_ctx *c, int *device)
{
    OFB_CHECK_CTX(c);
    if (device) *device = c->device;
    return OFB_OK;
}

int ofb_ctx_sm_count(const ofb_ctx *c, int *sm_count)
{
    OFB_CHECK_CTX(c);
    if (sm_count) *sm_count = c->sm_count;
    return OFB_OK;
}

int ofb_ctx_set_solve(ofb_ctx *c, int solve_mode)
{
    OFB_CHECK_CTX(c);
    if (solve_mode != OFB_SOLVE_EXACT && solve_mode != OFB_SOLVE_FAST) {
        set_error("unknown solve mode %d", solve_mode);
        return OFB_ERR_INVALID;
    }
    c->solve_fast = solve_mode == OFB_SOLVE_FAST;
    return OFB_OK;
}

int ofb_ctx_get_solve(const ofb_ctx *c, int *solve_mode)
{
    OFB_CHECK_CTX(c);
    if (solve_mode) *solve_mode = c->solve_fast ? OFB_SOLVE_FAST : OFB_SOLVE_EXACT;
    return OFB_OK;
}

int ofb_ctx_launch_count(const ofb_ctx *c, unsigned long long *count)
{
    OFB_CHECK_CTX(c);
    if (count) *count = c->launches;
    return OFB_OK;
}

int ofb_ctx_profile_enable(ofb_ctx *c, int on)
{
    OFB_CHECK_CTX(c);
    OFB_GUARD(c);
    c->prof_on = on != 0;
    for (auto &r : c->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    c->prof.clear();
    return OFB_OK;
}

int ofb_ctx_profile_read(ofb_ctx *c, int tag, double *ms_sum, unsigned long long *n_records)
{
    OFB_CHECK_CTX(c);
    OFB_GUARD(c);
    double sum = 0.0;
    unsigned long long n = 0;
    for (auto &r : c->prof) {
        if (r.tag != tag) continue;
        OFB_CUDA_TRY(cudaEventSynchronize(r.b));
        float ms = 0.0f;
        OFB_CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        sum += ms;
        n++;
    }
    if (ms_sum) *ms_sum = sum;
    if (n_records) *n_records = n;
    return OFB_OK;
}

int ofb_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) {
        set_error("ptr is NULL");
        return OFB_ERR_INVALID;
    }
    // (write-combined pinned memory for the upload buffers was measured: no difference, 4.58 k Mpx-pairs/s either way)
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return OFB_ERR_NOMEM;
    }
    return OFB_OK;
}

int ofb_host_free(void *ptr)
{
    if (ptr) OFB_CUDA_TRY(cudaFreeHost(ptr));
    return OFB_OK;
}

// ---------------------------------------------------------------------------------------------
// device-resident hot path
// ---------------------------------------------------------------------------------------------
int ofb_flow_pairs_device(ofb_ctx *c, const ofb_params *p, const uint8_t *prev_d, const uint8_t *next_d,
                          size_t pitch_bytes, size_t image_stride_bytes, float *const *flow_levels_d, float *total_flow_d,
                          void *stream)
{
    OFB_CHECK_CTX(c);
    int rc = check_params(p);
    if (rc) return rc;
    if (!prev_d || !next_d || !flow_levels_d) {
        set_error("NULL frame or flow pointer");
        return OFB_ERR_INVALID;
    }
    for (int k = 0; k < p->levels; k++)
        if (!flow_levels_d[k]) {
            set_error("flow_levels_d[%d] is NULL", k);
            return OFB_ERR_INVALID;
        }
    if (pitch_bytes < (size_t)p->w || (p->n_pairs > 1 && image_stride_bytes < pitch_bytes * (size_t)p->h)) {
        set_error("pitch %zu / image stride %zu too small for %dx%d", pitch_bytes, image_stride_bytes, p->w, p->h);
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    Carver cv;
    PairPlan pl;
    plan_pairs(p, &pl, &cv);
    rc = ws_reserve(c, pl.bytes);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = ws_acquire(c, st))) return rc;
    rc = run_pairs_device(c, p, pl, c->ws, prev_d, next_d, pitch_bytes, image_stride_bytes, flow_levels_d, total_flow_d, st);
    if (rc) return rc;
    return ws_release(c, st);
}

int ofb_ctx_reserve_pairs(ofb_ctx *c, const ofb_params *p)
{
    OFB_CHECK_CTX(c);
    int rc = check_params(p);
    if (rc) return rc;
    OFB_GUARD(c);
    Carver cv;
    PairPlan pl;
    plan_pairs(p, &pl, &cv);
    return ws_reserve(c, pl.bytes);
}

int ofb_pyr_down_device(ofb_ctx *c, const uint8_t *src_d, size_t src_pitch, size_t src_image_stride, int sw, int sh,
                        uint8_t *dst_d, size_t dst_pitch, size_t dst_image_stride, int n_images, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!src_d || !dst_d) {
        set_error("NULL image pointer");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    return launch_pyr_down(src_d, src_pitch, src_image_stride, sw, sh, dst_d, dst_pitch, dst_image_stride, n_images, 1,
                           static_cast<cudaStream_t>(stream), &c->launches);
}

int ofb_pyr_down2_device(ofb_ctx *c, const uint8_t *src_d, size_t src_pitch, size_t src_image_stride, int sw, int sh,
                         uint8_t *dst1_d, size_t dst1_pitch, size_t dst1_image_stride, uint8_t *dst2_d, size_t dst2_pitch,
                         size_t dst2_image_stride, int n_images, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!src_d || !dst1_d || !dst2_d) {
        set_error("NULL image pointer");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    return launch_pyr_roll(src_d, src_pitch, src_image_stride, sw, sh, dst1_d, dst1_pitch, dst1_image_stride, dst2_d, dst2_pitch,
                           dst2_image_stride, n_images, static_cast<cudaStream_t>(stream), &c->launches, nullptr, nullptr, nullptr,
                           c->sm_count);
}

int ofb_pyr_down_strip_device(ofb_ctx *c, const uint8_t *src_d, size_t src_pitch, int sw, int src_rows, int src_y_off,
                              uint8_t *dst_d, size_t dst_pitch, int dst_y0, int dst_y1, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!src_d || !dst_d) {
        set_error("NULL image pointer");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    return launch_pyr_down_strip(src_d, src_pitch, sw, src_rows, src_y_off, dst_d, dst_pitch, dst_y0, dst_y1,
                                 static_cast<cudaStream_t>(stream), &c->launches);
}

int ofb_lk_level_device(ofb_ctx *c, const uint8_t *prev_d, const uint8_t *next_d, size_t pitch_bytes,
                        size_t image_stride_bytes, int w, int h, int n_pairs, int win, int warp_mode, float flow_scale,
                        const float *cum_in_d, float *flow_out_d, float *cum_out_d, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!prev_d || !next_d || !flow_out_d) {
        set_error("NULL image or flow pointer");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    LkLevelArgs a{};
    a.prev = prev_d;
    a.next = next_d;
    a.pitch = pitch_bytes;
    a.image_stride = image_stride_bytes;
    a.w = w;
    a.h_local = h;
    a.y_off = 0;
    a.h_global = h;
    a.out_y0 = 0;
    a.out_y1 = h;
    a.n_pairs = n_pairs;
    a.win = win;
    a.warp_mode = warp_mode;
    a.flow_scale = flow_scale;
    a.cum_in = cum_in_d;
    a.cum_w = w >> 1;
    a.cum_h_global = h >> 1;
    a.cum_y_off = 0;
    a.cum_h_local = h >> 1;
    a.cum_pair_stride = (size_t)(w >> 1) * (h >> 1);
    a.flow_out = flow_out_d;
    a.cum_out = cum_out_d;
    a.flow_pair_stride = (size_t)w * h;
    a.sm_count = c->sm_count;
    a.solve_fast = c->solve_fast;
    if (cum_in_d && ((w >> 1) < 1 || (h >> 1) < 1)) {
        set_error("level %dx%d has no coarser level", w, h);
        return OFB_ERR_INVALID;
    }
    return launch_lk_level(a, static_cast<cudaStream_t>(stream), &c->launches);
}

int ofb_lk_level_strip_device(ofb_ctx *c, const uint8_t *prev_d, const uint8_t *next_d, size_t pitch_bytes, int w,
                              int h_local, int y_off, int h_global, int out_y0, int out_y1, int win, int warp_mode,
                              float flow_scale, const float *cum_in_d, int cum_y_off, int cum_h_local, float *flow_out_d,
                              float *cum_out_d, int *reach_overflow_d, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!prev_d || !next_d || !flow_out_d) {
        set_error("NULL image or flow pointer");
        return OFB_ERR_INVALID;
    }
    if (cum_in_d && warp_mode == OFB_WARP_AS_WRITTEN) {
        set_error("OFB_WARP_AS_WRITTEN needs pixel (0,0) of every coarser level and is not available on row strips");
        return OFB_ERR_UNSUPPORTED;
    }
    if (y_off < 0 || y_off + h_local > h_global) {
        set_error("strip rows [%d,%d) outside the level height %d", y_off, y_off + h_local, h_global);
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    LkLevelArgs a{};
    a.prev = prev_d;
    a.next = next_d;
    a.pitch = pitch_bytes;
    a.image_stride = pitch_bytes * (size_t)h_local;
    a.w = w;
    a.h_local = h_local;
    a.y_off = y_off;
    a.h_global = h_global;
    a.out_y0 = out_y0;
    a.out_y1 = out_y1;
    a.n_pairs = 1;
    a.win = win;
    a.warp_mode = warp_mode;
    a.flow_scale = flow_scale;
    a.cum_in = cum_in_d;
    a.cum_w = w >> 1;
    a.cum_h_global = h_global >> 1;
    a.cum_y_off = cum_y_off;
    a.cum_h_local = cum_h_local;
    a.cum_pair_stride = 0;
    a.flow_out = flow_out_d;
    a.cum_out = cum_out_d;
    a.flow_pair_stride = 0;
    a.reach_overflow = reach_overflow_d;
    a.sm_count = c->sm_count;
    a.solve_fast = c->solve_fast;
    return launch_lk_level(a, static_cast<cudaStream_t>(stream), &c->launches);
}

int ofb_c3_to_planar_device(ofb_ctx *c, const uint8_t *src_c3_d, int w, int h, int n_images, uint8_t *dst_d,
                            size_t dst_pitch, size_t dst_image_stride, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!src_c3_d || !dst_d) {
        set_error("NULL image pointer");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    return launch_c3_to_planar(src_c3_d, w, h, n_images, dst_d, dst_pitch, dst_image_stride,
                               static_cast<cudaStream_t>(stream), &c->launches);
}

// ---------------------------------------------------------------------------------------------
// host-pointer entry points (synchronous)
// ---------------------------------------------------------------------------------------------
int ofb_gauss_pyramid_host_u8c3(ofb_ctx *c, unsigned char **pyramid, int w, int h, int levels)
{
    OFB_CHECK_CTX(c);
    if (!pyramid || levels < 1 || levels > OFB_MAX_LEVELS || w < 1 || h < 1 || (w >> (levels - 1)) < 1 ||
        (h >> (levels - 1)) < 1) {
        set_error("gauss_pyramid: bad arguments (w %d h %d levels %d)", w, h, levels);
        return OFB_ERR_INVALID;
    }
    for (int k = 0; k < levels; k++)
        if (!pyramid[k]) {
            set_error("gauss_pyramid: pyramid[%d] is NULL", k);
            return OFB_ERR_INVALID;
        }
    if (levels == 1) return OFB_OK;
    OFB_GUARD(c);
    Carver cv;
    size_t off[OFB_MAX_LEVELS];
    for (int k = 0; k < levels; k++) off[k] = cv.take((size_t)(w >> k) * (h >> k) * 3);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + off[0], pyramid[0], (size_t)w * h * 3, cudaMemcpyHostToDevice, st));
    for (int k = 1; k < levels; k++) {
        const int sw = w >> (k - 1), sh = h >> (k - 1), dw = w >> k, dh = h >> k;
        rc = launch_pyr_down(c->ws + off[k - 1], (size_t)sw * 3, 0, sw, sh, c->ws + off[k], (size_t)dw * 3, 0, 1, 3, st,
                             &c->launches);
        if (rc) return rc;
        OFB_CUDA_TRY(cudaMemcpyAsync(pyramid[k], c->ws + off[k], (size_t)dw * dh * 3, cudaMemcpyDeviceToHost, st));
    }
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_calc_opt_flow_host_u8c3(ofb_ctx *c, const unsigned char *prev, const unsigned char *next, int w, int h,
                                float **optFlowPyramid, int level, int maxLevel, int win, int warp_mode, float flow_scale)
{
    OFB_CHECK_CTX(c);
    if (!prev || !next || !optFlowPyramid || w < 1 || h < 1 || level < 0 || level >= maxLevel ||
        maxLevel > OFB_MAX_LEVELS) {
        set_error("calc_opt_flow: bad arguments (w %d h %d level %d maxLevel %d)", w, h, level, maxLevel);
        return OFB_ERR_INVALID;
    }
    for (int k = level; k < maxLevel; k++) {
        if (!optFlowPyramid[k]) {
            set_error("calc_opt_flow: optFlowPyramid[%d] is NULL", k);
            return OFB_ERR_INVALID;
        }
        if ((w >> (k - level)) < 1 || (h >> (k - level)) < 1) {
            set_error("calc_opt_flow: level %d of a %dx%d level-%d image is empty", k, w, h, level);
            return OFB_ERR_INVALID;
        }
    }
    OFB_GUARD(c);
    const size_t pitch = align_up((size_t)w, 64);
    Carver cv;
    const size_t off_c3 = cv.take((size_t)w * h * 3 * 2);
    const size_t off_p = cv.take(pitch * h), off_n = cv.take(pitch * h);
    const size_t off_flow = cv.take((size_t)w * h * 8);
    size_t off_res[OFB_MAX_LEVELS], off_cum[OFB_MAX_LEVELS];
    for (int k = level + 1; k < maxLevel; k++) {
        const size_t n = (size_t)(w >> (k - level)) * (h >> (k - level));
        off_res[k] = cv.take(n * 8);
        off_cum[k] = cv.take(n * 8);
    }
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    uint8_t *B = c->ws;
    const size_t c3 = (size_t)w * h * 3;
    OFB_CUDA_TRY(cudaMemcpyAsync(B + off_c3, prev, c3, cudaMemcpyHostToDevice, st));
    OFB_CUDA_TRY(cudaMemcpyAsync(B + off_c3 + c3, next, c3, cudaMemcpyHostToDevice, st));
    rc = launch_c3_to_planar(B + off_c3, w, h, 1, B + off_p, pitch, pitch * h, st, &c->launches);
    if (rc) return rc;
    rc = launch_c3_to_planar(B + off_c3 + c3, w, h, 1, B + off_n, pitch, pitch * h, st, &c->launches);
    if (rc) return rc;
    // cumulative flow of level+1 from the residual flows the caller holds (main.cu:136-147)
    const float *cum_in = nullptr;
    for (int k = maxLevel - 1; k > level; k--) {
        const int wk = w >> (k - level), hk = h >> (k - level);
        OFB_CUDA_TRY(cudaMemcpyAsync(B + off_res[k], optFlowPyramid[k], (size_t)wk * hk * 8, cudaMemcpyHostToDevice, st));
        if (k == maxLevel - 1) {
            cum_in = reinterpret_cast<const float *>(B + off_res[k]);
        } else {
            rc = launch_compose_cum(reinterpret_cast<const float *>(B + off_res[k]), cum_in, wk, hk, 1,
                                    reinterpret_cast<float *>(B + off_cum[k]), st, &c->launches);
            if (rc) return rc;
            cum_in = reinterpret_cast<const float *>(B + off_cum[k]);
        }
    }
    LkLevelArgs a{};
    a.prev = B + off_p;
    a.next = B + off_n;
    a.pitch = pitch;
    a.image_stride = pitch * h;
    a.w = w;
    a.h_local = h;
    a.h_global = h;
    a.out_y1 = h;
    a.n_pairs = 1;
    a.win = win;
    a.warp_mode = warp_mode;
    a.flow_scale = flow_scale;
    a.cum_in = cum_in;
    a.cum_w = w >> 1;
    a.cum_h_global = h >> 1;
    a.cum_h_local = h >> 1;
    a.flow_out = reinterpret_cast<float *>(B + off_flow);
    a.sm_count = c->sm_count;
    a.solve_fast = c->solve_fast;
    rc = launch_lk_level(a, st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(optFlowPyramid[level], B + off_flow, (size_t)w * h * 8, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_conv_3ch_1ch_u8_f32_host(ofb_ctx *c, const unsigned char *src_h, int w, int h, float *dest_h, const float *mask,
                                 int mw, int mh)
{
    OFB_CHECK_CTX(c);
    if (!src_h || !dest_h || !mask || w < 1 || h < 1) {
        set_error("conv: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    Carver cv;
    const size_t off_s = cv.take((size_t)w * h * 3), off_d = cv.take((size_t)w * h * 4);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + off_s, src_h, (size_t)w * h * 3, cudaMemcpyHostToDevice, st));
    rc = launch_conv_c3_f32(c->ws + off_s, w, h, reinterpret_cast<float *>(c->ws + off_d), mask, mw, mh, st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(dest_h, c->ws + off_d, (size_t)w * h * 4, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_srm_1ch_f32_host(ofb_ctx *c, const float *arr1_h, const float *arr2_h, int w, int h, int ww, int wh,
                         float *dest_h)
{
    OFB_CHECK_CTX(c);
    if (!arr1_h || !arr2_h || !dest_h || w < 1 || h < 1) {
        set_error("srm: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    const size_t n = (size_t)w * h * 4;
    Carver cv;
    const size_t oa = cv.take(n), ob = cv.take(n), od = cv.take(n);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + oa, arr1_h, n, cudaMemcpyHostToDevice, st));
    const float *bdev = reinterpret_cast<const float *>(c->ws + oa);
    if (arr2_h != arr1_h) {
        OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + ob, arr2_h, n, cudaMemcpyHostToDevice, st));
        bdev = reinterpret_cast<const float *>(c->ws + ob);
    }
    rc = launch_srm_f32(reinterpret_cast<const float *>(c->ws + oa), bdev, w, h, ww, wh,
                        reinterpret_cast<float *>(c->ws + od), st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(dest_h, c->ws + od, n, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_inverse_matrix_f32_host(ofb_ctx *c, const float *sumIx2, const float *sumIy2, const float *sumIxIy,
                                const float *sumIxIt, const float *sumIyIt, float **optFlowPyramid, int level, int w, int h)
{
    OFB_CHECK_CTX(c);
    if (!sumIx2 || !sumIy2 || !sumIxIy || !sumIxIt || !sumIyIt || !optFlowPyramid || level < 0 || !optFlowPyramid[level] ||
        w < 1 || h < 1) {
        set_error("inverse_matrix: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    const size_t n = (size_t)w * h * 4;
    Carver cv;
    size_t o[5];
    for (int k = 0; k < 5; k++) o[k] = cv.take(n);
    const size_t of = cv.take(2 * n);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    const float *src[5] = {sumIx2, sumIy2, sumIxIy, sumIxIt, sumIyIt};
    for (int k = 0; k < 5; k++) OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + o[k], src[k], n, cudaMemcpyHostToDevice, st));
    rc = launch_inverse_f32(reinterpret_cast<const float *>(c->ws + o[0]), reinterpret_cast<const float *>(c->ws + o[1]),
                            reinterpret_cast<const float *>(c->ws + o[2]), reinterpret_cast<const float *>(c->ws + o[3]),
                            reinterpret_cast<const float *>(c->ws + o[4]), reinterpret_cast<float *>(c->ws + of), w * h,
                            st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(optFlowPyramid[level], c->ws + of, 2 * n, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

// flow_levels_h: NULL, or an array whose NULL entries are levels the caller does not want downloaded; total_flow_h: NULL
// or the level-0 composition (main.cu:136-147).  At least one output.
static int flow_pairs_host_impl(ofb_ctx *c, const ofb_params *p, const unsigned char *prev_h, const unsigned char *next_h,
                                int channels, float *const *flow_levels_h, float *total_flow_h)
{
    OFB_CHECK_CTX(c);
    int rc = check_params(p);
    if (rc) return rc;
    if (!prev_h || !next_h || (!flow_levels_h && !total_flow_h) || (channels != 1 && channels != 3)) {
        set_error("flow_pairs_host: bad arguments (channels %d)", channels);
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    // Software pipeline over sub-batches on OFB_LANES streams: while one lane downloads its flow the
    // other uploads and computes, so the PCIe directions and the SMs overlap (with pinned host memory).
    // Sub-batches are kept small (at least eight per lane when the batch allows): the first upload and the last
    // download are the part of the pipeline nothing overlaps, and a 1080p pair already is 12 MB in and 22 MB out.
    const int n = p->n_pairs;
    const int LANES = c->lanes;
    int sub = n / (8 * LANES);
    if (sub < 1) sub = 1;
    if (sub > 8) sub = 8;
    if (c->sub_pairs > 0) sub = c->sub_pairs < n ? c->sub_pairs : n;
    ofb_params ps = *p;
    ps.n_pairs = sub;
    const size_t pitch0 = align_up((size_t)p->w, 64), istride0 = pitch0 * (size_t)p->h;
    const size_t c3 = (size_t)p->w * p->h * 3;
    size_t lane_bytes = 0;
    size_t off_p0, off_n0, off_c3 = 0, off_flow[OFB_MAX_LEVELS], off_total = 0, plan_base;
    PairPlan pl;
    {
        Carver cv;
        off_p0 = cv.take(istride0 * sub);
        off_n0 = cv.take(istride0 * sub);
        if (channels == 3) off_c3 = cv.take(c3 * sub * 2);
        for (int k = 0; k < p->levels; k++) off_flow[k] = cv.take((size_t)(p->w >> k) * (p->h >> k) * 8 * sub);
        if (total_flow_h) off_total = cv.take((size_t)p->w * p->h * 8 * sub);
        plan_base = cv.off;
        Carver cv2;
        plan_pairs(&ps, &pl, &cv2);
        lane_bytes = align_up(plan_base + pl.bytes, 256);
    }
    rc = ws_reserve(c, lane_bytes * LANES);
    if (rc) return rc;
    // 3-channel input with host threads: channel 0 is extracted on the host into pinned planar staging buffers (one
    // prev + one next buffer per lane) and only those bytes are uploaded (host_extract.cpp)
    const bool host_extract = channels == 3 && c->host_threads > 0;
    const size_t frame_px = (size_t)p->w * p->h, stage_lane = 2 * frame_px * sub;
    bool stage_used[OFB_LANES] = {};
    if (host_extract) {
        if (!c->pool) {
            try {
                c->pool = new ofb::HostPool(c->host_threads);
            } catch (...) {
                set_error("flow_pairs_host: could not start %d host threads", c->host_threads);
                return OFB_ERR_NOMEM;
            }
        }
        if (c->stage_bytes < stage_lane * LANES) {
            if (c->stage) OFB_CUDA_TRY(cudaFreeHost(c->stage));
            c->stage = nullptr;
            c->stage_bytes = 0;
            OFB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&c->stage), stage_lane * LANES, cudaHostAllocDefault));
            c->stage_bytes = stage_lane * LANES;
        }
        for (int lane = 0; lane < LANES; lane++)
            if (!c->stage_ev[lane]) OFB_CUDA_TRY(cudaEventCreateWithFlags(&c->stage_ev[lane], cudaEventDisableTiming));
    }
    for (int lane = 0; lane < LANES; lane++)
        if ((rc = ws_acquire(c, c->lane_stream[lane]))) return rc;
    for (int first = 0, sb = 0; first < n; first += sub, sb++) {
        const int lane = sb % LANES;
        const int cnt = (n - first < sub) ? n - first : sub;
        cudaStream_t st = c->lane_stream[lane];
        uint8_t *B = c->ws + lane_bytes * lane;
        ofb_params pc = *p;
        pc.n_pairs = cnt;
        if (channels == 1) {
            OFB_CUDA_TRY(cudaMemcpy2DAsync(B + off_p0, pitch0, prev_h + (size_t)first * p->w * p->h, (size_t)p->w, (size_t)p->w,
                                           (size_t)p->h * cnt, cudaMemcpyHostToDevice, st));
            OFB_CUDA_TRY(cudaMemcpy2DAsync(B + off_n0, pitch0, next_h + (size_t)first * p->w * p->h, (size_t)p->w, (size_t)p->w,
                                           (size_t)p->h * cnt, cudaMemcpyHostToDevice, st));
        } else if (host_extract) {
            uint8_t *sp = c->stage + stage_lane * lane, *sn = sp + frame_px * sub;
            if (stage_used[lane]) OFB_CUDA_TRY(cudaEventSynchronize(c->stage_ev[lane])); // the lane's previous upload has left the buffers
            ofb::c3_extract_pair_sets(c->pool, prev_h + c3 * first, next_h + c3 * first, sp, sn, frame_px * cnt);
            OFB_CUDA_TRY(cudaMemcpy2DAsync(B + off_p0, pitch0, sp, (size_t)p->w, (size_t)p->w, (size_t)p->h * cnt,
                                           cudaMemcpyHostToDevice, st));
            OFB_CUDA_TRY(cudaMemcpy2DAsync(B + off_n0, pitch0, sn, (size_t)p->w, (size_t)p->w, (size_t)p->h * cnt,
                                           cudaMemcpyHostToDevice, st));
            OFB_CUDA_TRY(cudaEventRecord(c->stage_ev[lane], st));
            stage_used[lane] = true;
        } else {
            OFB_CUDA_TRY(cudaMemcpyAsync(B + off_c3, prev_h + c3 * first, c3 * cnt, cudaMemcpyHostToDevice, st));
            OFB_CUDA_TRY(cudaMemcpyAsync(B + off_c3 + c3 * sub, next_h + c3 * first, c3 * cnt, cudaMemcpyHostToDevice, st));
            rc = launch_c3_to_planar(B + off_c3, p->w, p->h, cnt, B + off_p0, pitch0, istride0, st, &c->launches);
            if (rc) return rc;
            rc = launch_c3_to_planar(B + off_c3 + c3 * sub, p->w, p->h, cnt, B + off_n0, pitch0, istride0, st, &c->launches);
            if (rc) return rc;
        }
        float *flow_d[OFB_MAX_LEVELS];
        for (int k = 0; k < p->levels; k++) flow_d[k] = reinterpret_cast<float *>(B + off_flow[k]);
        float *total_d = total_flow_h ? reinterpret_cast<float *>(B + off_total) : nullptr;
        // (one level: the total flow is the residual flow, the kernel has nothing to compose)
        rc = run_pairs_device(c, &pc, pl, B + plan_base, B + off_p0, B + off_n0, pitch0, istride0, flow_d,
                              p->levels > 1 ? total_d : nullptr, st);
        if (rc) return rc;
        for (int k = p->levels - 1; k >= 0; k--) {
            if (!flow_levels_h || !flow_levels_h[k]) continue;
            const size_t per_pair = (size_t)(p->w >> k) * (p->h >> k) * 2; // floats
            OFB_CUDA_TRY(cudaMemcpyAsync(flow_levels_h[k] + per_pair * first, flow_d[k], per_pair * 4 * cnt,
                                         cudaMemcpyDeviceToHost, st));
        }
        if (total_flow_h) {
            const size_t per_pair = (size_t)p->w * p->h * 2;
            OFB_CUDA_TRY(cudaMemcpyAsync(total_flow_h + per_pair * first, p->levels > 1 ? total_d : flow_d[0], per_pair * 4 * cnt,
                                         cudaMemcpyDeviceToHost, st));
        }
    }
    for (int lane = 0; lane < LANES; lane++) OFB_CUDA_TRY(cudaStreamSynchronize(c->lane_stream[lane]));
    return OFB_OK;
}

int ofb_flow_pairs_host(ofb_ctx *c, const ofb_params *p, const unsigned char *prev_h, const unsigned char *next_h,
                        int channels, float *const *flow_levels_h)
{
    if (!flow_levels_h) {
        set_error("flow_pairs_host: flow_levels_h is NULL");
        return OFB_ERR_INVALID;
    }
    if (p)
        for (int k = 0; k < p->levels && k < OFB_MAX_LEVELS; k++)
            if (!flow_levels_h[k]) {
                set_error("flow_levels_h[%d] is NULL", k);
                return OFB_ERR_INVALID;
            }
    return flow_pairs_host_impl(c, p, prev_h, next_h, channels, flow_levels_h, nullptr);
}

int ofb_flow_pairs_host_ex(ofb_ctx *c, const ofb_params *p, const unsigned char *prev_h, const unsigned char *next_h,
                           int channels, float *const *flow_levels_h, float *total_flow_h)
{
    return flow_pairs_host_impl(c, p, prev_h, next_h, channels, flow_levels_h, total_flow_h);
}

// ---------------------------------------------------------------------------------------------
// pre-processing and the frame loop
// ---------------------------------------------------------------------------------------------
int ofb_grayscale_avg_host_u8c3(ofb_ctx *c, const unsigned char *src_h, unsigned char *dest_h, int h, int w)
{
    OFB_CHECK_CTX(c);
    if (!src_h || !dest_h || w < 1 || h < 1) {
        set_error("grayscale_avg: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    const size_t n = (size_t)w * h * 3;
    Carver cv;
    const size_t os = cv.take(n), od = cv.take(n);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + os, src_h, n, cudaMemcpyHostToDevice, st));
    rc = launch_grayscale(c->ws + os, w, h, c->ws + od, nullptr, 0, st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(dest_h, c->ws + od, n, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_bilinear_filter_host_u8c3(ofb_ctx *c, const unsigned char *src, const unsigned char *gray, unsigned char *dest,
                                  int w, int h, int ww, int wh, double sigmaS, double sigmaB)
{
    OFB_CHECK_CTX(c);
    if (!src || !gray || !dest || w < 1 || h < 1) {
        set_error("bilinear_filter: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    const size_t n = (size_t)w * h * 3;
    Carver cv;
    const size_t os = cv.take(n), og = cv.take(n), od = cv.take(n), ol = cv.take(256 * sizeof(double));
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + os, src, n, cudaMemcpyHostToDevice, st));
    const uint8_t *gd = c->ws + os;
    if (gray != src) {
        OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + og, gray, n, cudaMemcpyHostToDevice, st));
        gd = c->ws + og;
    }
    rc = launch_bilateral(c->ws + os, gd, (size_t)w * 3, 3, w, h, ww, wh, sigmaS, sigmaB,
                          reinterpret_cast<double *>(c->ws + ol), c->ws + od, (size_t)w * 3, st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(dest, c->ws + od, n, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_bilateral_planar_device(ofb_ctx *c, const uint8_t *gray_d, size_t pitch, int w, int h, int ww, int wh, double sigmaS,
                                double sigmaB, uint8_t *dst_d, size_t dst_pitch, void *stream)
{
    OFB_CHECK_CTX(c);
    if (!gray_d || !dst_d) {
        set_error("NULL image pointer");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    if (!c->bil_lut) OFB_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&c->bil_lut), 256 * sizeof(double)));
    return launch_bilateral(gray_d, gray_d, pitch, 1, w, h, ww, wh, sigmaS, sigmaB, c->bil_lut, dst_d, dst_pitch,
                            static_cast<cudaStream_t>(stream), &c->launches);
}

// ---- debug derivative views (main.cu:19-92) -----------------------------------------------------

int ofb_conv_3ch_1ch_u8_u8_host(ofb_ctx *c, const unsigned char *src_h, int w, int h, unsigned char *dest_h, const float *mask,
                                int mw, int mh)
{
    OFB_CHECK_CTX(c);
    if (!src_h || !dest_h || !mask || w < 1 || h < 1) {
        set_error("conv (u8): bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    const size_t n = (size_t)w * h;
    Carver cv;
    const size_t os = cv.take(n * 3), od = cv.take(n);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + os, src_h, n * 3, cudaMemcpyHostToDevice, st));
    rc = launch_conv_c3_u8(c->ws + os, w, h, c->ws + od, mask, mw, mh, st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(dest_h, c->ws + od, n, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_debug_view_host_u8c3(ofb_ctx *c, const unsigned char *prev_level_h, const unsigned char *cur_level_h, int w, int h,
                             int level, int which, unsigned char *out_h)
{
    OFB_CHECK_CTX(c);
    if (!cur_level_h || !out_h || (which == OFB_VIEW_T && !prev_level_h) || w < 1 || h < 1 || level < 0 || level > 12) {
        set_error("debug_view: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_GUARD(c);
    const size_t n = (size_t)w * h, no = ((size_t)w << level) * ((size_t)h << level);
    Carver cv;
    const size_t oc = cv.take(n * 3), op = cv.take(n * 3), oo = cv.take(no);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + oc, cur_level_h, n * 3, cudaMemcpyHostToDevice, st));
    if (which == OFB_VIEW_T) OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + op, prev_level_h, n * 3, cudaMemcpyHostToDevice, st));
    rc = launch_debug_view(c->ws + op, c->ws + oc, w, h, level, which, c->ws + oo, st, &c->launches);
    if (rc) return rc;
    OFB_CUDA_TRY(cudaMemcpyAsync(out_h, c->ws + oo, no, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

// ---- flow composition and export (the headless part of main.cu:114-174) ----------------------

int ofb_compose_flow_host(ofb_ctx *c, float *const *flow_pyramid_h, int w, int h, int levels, int level, float *total_h)
{
    OFB_CHECK_CTX(c);
    if (!flow_pyramid_h || !total_h || w < 1 || h < 1 || levels < 1 || levels > OFB_MAX_LEVELS || level < 0 || level >= levels ||
        (w >> (levels - 1)) < 1 || (h >> (levels - 1)) < 1) {
        set_error("compose_flow: bad arguments (w %d h %d levels %d level %d)", w, h, levels, level);
        return OFB_ERR_INVALID;
    }
    for (int k = level; k < levels; k++)
        if (!flow_pyramid_h[k]) {
            set_error("compose_flow: flow_pyramid[%d] is NULL", k);
            return OFB_ERR_INVALID;
        }
    OFB_GUARD(c);
    // device: residual flow of levels level..levels-1 and two ping-pong cumulative buffers of the finest size
    Carver cv;
    size_t off_res[OFB_MAX_LEVELS];
    for (int k = level; k < levels; k++) off_res[k] = cv.take((size_t)(w >> k) * (h >> k) * 8);
    const size_t n0 = (size_t)(w >> level) * (h >> level) * 8;
    const size_t off_a = cv.take(n0), off_b = cv.take(n0);
    int rc = ws_reserve(c, cv.off);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    if ((rc = ws_acquire(c, st))) return rc;
    for (int k = level; k < levels; k++)
        OFB_CUDA_TRY(cudaMemcpyAsync(c->ws + off_res[k], flow_pyramid_h[k], (size_t)(w >> k) * (h >> k) * 8,
                                     cudaMemcpyHostToDevice, st));
    // coarse to fine: cum_{L-1} = flow_{L-1}; cum_k = 2*cum_{k+1}[i>>1][j>>1] + flow_k
    const float *cum = nullptr;
    size_t cur = off_a, other = off_b;
    for (int k = levels - 1; k >= level; k--) {
        rc = launch_compose_cum(reinterpret_cast<const float *>(c->ws + off_res[k]), cum, w >> k, h >> k, 1,
                                reinterpret_cast<float *>(c->ws + cur), st, &c->launches);
        if (rc) return rc;
        cum = reinterpret_cast<const float *>(c->ws + cur);
        std::swap(cur, other);
    }
    OFB_CUDA_TRY(cudaMemcpyAsync(total_h, cum, n0, cudaMemcpyDeviceToHost, st));
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    return OFB_OK;
}

int ofb_flow_arrows_host(ofb_ctx *c, float *const *flow_pyramid_h, int w, int h, int levels, int level, int arrow_res,
                         int *arrows_xyxy, int max_arrows, int *n_arrows)
{
    OFB_CHECK_CTX(c);
    if (!n_arrows || arrow_res < 1 || max_arrows < 0 || (max_arrows > 0 && !arrows_xyxy)) {
        set_error("flow_arrows: bad arguments");
        return OFB_ERR_INVALID;
    }
    if (level < 0 || level >= levels || w < 1 || h < 1) {
        set_error("flow_arrows: bad geometry (w %d h %d levels %d level %d)", w, h, levels, level);
        return OFB_ERR_INVALID;
    }
    const int wl = w >> level, hl = h >> level;
    const int step = wl / arrow_res; // main.cu:124
    if (step < 1) {
        set_error("flow_arrows: arrow_res %d exceeds the level width %d", arrow_res, wl);
        return OFB_ERR_INVALID;
    }
    std::vector<float> total((size_t)wl * hl * 2);
    int rc = ofb_compose_flow_host(c, flow_pyramid_h, w, h, levels, level, total.data());
    if (rc) return rc;
    int n = 0;
    for (int i = 0; i < hl; i += step)
        for (int j = 0; j < wl; j += step) {
            float u = total[((size_t)i * wl + j) * 2], v = total[((size_t)i * wl + j) * 2 + 1];
            // main.cu:150-157: clamp to one grid step; NaN compares false and passes through, as in the reference
            if (u > step) u = (float)step;
            else if (u < -step) u = (float)-step;
            if (v > step) v = (float)step;
            else if (v < -step) v = (float)-step;
            // main.cu:159-160 casts float to int; the result is undefined there for NaN (x86 gives INT_MIN, which the
            // sign test then drops): defined here as "no arrow"
            if (u != u || v != v) continue;
            const int ni = (int)(v + (float)i), nj = (int)(u + (float)j);
            if (ni < 0 || nj < 0) continue; // main.cu:163
            if (n < max_arrows) {
                arrows_xyxy[4 * n + 0] = j;
                arrows_xyxy[4 * n + 1] = i;
                arrows_xyxy[4 * n + 2] = nj;
                arrows_xyxy[4 * n + 3] = ni;
            }
            n++;
        }
    *n_arrows = n;
    return OFB_OK;
}

int ofb_write_flo(const char *path, const float *flow_h, int w, int h)
{
    if (!path || !flow_h || w < 1 || h < 1) {
        set_error("write_flo: bad arguments");
        return OFB_ERR_INVALID;
    }
    FILE *f = fopen(path, "wb");
    if (!f) {
        set_error("write_flo: cannot open %s", path);
        return OFB_ERR_INVALID;
    }
    const float tag = 202021.25f; // "PIEH"
    const int32_t dims[2] = {w, h};
    bool ok = fwrite(&tag, 4, 1, f) == 1 && fwrite(dims, 4, 2, f) == 2 &&
              fwrite(flow_h, 8, (size_t)w * h, f) == (size_t)w * h;
    ok = (fclose(f) == 0) && ok;
    if (!ok) {
        set_error("write_flo: short write to %s", path);
        return OFB_ERR_INVALID;
    }
    return OFB_OK;
}

} // extern "C"

// One frame sequence: two device-resident planar pyramids that swap roles every frame (main.cu:270-272).
struct ofb_stream {
    ofb_ctx *ctx = nullptr;
    ofb_params p{};
    int bil_win = 0;
    double sig_s = 0, sig_b = 0;
    int frames = 0;
    uint8_t *mem = nullptr; // all device buffers of the stream in one allocation
    size_t pitch[OFB_MAX_LEVELS] = {}, off_pyr[2][OFB_MAX_LEVELS] = {}, off_bgr = 0, off_gray = 0, off_lut = 0;
    size_t off_flow[OFB_MAX_LEVELS] = {}, off_cum[OFB_MAX_LEVELS] = {}, off_total = 0;
    int cur = 0; // which pyramid receives the next frame
};

extern "C" {

int ofb_stream_create(ofb_ctx *c, const ofb_params *p, int bil_win, double bil_sigma_s, double bil_sigma_b, ofb_stream **out)
{
    OFB_CHECK_CTX(c);
    if (!out) {
        set_error("out is NULL");
        return OFB_ERR_INVALID;
    }
    *out = nullptr;
    int rc = check_params(p);
    if (rc) return rc;
    if (p->n_pairs != 1) {
        set_error("a frame sequence solves one pair per frame (n_pairs must be 1)");
        return OFB_ERR_INVALID;
    }
    if (bil_win != 0 && (bil_win < 3 || bil_win > 9 || !(bil_win & 1))) {
        set_error("bilateral window %d not supported (0 = off, or odd 3..9)", bil_win);
        return OFB_ERR_UNSUPPORTED;
    }
    OFB_GUARD(c);
    ofb_stream *s = new (std::nothrow) ofb_stream();
    if (!s) {
        set_error("out of host memory");
        return OFB_ERR_NOMEM;
    }
    s->ctx = c;
    s->p = *p;
    s->bil_win = bil_win;
    s->sig_s = bil_sigma_s;
    s->sig_b = bil_sigma_b;
    Carver cv;
    s->off_bgr = cv.take((size_t)p->w * p->h * 3);
    for (int k = 0; k < p->levels; k++) {
        const int wk = p->w >> k, hk = p->h >> k;
        s->pitch[k] = align_up((size_t)wk, 64);
        s->off_pyr[0][k] = cv.take(s->pitch[k] * hk);
        s->off_pyr[1][k] = cv.take(s->pitch[k] * hk);
        s->off_flow[k] = cv.take((size_t)wk * hk * 8);
        s->off_cum[k] = cv.take((size_t)wk * hk * 8);
    }
    s->off_gray = cv.take(s->pitch[0] * p->h);
    s->off_total = cv.take((size_t)p->w * p->h * 8);
    s->off_lut = cv.take(256 * sizeof(double));
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&s->mem), cv.off);
    if (e != cudaSuccess) {
        set_error("stream cudaMalloc(%zu) failed: %s", cv.off, cudaGetErrorString(e));
        cudaGetLastError();
        delete s;
        return OFB_ERR_NOMEM;
    }
    e = cudaMemsetAsync(s->mem, 0, cv.off, c->stream);
    if (e != cudaSuccess) {
        set_error("stream cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaFree(s->mem);
        delete s;
        return OFB_ERR_CUDA;
    }
    *out = s;
    return OFB_OK;
}

int ofb_stream_destroy(ofb_stream *s)
{
    if (!s) {
        set_error("stream is NULL");
        return OFB_ERR_INVALID;
    }
    {
        DeviceGuard g(s->ctx->device);
        if (g.ok) {
            cudaStreamSynchronize(s->ctx->stream);
            if (s->mem) cudaFree(s->mem);
        }
    }
    delete s;
    return OFB_OK;
}

int ofb_stream_push_bgr_host(ofb_stream *s, const unsigned char *frame_bgr, float *const *flow_levels_h, float *total_flow_h,
                             int *has_flow)
{
    if (!s || !frame_bgr) {
        set_error("stream_push: NULL stream or frame");
        return OFB_ERR_INVALID;
    }
    ofb_ctx *c = s->ctx;
    const ofb_params &p = s->p;
    if (s->frames > 0) { // (before anything is enqueued: a rejected call leaves the sequence where it was)
        if (!flow_levels_h) {
            set_error("stream_push: flow_levels_h is NULL");
            return OFB_ERR_INVALID;
        }
        for (int k = 0; k < p.levels; k++)
            if (!flow_levels_h[k]) {
                set_error("stream_push: flow_levels_h[%d] is NULL", k);
                return OFB_ERR_INVALID;
            }
    }
    OFB_GUARD(c);
    cudaStream_t st = c->stream;
    uint8_t *B = s->mem;
    const int cur = s->cur, prv = cur ^ 1;
    if (has_flow) *has_flow = 0;
    // upload, grayscale (main.cu:232), bilateral pre-filter (main.cu:240) into level 0 of the current pyramid
    OFB_CUDA_TRY(cudaMemcpyAsync(B + s->off_bgr, frame_bgr, (size_t)p.w * p.h * 3, cudaMemcpyHostToDevice, st));
    uint8_t *lvl0 = B + s->off_pyr[cur][0];
    int rc = launch_grayscale(B + s->off_bgr, p.w, p.h, nullptr, s->bil_win ? B + s->off_gray : lvl0, s->pitch[0], st,
                              &c->launches);
    if (rc) return rc;
    if (s->bil_win) {
        rc = launch_bilateral(B + s->off_gray, B + s->off_gray, s->pitch[0], 1, p.w, p.h, s->bil_win, s->bil_win, s->sig_s,
                              s->sig_b, reinterpret_cast<double *>(B + s->off_lut), lvl0, s->pitch[0], st, &c->launches);
        if (rc) return rc;
    }
    for (int k = 1; k < p.levels;) { // main.cu:250; two steps per launch where two are left
        const bool two = pyr_fuse_enabled() && k + 1 < p.levels;
        rc = launch_pyr_roll(B + s->off_pyr[cur][k - 1], s->pitch[k - 1], s->pitch[k - 1] * (size_t)(p.h >> (k - 1)), p.w >> (k - 1),
                             p.h >> (k - 1), B + s->off_pyr[cur][k], s->pitch[k], s->pitch[k] * (size_t)(p.h >> k),
                             two ? B + s->off_pyr[cur][k + 1] : nullptr, two ? s->pitch[k + 1] : 0,
                             two ? s->pitch[k + 1] * (size_t)(p.h >> (k + 1)) : 0, 1, st, &c->launches, nullptr, nullptr, nullptr,
                             c->sm_count);
        if (rc) return rc;
        k += two ? 2 : 1;
    }
    if (s->frames > 0) {
        for (int k = p.levels - 1; k >= 0; k--) { // main.cu:256-262
            LkLevelArgs a{};
            a.prev = B + s->off_pyr[prv][k];
            a.next = B + s->off_pyr[cur][k];
            a.pitch = s->pitch[k];
            a.image_stride = s->pitch[k] * (size_t)(p.h >> k);
            a.w = p.w >> k;
            a.h_local = a.h_global = p.h >> k;
            a.out_y1 = p.h >> k;
            a.n_pairs = 1;
            a.win = p.win;
            a.warp_mode = p.warp_mode;
            a.flow_scale = p.flow_scale;
            a.flow_out = reinterpret_cast<float *>(B + s->off_flow[k]);
            a.sm_count = c->sm_count;
            a.solve_fast = c->solve_fast;
            if (k < p.levels - 1) {
                a.cum_in = reinterpret_cast<const float *>(k + 1 == p.levels - 1 ? B + s->off_flow[k + 1] : B + s->off_cum[k + 1]);
                a.cum_w = p.w >> (k + 1);
                a.cum_h_global = a.cum_h_local = p.h >> (k + 1);
            }
            if (k == 0) a.cum_out = total_flow_h ? reinterpret_cast<float *>(B + s->off_total) : nullptr;
            else if (k <= p.levels - 2) a.cum_out = reinterpret_cast<float *>(B + s->off_cum[k]);
            rc = launch_lk_level(a, st, &c->launches);
            if (rc) return rc;
            OFB_CUDA_TRY(cudaMemcpyAsync(flow_levels_h[k], B + s->off_flow[k], (size_t)(p.w >> k) * (p.h >> k) * 8,
                                         cudaMemcpyDeviceToHost, st));
        }
        if (total_flow_h)
            OFB_CUDA_TRY(cudaMemcpyAsync(total_flow_h, B + s->off_total, (size_t)p.w * p.h * 8, cudaMemcpyDeviceToHost, st));
        if (has_flow) *has_flow = 1;
    }
    OFB_CUDA_TRY(cudaStreamSynchronize(st));
    s->cur = prv; // main.cu:270-272: the current pyramid becomes the previous one
    s->frames++;
    return OFB_OK;
}

} // extern "C"

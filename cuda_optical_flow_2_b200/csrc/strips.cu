// strips.cu -- row-strip solve of ONE large pair over several GPUs (BASELINE configs[4]), native host side.
//
// One process per GPU.  The strip schedule is the one of cuda_optical_flow_2_b200/dist.py (StripPlan: bounds fixed
// on the coarsest level and doubled per finer level, image halo = window radius + stencil + warp reach, coarser
// cumulative-flow halo), but here the whole pair -- own-row upload, per-level halo exchange, pyramid, fused level
// kernels -- is enqueued on one CUDA stream from C++ with no host synchronisation, and the halo rows move by NCCL
// send/recv (NVLink / NVSwitch) straight between the ranks' image and flow buffers: rows are contiguous, so there
// is no staging copy.  Only level-0 image rows are exchanged (the coarser levels' halo rows are rebuilt locally), so a
// pair of L levels costs L exchanges: one for the images, L-1 for the cumulative flow.  The Python StripRunner runs
// the 2L-1-exchange form of the schedule through torch and is bound by ~100 us of host work per exchange.
//
// Second transport, peer memory (ofb_strips_peer_*): every rank keeps the buffers its neighbours write into -- the
// level-0 image rows and the coarser cumulative-flow rows -- in ONE device allocation (the arena) whose CUDA IPC
// handle the application all-gathers once.  An exchange is then a copy kernel of the SENDER that stores the halo rows
// straight into the receiver's arena over NVLink and raises an epoch flag there, and a one-block wait kernel on the
// receiver's stream; no NCCL call, no host work beyond two launches (~5 us instead of ~30 us per exchange).
//
// NCCL is not a link-time dependency of the library: libnccl.so.2 is opened on first use (inside a torch process
// that is the copy torch already loaded), and every symbol is resolved with dlsym.
#include "ofb_common.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

struct ofb_ctx; // ofb_api.cu
namespace ofb {
// accessors implemented in ofb_api.cu (ofb_ctx is private to it)
int ctx_device(const ofb_ctx *c);
int ctx_sm_count(const ofb_ctx *c);
int ctx_solve_fast(const ofb_ctx *c);
unsigned long long *ctx_launch_counter(ofb_ctx *c);

// ---- NCCL through dlopen ------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
struct NcclUniqueId {
    char internal[128];
};
enum { NCCL_SUCCESS = 0, NCCL_UINT8 = 1 };
struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
    bool ok = false;
};
static NcclApi *nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        bool all = true;
        auto sym = [&](const char *n) {
            void *p = dlsym(h, n);
            if (!p) all = false;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = all;
    });
    return api.ok ? &api : nullptr;
}
#define OFB_NCCL_TRY(expr)                                                                                    \
    do {                                                                                                      \
        int _r = (expr);                                                                                      \
        if (_r != NCCL_SUCCESS) {                                                                             \
            ofb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, ofb::nccl()->GetErrorString(_r)); \
            return OFB_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)

// ---- the strip schedule (dist.py: StripPlan) ---------------------------------------------------------------------
struct LevelStrip {
    int w, h;     // size of the whole level
    int y0, y1;   // rows this rank owns (and produces flow for)
    int by0, by1; // rows the level kernel needs: own rows plus halo, clipped to the image
    int cy0, cy1; // rows of the next-coarser level's cumulative flow its buffer holds
    int eb0, eb1; // rows its image buffers hold: [by0, by1) plus what building [eb0, eb1) of the next level takes
};
struct StripPlan {
    int W, H, levels, win, world, reach, r, img_halo;
    std::vector<int> coarse_bounds;
    LevelStrip level(int k, int rank) const
    {
        const int sh = levels - 1 - k;
        LevelStrip s;
        s.w = W >> k;
        s.h = H >> k;
        s.y0 = coarse_bounds[rank] << sh;
        s.y1 = rank == world - 1 ? s.h : coarse_bounds[rank + 1] << sh;
        s.by0 = std::max(0, s.y0 - img_halo);
        s.by1 = std::min(s.h, s.y1 + img_halo);
        if (k < levels - 1) {
            const int hc = H >> (k + 1);
            // the kernel looks up cum[(y >> 1)] for the rows it warps (its packed tile: own rows -r-2 .. +r+1)
            s.cy0 = std::max(0, (s.y0 - r - 2) >> 1);
            s.cy1 = std::min(hc, ((s.y1 + r + 1) >> 1) + 1);
        } else {
            s.cy0 = s.cy1 = 0;
        }
        s.eb0 = s.by0;
        s.eb1 = s.by1;
        return s;
    }
    // Every level of one rank.  Only level-0 rows are exchanged; the halo rows of the coarser levels are built
    // locally (each rank repeats a few rows of pyramid work instead of three more exchanges), so a level's buffer
    // also holds the rows 2y-1 .. 2y+1 under every row y of the next level's buffer.
    std::vector<LevelStrip> strips(int rank) const
    {
        std::vector<LevelStrip> v(levels);
        for (int k = levels - 1; k >= 0; k--) {
            v[k] = level(k, rank);
            if (k + 1 < levels) {
                v[k].eb0 = std::min(v[k].by0, std::max(0, 2 * v[k + 1].eb0 - 1));
                v[k].eb1 = std::max(v[k].by1, std::min(v[k].h, 2 * (v[k + 1].eb1 - 1) + 2));
            }
        }
        return v;
    }
};

// ---- peer-memory transport ------------------------------------------------------------------------------------
// Arena of one rank: [prev level 0][next level 0][cum_in level 0 .. L-2][flags].  Every rank can compute every other
// rank's layout from the plan, so a peer address is (peer arena base) + (offset computed here).
constexpr int PEER_MAX_SEG = 8;   // copy segments per exchange (prev + next rows per destination)
constexpr int PEER_MAX_DST = 4;   // ranks one exchange sends to / receives from
struct ArenaLayout {
    size_t prev0, next0, cum_in[OFB_MAX_LEVELS], flags, bytes;
    // flags (unsigned): arrive[x * world + src] = epoch of the last exchange x that src completed into this arena;
    //                   done[src] at L * world + src = last epoch rank src finished reading what this rank pushed;
    //                   counter[x] at (L + 1) * world + x: blocks of this rank's push kernel that have finished (local);
    //                   pairs at (L + 1) * world + L: pairs this rank has completed (local).  The epoch of a pair is
    //                   pairs + 1, read from here by the kernels, so that a captured CUDA graph of one pair replays;
    //                   fused[2 * x + d] after that: CTAs of the level kernel that have pushed their rows of exchange x to
    //                   target d (local; the halo push fused into the level kernels).
};
static ArenaLayout arena_layout(const StripPlan &pl, int rank)
{
    ArenaLayout a{};
    const std::vector<LevelStrip> v = pl.strips(rank);
    size_t off = 0;
    auto take = [&](size_t n) {
        const size_t o = off;
        off += (n + 255) / 256 * 256;
        return o;
    };
    const size_t pitch0 = ((size_t)v[0].w + 63) / 64 * 64;
    a.prev0 = take(pitch0 * (v[0].eb1 - v[0].eb0));
    a.next0 = take(pitch0 * (v[0].eb1 - v[0].eb0));
    for (int k = 0; k < pl.levels; k++) a.cum_in[k] = take((size_t)(v[k].w >> 1) * std::max(v[k].cy1 - v[k].cy0, 1) * 8);
    a.flags = take(((size_t)(pl.levels + 1) * pl.world + pl.levels + 1 + 2 * (size_t)pl.levels) * sizeof(unsigned));
    a.bytes = off;
    return a;
}

struct PeerSeg {
    const void *src;
    void *dst; // inside a peer's arena
    size_t bytes; // multiple of 8
};
struct PeerPushArgs {
    PeerSeg seg[PEER_MAX_SEG];
    int nseg, ndst;
    unsigned *arrive[PEER_MAX_DST];           // arrive[x * world + me] in the arena of each destination
    const unsigned *done_local[PEER_MAX_DST]; // done[dst] in MY arena: dst has finished with what I pushed last pair
    unsigned *counter;                        // counter[x] in my arena
    const unsigned *pairs;                    // pairs completed (my arena): this pair's epoch is *pairs + 1
    int *err;
};
struct PeerWaitArgs {
    const unsigned *flag[PEER_MAX_DST]; // arrive[x * world + src] in my arena
    int n;
    const unsigned *pairs;
    int *err;
};
struct PeerDoneArgs {
    unsigned *flag[2 * PEER_MAX_DST]; // done[me] in the arena of every rank that pushes to me
    int n;
    unsigned *pairs;
};
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ull; // a peer that never shows up: flag the pair, do not hang

__device__ __forceinline__ unsigned long long peer_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned peer_ld(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// true when *p has reached `epoch` (wrap-safe), false after PEER_TIMEOUT_NS
__device__ bool peer_spin(const unsigned *p, unsigned epoch)
{
    const unsigned long long t0 = peer_now();
    while ((int)(peer_ld(p) - epoch) < 0) {
        if (peer_now() - t0 > PEER_TIMEOUT_NS) return false;
        __nanosleep(100);
    }
    return true;
}

// Sender side of one exchange: wait until every destination has finished with last pair's rows, store the segments into
// the destinations' arenas (peer stores over NVLink), and -- last block out -- raise the epoch flag in each of them.
__global__ void __launch_bounds__(256) peer_push_kernel(const __grid_constant__ PeerPushArgs a)
{
    const unsigned epoch = *a.pairs + 1u;
    if (threadIdx.x < a.ndst && !peer_spin(a.done_local[threadIdx.x], epoch - 1u)) atomicOr(a.err, 2);
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    for (int s = 0; s < a.nseg; s++) {
        const PeerSeg g = a.seg[s];
        if (((reinterpret_cast<uintptr_t>(g.src) | reinterpret_cast<uintptr_t>(g.dst) | g.bytes) & 15) == 0) {
            const uint4 *src = static_cast<const uint4 *>(g.src);
            uint4 *dst = static_cast<uint4 *>(g.dst);
            for (size_t i = t; i < g.bytes / 16; i += nt) dst[i] = src[i];
        } else {
            const uint2 *src = static_cast<const uint2 *>(g.src);
            uint2 *dst = static_cast<uint2 *>(g.dst);
            for (size_t i = t; i < g.bytes / 8; i += nt) dst[i] = src[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned arrived = atomicAdd(a.counter, 1u);
        if (arrived == gridDim.x - 1) {
            *a.counter = 0u;
            __threadfence_system();
            for (int d = 0; d < a.ndst; d++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.arrive[d]), "r"(epoch) : "memory");
        }
    }
}
// Receiver side: one thread per source spins on its epoch flag.
__global__ void peer_wait_kernel(const __grid_constant__ PeerWaitArgs a)
{
    if (threadIdx.x < a.n && !peer_spin(a.flag[threadIdx.x], *a.pairs + 1u)) atomicOr(a.err, 2);
    __threadfence_system();
}
// Teardown: wait until every rank this one pushed to has acknowledged this rank's last pair (done[dst] == pairs).
struct PeerDrainArgs {
    const unsigned *flag[2 * PEER_MAX_DST];
    int n;
    const unsigned *pairs;
    int *err;
};
__global__ void peer_drain_kernel(const __grid_constant__ PeerDrainArgs a)
{
    if (threadIdx.x < a.n && !peer_spin(a.flag[threadIdx.x], *a.pairs)) atomicOr(a.err, 2);
}
// End of a pair: tell every rank that pushes to me that its rows have been used, and count the pair.
__global__ void peer_done_kernel(const __grid_constant__ PeerDoneArgs a)
{
    const unsigned epoch = *a.pairs + 1u;
    if (threadIdx.x < a.n) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag[threadIdx.x]), "r"(epoch) : "memory");
    __syncthreads();
    if (threadIdx.x == 0) *a.pairs = epoch;
}
} // namespace ofb

using namespace ofb;

#define OFB_STRIPS_GUARD(st)                       \
    DeviceGuard _guard(ctx_device((st)->ctx));     \
    if (!_guard.ok) return OFB_ERR_CUDA

struct ofb_strips {
    ofb_ctx *ctx = nullptr;
    StripPlan plan;
    int rank = 0, warp_mode = OFB_WARP_BILINEAR;
    float flow_scale = 1.0f;
    ncclComm_t comm = nullptr;
    std::vector<LevelStrip> s;                 // this rank's strip per level
    std::vector<LevelStrip> peer0;             // every rank's level-0 strip (for the one image exchange)
    std::vector<size_t> pitch;                 // image pitch per level
    std::vector<uint8_t *> prev, next;         // [eb0, eb1) rows
    std::vector<float *> flow, cum, cum_in;    // flow / cum: buffer rows (origin eb0); cum_in[k]: rows [cy0, cy1) of cum_{k+1}
    std::vector<size_t> img_stride;            // next[k] - prev[k]: both frames of a level are one batch of two images
    std::vector<float *> cum_dst;              // where level k's kernel writes its cumulative flow (buffer-row indexing like
                                               // flow[k]): straight into cum_in[k-1], whose rows [cy0, cy1) contain the own rows
    int want_total = 1;                        // level 0 also writes its cumulative flow (the total flow of the pair)
    int fused = 1;                             // peer-memory transport: cumulative-flow halo rows are pushed by the level kernels themselves
    int *overflow = nullptr;
    std::vector<void *> allocs;
    // peer-memory transport
    uint8_t *arena = nullptr;               // my arena (one cudaMalloc: prev[0], next[0], cum_in[], flags live in it)
    ArenaLayout lay;
    std::vector<uint8_t *> peer_arena;      // base of every rank's arena in this process's address space (nullptr: not mapped)
    std::vector<ArenaLayout> peer_lay;
    std::vector<void *> ipc_opened;
    bool peers_connected = false;
    std::vector<int> done_from;             // ranks this one pushes to: they write done[rank] into this arena (teardown waits for it)
};

namespace {
int dev_alloc(ofb_strips *st, void **p, size_t bytes)
{
    OFB_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    OFB_CUDA_TRY(cudaMemset(*p, 0, bytes ? bytes : 16));
    st->allocs.push_back(*p);
    return OFB_OK;
}
// rows [lo, hi) that `owner` owns and `peer`'s buffer [b0, b1) holds
inline bool overlap(int own0, int own1, int b0, int b1, int *lo, int *hi)
{
    *lo = std::max(own0, b0);
    *hi = std::min(own1, b1);
    return *lo < *hi;
}

// ---- peer-memory transport: one exchange ------------------------------------------------------------------------------
struct PeerSend {
    int peer;
    const void *src;
    size_t dst_off; // in the peer's arena
    size_t bytes;
};
inline unsigned *flag_ptr(uint8_t *arena, const ArenaLayout &lay, size_t index)
{
    return reinterpret_cast<unsigned *>(arena + lay.flags) + index;
}
// Exchange x of the pair with epoch `epoch`: push `sends`, then wait for the ranks in `sources`.
// (do_push / do_wait: the two halves, which ofb_strips_run_phase_device enqueues separately.)
int peer_exchange(ofb_strips *st, int x, const std::vector<PeerSend> &sends, const std::vector<int> &sources, cudaStream_t q,
                  unsigned long long *launches, bool do_push = true, bool do_wait = true)
{
    const int world = st->plan.world, L = st->plan.levels, me = st->rank;
    unsigned *pairs = flag_ptr(st->arena, st->lay, (size_t)(L + 1) * world + L);
    if (do_push && !sends.empty()) {
        PeerPushArgs a{};
        size_t total = 0;
        for (const PeerSend &sd : sends) {
            if (a.nseg == PEER_MAX_SEG || !st->peer_arena[sd.peer]) {
                set_error("strips: exchange %d has too many segments or an unmapped peer %d (strips thinner than the halo?)", x, sd.peer);
                return OFB_ERR_UNSUPPORTED;
            }
            a.seg[a.nseg++] = PeerSeg{sd.src, st->peer_arena[sd.peer] + sd.dst_off, sd.bytes};
            total += sd.bytes;
            bool seen = false;
            for (int d = 0; d < a.ndst; d++) seen |= a.arrive[d] == flag_ptr(st->peer_arena[sd.peer], st->peer_lay[sd.peer], (size_t)x * world + me);
            if (!seen) {
                if (a.ndst == PEER_MAX_DST) {
                    set_error("strips: exchange %d sends to more than %d ranks", x, PEER_MAX_DST);
                    return OFB_ERR_UNSUPPORTED;
                }
                a.arrive[a.ndst] = flag_ptr(st->peer_arena[sd.peer], st->peer_lay[sd.peer], (size_t)x * world + me);
                a.done_local[a.ndst] = flag_ptr(st->arena, st->lay, (size_t)L * world + sd.peer);
                a.ndst++;
                if (std::find(st->done_from.begin(), st->done_from.end(), sd.peer) == st->done_from.end()) st->done_from.push_back(sd.peer);
            }
        }
        a.counter = flag_ptr(st->arena, st->lay, (size_t)(L + 1) * world + x);
        a.pairs = pairs;
        a.err = st->overflow;
        const unsigned blocks = (unsigned)std::min<size_t>(64, std::max<size_t>(1, total / (256 * 64)));
        peer_push_kernel<<<blocks, 256, 0, q>>>(a);
        OFB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
    }
    if (do_wait && !sources.empty()) {
        PeerWaitArgs w{};
        for (int src : sources) {
            if (w.n == PEER_MAX_DST) {
                set_error("strips: exchange %d receives from more than %d ranks", x, PEER_MAX_DST);
                return OFB_ERR_UNSUPPORTED;
            }
            w.flag[w.n++] = flag_ptr(st->arena, st->lay, (size_t)x * world + src);
        }
        w.pairs = pairs;
        w.err = st->overflow;
        peer_wait_kernel<<<1, 32, 0, q>>>(w);
        OFB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
    }
    return OFB_OK;
}
void add_unique(std::vector<int> &v, int x)
{
    if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x);
}
} // namespace

extern "C" {

// The strip schedule as this library computes it, without a device: rows of `level` on `rank` as
// out8 = {y0, y1, by0, by1, cy0, cy1, eb0, eb1} (own rows; rows the level kernel needs; rows of the next-coarser
// cumulative flow held; rows the image buffers hold, which adds what building the coarser levels' halo rows locally
// takes).  Lets a host check its own partitioning -- and the tests check cuda_optical_flow_2_b200/dist.py -- against it.
int ofb_strips_plan_query(int w, int h, int levels, int win, int world, int rank, int reach, int level, int *out8)
{
    if (!out8 || w < 1 || h < 1 || levels < 1 || levels > OFB_MAX_LEVELS || (win & 1) == 0 || win < 3 || win > OFB_MAX_WINDOW ||
        world < 1 || rank < 0 || rank >= world || reach < 0 || level < 0 || level >= levels) {
        set_error("strips_plan_query: bad arguments");
        return OFB_ERR_INVALID;
    }
    const int hc = h >> (levels - 1);
    if (hc < world || (w >> (levels - 1)) < 1) {
        set_error("coarsest level has %d rows, cannot cut it into %d strips", hc, world);
        return OFB_ERR_INVALID;
    }
    StripPlan pl;
    pl.W = w, pl.H = h, pl.levels = levels, pl.win = win, pl.world = world, pl.reach = reach;
    pl.r = win / 2;
    pl.img_halo = pl.r + 2 + reach + 2;
    for (int rk = 0; rk < world; rk++) pl.coarse_bounds.push_back((int)(((long long)rk * hc) / world));
    pl.coarse_bounds.push_back(hc);
    const LevelStrip s = pl.strips(rank)[level];
    const int v[8] = {s.y0, s.y1, s.by0, s.by1, s.cy0, s.cy1, s.eb0, s.eb1};
    memcpy(out8, v, sizeof v);
    return OFB_OK;
}

int ofb_strips_nccl_unique_id(void *id128)
{
    NcclApi *n = nccl();
    if (!n) {
        set_error("libnccl.so.2 could not be loaded");
        return OFB_ERR_UNSUPPORTED;
    }
    if (!id128) {
        set_error("NULL id buffer");
        return OFB_ERR_INVALID;
    }
    NcclUniqueId id;
    OFB_NCCL_TRY(n->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return OFB_OK;
}

int ofb_strips_create(ofb_ctx *ctx, int w, int h, int levels, int win, int warp_mode, float flow_scale, int world, int rank,
                      int reach, const void *nccl_id128, ofb_strips **out)
{
    if (!ctx || !out || w < 1 || h < 1 || levels < 1 || levels > OFB_MAX_LEVELS || (win & 1) == 0 || win < 3 ||
        win > OFB_MAX_WINDOW || world < 1 || rank < 0 || rank >= world || reach < 0) {
        set_error("strips_create: bad arguments");
        return OFB_ERR_INVALID;
    }
    if (warp_mode == OFB_WARP_AS_WRITTEN && levels > 1) {
        set_error("OFB_WARP_AS_WRITTEN needs pixel (0,0) of every coarser level: not available on strips");
        return OFB_ERR_UNSUPPORTED;
    }
    const int hc = h >> (levels - 1);
    if (hc < world || (w >> (levels - 1)) < 1) {
        set_error("coarsest level has %d rows, cannot cut it into %d strips", hc, world);
        return OFB_ERR_INVALID;
    }
    DeviceGuard guard(ctx_device(ctx));
    if (!guard.ok) return OFB_ERR_CUDA;
    ofb_strips *st = new (std::nothrow) ofb_strips;
    if (!st) {
        set_error("strips_create: out of host memory");
        return OFB_ERR_NOMEM;
    }
    try { // (std::vector growth below must not throw across the C ABI)
    st->ctx = ctx;
    st->rank = rank;
    st->warp_mode = warp_mode;
    st->flow_scale = flow_scale;
    StripPlan &pl = st->plan;
    pl.W = w, pl.H = h, pl.levels = levels, pl.win = win, pl.world = world, pl.reach = reach;
    pl.r = win / 2;
    // image halo: 3x3 stencil + window radius + one row of even-row alignment + warp reach + bilinear tap
    pl.img_halo = pl.r + 2 + reach + 2;
    for (int rk = 0; rk < world; rk++) pl.coarse_bounds.push_back((int)(((long long)rk * hc) / world));
    pl.coarse_bounds.push_back(hc);
    int rc = OFB_OK;
    st->s = pl.strips(rank);
    for (int peer = 0; peer < world; peer++) st->peer0.push_back(pl.strips(peer)[0]);
    // what neighbours write into (level-0 image rows, coarser cumulative-flow rows, flags): one allocation, the arena
    st->lay = arena_layout(pl, rank);
    {
        void *p = nullptr;
        rc = dev_alloc(st, &p, st->lay.bytes);
        st->arena = static_cast<uint8_t *>(p);
    }
    for (int k = 0; k < levels && rc == OFB_OK; k++) {
        const LevelStrip s = st->s[k];
        const size_t pitch = ((size_t)s.w + 63) / 64 * 64;
        st->pitch.push_back(pitch);
        void *p = nullptr;
        const int rows = s.eb1 - s.eb0;
        if (k == 0) {
            st->prev.push_back(st->arena + st->lay.prev0);
            st->next.push_back(st->arena + st->lay.next0);
            st->img_stride.push_back(st->lay.next0 - st->lay.prev0);
        } else { // one block: the pyramid step handles prev and next as a batch of two
            const size_t stride = (pitch * rows + 255) / 256 * 256;
            if ((rc = dev_alloc(st, &p, 2 * stride))) break;
            st->prev.push_back(static_cast<uint8_t *>(p));
            st->next.push_back(static_cast<uint8_t *>(p) + stride);
            st->img_stride.push_back(stride);
        }
        if ((rc = dev_alloc(st, &p, (size_t)s.w * rows * 8))) break;
        st->flow.push_back(static_cast<float *>(p));
        // the cumulative flow of level 0 (the total flow) has its own buffer; levels >= 1 write theirs straight into the
        // buffer the next finer level reads (below), whose other rows the neighbours fill
        p = nullptr;
        if (k == 0 && (rc = dev_alloc(st, &p, (size_t)s.w * rows * 8))) break;
        st->cum.push_back(static_cast<float *>(p));
        st->cum_in.push_back(reinterpret_cast<float *>(st->arena + st->lay.cum_in[k]));
    }
    for (int k = 0; k < levels && rc == OFB_OK; k++) {
        // buffer row b of level k is global row eb0 + b; row g of cum_in[k-1] sits at (g - cy0) * w
        float *dst = st->cum[k];
        if (k > 0) {
            const LevelStrip &s = st->s[k], &fine = st->s[k - 1];
            dst = st->cum_in[k - 1] + ((ptrdiff_t)s.eb0 - fine.cy0) * (ptrdiff_t)s.w * 2;
            if (s.y0 < fine.cy0 || s.y1 > fine.cy1) {
                set_error("strips_create: level %d own rows [%d,%d) outside the finer level's coarse window [%d,%d)", k, s.y0, s.y1,
                          fine.cy0, fine.cy1);
                rc = OFB_ERR_INVALID;
            }
        }
        st->cum_dst.push_back(dst);
    }
    if (rc == OFB_OK) {
        void *p = nullptr;
        rc = dev_alloc(st, &p, sizeof(int));
        st->overflow = static_cast<int *>(p);
    }
    if (rc == OFB_OK && world > 1 && nccl_id128) {
        NcclApi *n = nccl();
        if (!n) {
            set_error("libnccl.so.2 could not be loaded");
            rc = OFB_ERR_UNSUPPORTED;
        } else {
            NcclUniqueId id;
            memcpy(&id, nccl_id128, sizeof id);
            const int r = n->CommInitRank(&st->comm, world, id, rank);
            if (r != NCCL_SUCCESS) {
                set_error("ncclCommInitRank failed: %s", n->GetErrorString(r));
                rc = OFB_ERR_CUDA;
            }
        }
    }
    // every kernel a pair launches is loaded now: a lazily loaded kernel may synchronise the context at its first
    // launch, which must not happen behind a kernel that waits for a neighbour
    if (rc == OFB_OK) rc = preload_lk_level(win);
    if (rc == OFB_OK) rc = preload_pyramid();
    if (rc == OFB_OK) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, peer_push_kernel) != cudaSuccess || cudaFuncGetAttributes(&fa, peer_wait_kernel) != cudaSuccess ||
            cudaFuncGetAttributes(&fa, peer_done_kernel) != cudaSuccess) {
            set_error("strips_create: loading the exchange kernels failed");
            rc = OFB_ERR_CUDA;
        }
    }
    if (rc == OFB_OK && cudaDeviceSynchronize() != cudaSuccess) { // the zeroed flags must be in place before a peer can see the arena
        set_error("strips_create: device synchronisation failed");
        rc = OFB_ERR_CUDA;
    }
    if (rc != OFB_OK) {
        for (void *p : st->allocs) cudaFree(p);
        delete st;
        return rc;
    }
    } catch (const std::bad_alloc &) {
        for (void *p : st->allocs) cudaFree(p);
        delete st;
        set_error("strips_create: out of host memory");
        return OFB_ERR_NOMEM;
    }
    *out = st;
    return OFB_OK;
}

// ---- peer-memory transport: set-up -------------------------------------------------------------------------------
// The 128-byte blob a rank publishes: its arena's CUDA IPC handle and the arena's offset inside the allocation the
// handle stands for (small cudaMalloc blocks share one).
int ofb_strips_peer_handle(ofb_strips *st, void *blob128)
{
    if (!st || !blob128) {
        set_error("strips_peer_handle: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_STRIPS_GUARD(st);
    memset(blob128, 0, 128);
    cudaIpcMemHandle_t h;
    OFB_CUDA_TRY(cudaIpcGetMemHandle(&h, st->arena));
    static_assert(sizeof h == 64, "CUDA IPC handle size");
    memcpy(blob128, &h, sizeof h);
    unsigned long long off = 0;
    {
        typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        unsigned long long base = 0;
        size_t size = 0;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && sym &&
            reinterpret_cast<GetRange>(sym)(&base, &size, (unsigned long long)reinterpret_cast<uintptr_t>(st->arena)) == 0)
            off = (unsigned long long)reinterpret_cast<uintptr_t>(st->arena) - base;
        else {
            set_error("strips_peer_handle: cuMemGetAddressRange failed");
            return OFB_ERR_CUDA;
        }
    }
    memcpy(static_cast<char *>(blob128) + 64, &off, sizeof off);
    return OFB_OK;
}

int ofb_strips_peer_arena(ofb_strips *st, void **arena)
{
    if (!st || !arena) {
        set_error("strips_peer_arena: bad arguments");
        return OFB_ERR_INVALID;
    }
    *arena = st->arena;
    return OFB_OK;
}

namespace {
// ranks whose buffers I write into, or that write into mine, at any exchange
bool is_neighbour(const ofb_strips *st, int peer)
{
    const StripPlan &pl = st->plan;
    int lo, hi;
    const LevelStrip &s = st->s[0], &ps = st->peer0[peer];
    if (overlap(s.y0, s.y1, ps.eb0, ps.eb1, &lo, &hi) || overlap(ps.y0, ps.y1, s.eb0, s.eb1, &lo, &hi)) return true;
    for (int k = 0; k + 1 < pl.levels; k++) {
        const LevelStrip up = st->s[k + 1], sk = st->s[k], pk = pl.level(k, peer), pu = pl.level(k + 1, peer);
        if (overlap(up.y0, up.y1, pk.cy0, pk.cy1, &lo, &hi) || overlap(pu.y0, pu.y1, sk.cy0, sk.cy1, &lo, &hi)) return true;
    }
    return false;
}
int finish_connect(ofb_strips *st)
{
    st->peer_lay.clear();
    for (int peer = 0; peer < st->plan.world; peer++) st->peer_lay.push_back(arena_layout(st->plan, peer));
    st->peers_connected = true;
    return OFB_OK;
}
} // namespace

// blobs: world x 128 bytes, rank r's ofb_strips_peer_handle blob at 128 * r (an all-gather by the application).
int ofb_strips_peer_connect(ofb_strips *st, const void *blobs)
{
    if (!st || !blobs) {
        set_error("strips_peer_connect: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_STRIPS_GUARD(st);
    const int world = st->plan.world;
    st->peer_arena.assign(world, nullptr);
    st->peer_arena[st->rank] = st->arena;
    for (int peer = 0; peer < world; peer++) {
        if (peer == st->rank || !is_neighbour(st, peer)) continue;
        cudaIpcMemHandle_t h;
        unsigned long long off = 0;
        memcpy(&h, static_cast<const char *>(blobs) + 128 * (size_t)peer, sizeof h);
        memcpy(&off, static_cast<const char *>(blobs) + 128 * (size_t)peer + 64, sizeof off);
        void *base = nullptr;
        OFB_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        st->ipc_opened.push_back(base);
        st->peer_arena[peer] = static_cast<uint8_t *>(base) + off;
    }
    return finish_connect(st);
}

// Same, for ranks that live in ONE process (tests; a single-process multi-GPU host with peer access enabled by the
// caller): arenas[r] = ofb_strips_peer_arena of rank r.
int ofb_strips_peer_connect_local(ofb_strips *st, void *const *arenas)
{
    if (!st || !arenas) {
        set_error("strips_peer_connect_local: bad arguments");
        return OFB_ERR_INVALID;
    }
    st->peer_arena.assign(st->plan.world, nullptr);
    for (int peer = 0; peer < st->plan.world; peer++) st->peer_arena[peer] = static_cast<uint8_t *>(arenas[peer]);
    st->peer_arena[st->rank] = st->arena;
    return finish_connect(st);
}

// Teardown.  The last thing a neighbour writes into this rank's arena is its `done` flag at the end of its last pair,
// which can land after this rank has finished its own: so, before the arena is unmapped and freed, wait (on the device,
// bounded like every other wait) until every rank this one pushes to has acknowledged this rank's last pair.  Ranks must
// have run the same number of pairs; a rank that is still PUSHING a further pair into a freed arena is a caller error.
int ofb_strips_destroy(ofb_strips *st)
{
    if (!st) return OFB_OK;
    {
        DeviceGuard guard(ctx_device(st->ctx));
        if (guard.ok) {
            cudaDeviceSynchronize();
            if (st->plan.world > 1 && st->peers_connected && !st->comm && !st->done_from.empty()) {
                PeerDrainArgs d{};
                const int L = st->plan.levels, world = st->plan.world;
                for (int dst : st->done_from)
                    if (d.n < 2 * PEER_MAX_DST) d.flag[d.n++] = flag_ptr(st->arena, st->lay, (size_t)L * world + dst);
                d.pairs = flag_ptr(st->arena, st->lay, (size_t)(L + 1) * world + L);
                d.err = st->overflow;
                peer_drain_kernel<<<1, 32>>>(d);
                cudaDeviceSynchronize();
            }
            for (void *p : st->ipc_opened) cudaIpcCloseMemHandle(p);
            if (st->comm && nccl()) nccl()->CommDestroy(st->comm);
            for (void *p : st->allocs) cudaFree(p);
        }
    }
    delete st;
    return OFB_OK;
}

int ofb_strips_own_rows(const ofb_strips *st, int level, int *y0, int *y1)
{
    if (!st || level < 0 || level >= st->plan.levels) {
        set_error("strips_own_rows: bad arguments");
        return OFB_ERR_INVALID;
    }
    if (y0) *y0 = st->s[level].y0;
    if (y1) *y1 = st->s[level].y1;
    return OFB_OK;
}

int ofb_strips_result(const ofb_strips *st, int level, float **flow_own_d, float **total_own_d)
{
    if (!st || level < 0 || level >= st->plan.levels) {
        set_error("strips_result: bad arguments");
        return OFB_ERR_INVALID;
    }
    const LevelStrip &s = st->s[level];
    const size_t off = (size_t)(s.y0 - s.eb0) * s.w * 2;
    if (flow_own_d) *flow_own_d = st->flow[level] + off;
    // cumulative flow exists where a level writes it: every level but the coarsest (whose cumulative flow is its flow)
    if (total_own_d) *total_own_d = (st->plan.levels == 1 ? st->flow[level] : st->cum_dst[level]) + off;
    return OFB_OK;
}

// Where the own rows [y0, y1) of level 0 live: a producer that writes them there and passes the same pointers to
// ofb_strips_run_device saves the upload copy.
int ofb_strips_input(const ofb_strips *st, uint8_t **prev_own_d, uint8_t **next_own_d, size_t *pitch)
{
    if (!st) {
        set_error("strips_input: NULL handle");
        return OFB_ERR_INVALID;
    }
    const LevelStrip &s = st->s[0];
    const size_t off = (size_t)(s.y0 - s.eb0) * st->pitch[0];
    if (prev_own_d) *prev_own_d = st->prev[0] + off;
    if (next_own_d) *next_own_d = st->next[0] + off;
    if (pitch) *pitch = st->pitch[0];
    return OFB_OK;
}

// Whether level 0 also writes the total (cumulative) flow of the pair; on by default, 8 more bytes per pixel.
int ofb_strips_set_total(ofb_strips *st, int on)
{
    if (!st) {
        set_error("strips_set_total: NULL handle");
        return OFB_ERR_INVALID;
    }
    st->want_total = on ? 1 : 0;
    return OFB_OK;
}

// Whether the cumulative-flow halo rows are pushed by the level kernels themselves (default) or by separate copy / wait
// kernels (peer-memory transport; for comparison).
int ofb_strips_set_fused(ofb_strips *st, int on)
{
    if (!st) {
        set_error("strips_set_fused: NULL handle");
        return OFB_ERR_INVALID;
    }
    st->fused = on ? 1 : 0;
    return OFB_OK;
}

// One pair: own rows of level 0 in, residual (and cumulative) flow of the own rows of every level out, all
// asynchronous on `stream`.  prev_own_d / next_own_d: rows [y0, y1) of level 0, planar u8 with `pitch` bytes per row.
//
// `phase` < 0: the whole pair.  Otherwise one of its 2L phases (peer-memory transport only), so that a host can interleave
// the ranks of ONE process on ONE stream in an order in which no kernel ever waits for a later one:
//   phase 0      own-row upload, push of the level-0 image rows (exchange 0)
//   phase 1      wait for exchange 0, pyramid, coarsest level L-1
//   phase 2j     push of the cumulative-flow rows of level L-j (exchange j),            j = 1 .. L-1
//   phase 2j+1   wait for exchange j, level L-1-j; after level 0 the end-of-pair handshake
static int strips_run_impl(ofb_strips *st, const uint8_t *prev_own_d, const uint8_t *next_own_d, size_t pitch, void *stream, int phase)
{
    if (!st || !prev_own_d || !next_own_d) {
        set_error("strips_run: bad arguments");
        return OFB_ERR_INVALID;
    }
    OFB_STRIPS_GUARD(st);
    try { // (std::vector growth must not throw across the C ABI)
    auto on = [phase](int ph) { return phase < 0 || phase == ph; };
    cudaStream_t q = static_cast<cudaStream_t>(stream);
    const StripPlan &pl = st->plan;
    const bool peer_mode = pl.world > 1 && !st->comm;
    if (peer_mode && !st->peers_connected) {
        set_error("strips_run: no transport (create with a NCCL id, or call ofb_strips_peer_connect first)");
        return OFB_ERR_INVALID;
    }
    NcclApi *n = pl.world > 1 && !peer_mode ? nccl() : nullptr;
    if (phase >= 0 && (n || phase >= 2 * pl.levels)) {
        set_error("strips_run_phase: phase %d of %d (phases exist for the peer-memory transport and one rank only)", phase, 2 * pl.levels);
        return OFB_ERR_INVALID;
    }
    unsigned long long *launches = ctx_launch_counter(st->ctx);
    const int L = pl.levels, me = st->rank;
    std::vector<int> all_sources; // every rank that pushes to me during the pair
    if (pitch < (size_t)pl.W) {
        set_error("strips_run: pitch %zu smaller than the width %d", pitch, pl.W);
        return OFB_ERR_INVALID;
    }
    if (on(0)) {
        const LevelStrip &s = st->s[0];
        const size_t off = (size_t)(s.y0 - s.eb0) * st->pitch[0];
        // a producer that wrote the own rows in place (ofb_strips_input) saves this copy
        if (prev_own_d != st->prev[0] + off)
            OFB_CUDA_TRY(cudaMemcpy2DAsync(st->prev[0] + off, st->pitch[0], prev_own_d, pitch, (size_t)s.w, (size_t)(s.y1 - s.y0),
                                           cudaMemcpyDeviceToDevice, q));
        if (next_own_d != st->next[0] + off)
            OFB_CUDA_TRY(cudaMemcpy2DAsync(st->next[0] + off, st->pitch[0], next_own_d, pitch, (size_t)s.w, (size_t)(s.y1 - s.y0),
                                           cudaMemcpyDeviceToDevice, q));
    }
    // (the error flag st->overflow is sticky: kernels OR into it, only ofb_strips_check clears it)
    // ONE image exchange, on level 0: the rows of my buffer that other ranks own (and vice versa), in place
    if (n) {
        const LevelStrip &s = st->s[0];
        const size_t P = st->pitch[0];
        OFB_NCCL_TRY(n->GroupStart());
        for (int peer = 0; peer < pl.world; peer++) {
            if (peer == me) continue;
            const LevelStrip &ps = st->peer0[peer];
            int lo, hi;
            if (overlap(s.y0, s.y1, ps.eb0, ps.eb1, &lo, &hi)) { // my rows the peer's buffer holds
                OFB_NCCL_TRY(n->Send(st->prev[0] + (size_t)(lo - s.eb0) * P, (size_t)(hi - lo) * P, NCCL_UINT8, peer, st->comm, q));
                OFB_NCCL_TRY(n->Send(st->next[0] + (size_t)(lo - s.eb0) * P, (size_t)(hi - lo) * P, NCCL_UINT8, peer, st->comm, q));
            }
            if (overlap(ps.y0, ps.y1, s.eb0, s.eb1, &lo, &hi)) { // the peer's rows my buffer holds
                OFB_NCCL_TRY(n->Recv(st->prev[0] + (size_t)(lo - s.eb0) * P, (size_t)(hi - lo) * P, NCCL_UINT8, peer, st->comm, q));
                OFB_NCCL_TRY(n->Recv(st->next[0] + (size_t)(lo - s.eb0) * P, (size_t)(hi - lo) * P, NCCL_UINT8, peer, st->comm, q));
            }
        }
        OFB_NCCL_TRY(n->GroupEnd());
    }
    if (peer_mode) {
        const LevelStrip &s = st->s[0];
        const size_t P = st->pitch[0];
        std::vector<PeerSend> sends;
        std::vector<int> sources;
        for (int peer = 0; peer < pl.world; peer++) {
            if (peer == me) continue;
            const LevelStrip &ps = st->peer0[peer];
            int lo, hi;
            if (overlap(s.y0, s.y1, ps.eb0, ps.eb1, &lo, &hi)) { // my rows the peer's buffer holds
                const ArenaLayout &pa = st->peer_lay[peer];
                sends.push_back({peer, st->prev[0] + (size_t)(lo - s.eb0) * P, pa.prev0 + (size_t)(lo - ps.eb0) * P, (size_t)(hi - lo) * P});
                sends.push_back({peer, st->next[0] + (size_t)(lo - s.eb0) * P, pa.next0 + (size_t)(lo - ps.eb0) * P, (size_t)(hi - lo) * P});
            }
            if (overlap(ps.y0, ps.y1, s.eb0, s.eb1, &lo, &hi)) { // the peer's rows my buffer holds
                sources.push_back(peer);
                add_unique(all_sources, peer);
            }
        }
        int rc = peer_exchange(st, 0, sends, sources, q, launches, on(0), on(1));
        if (rc) return rc;
    }
    // pyramid: every buffer row of level k+1 from the buffer rows of level k (halo rows are built locally)
    for (int k = 0; on(1) && k + 1 < L; k++) {
        const LevelStrip &s = st->s[k], &d = st->s[k + 1];
        int rc = launch_pyr_down_strip(st->prev[k], st->pitch[k], s.w, s.eb1 - s.eb0, s.eb0, st->prev[k + 1], st->pitch[k + 1], d.eb0,
                                       d.eb1, q, launches, 2, st->img_stride[k], st->img_stride[k + 1]);
        if (rc) return rc;
    }
    // coarse to fine: rows [cy0, cy1) of cum_{k+1} into cum_in[k] (own part copied, the rest received), then the level
    bool pushed_by_kernel = false; // the exchange that feeds the level at hand was pushed by the previous (coarser) level's kernel
    for (int k = L - 1; k >= 0; k--) {
        const LevelStrip &s = st->s[k];
        struct WaitSrc {
            int peer, crow_lo, crow_hi;
        } wait_src[2] = {};
        int nwait = 0;
        bool wait_in_kernel = false;
        if (k < L - 1) {
            const LevelStrip &up = st->s[k + 1];
            // level k+1 wrote its own rows of cum_{k+1} straight into cum_in[k] (cum_dst); the neighbours' rows follow
            const float *src = st->cum_dst[k + 1]; // buffer-row indexing of level k+1 (origin up.eb0)
            const size_t rowf = (size_t)up.w * 2;  // floats per row
            int lo, hi;
            if (n) {
                OFB_NCCL_TRY(n->GroupStart());
                for (int peer = 0; peer < pl.world; peer++) {
                    if (peer == me) continue;
                    const LevelStrip pk = pl.level(k, peer), pu = pl.level(k + 1, peer);
                    if (overlap(up.y0, up.y1, pk.cy0, pk.cy1, &lo, &hi)) // my cum rows the peer needs
                        OFB_NCCL_TRY(n->Send(src + (size_t)(lo - up.eb0) * rowf, (size_t)(hi - lo) * rowf * 4, NCCL_UINT8, peer, st->comm, q));
                    if (overlap(pu.y0, pu.y1, s.cy0, s.cy1, &lo, &hi)) // the peer's cum rows I need
                        OFB_NCCL_TRY(n->Recv(st->cum_in[k] + (size_t)(lo - s.cy0) * rowf, (size_t)(hi - lo) * rowf * 4, NCCL_UINT8, peer,
                                             st->comm, q));
                }
                OFB_NCCL_TRY(n->GroupEnd());
            }
            if (peer_mode) {
                std::vector<PeerSend> sends;
                std::vector<int> sources;
                for (int peer = 0; peer < pl.world; peer++) {
                    if (peer == me) continue;
                    const LevelStrip pk = pl.level(k, peer), pu = pl.level(k + 1, peer);
                    if (overlap(up.y0, up.y1, pk.cy0, pk.cy1, &lo, &hi)) // my cum rows the peer needs
                        sends.push_back({peer, src + (size_t)(lo - up.eb0) * rowf, st->peer_lay[peer].cum_in[k] + (size_t)(lo - pk.cy0) * rowf * 4,
                                         (size_t)(hi - lo) * rowf * 4});
                    if (overlap(pu.y0, pu.y1, s.cy0, s.cy1, &lo, &hi)) { // the peer's cum rows I need
                        sources.push_back(peer);
                        add_unique(all_sources, peer);
                        if (nwait < 2) wait_src[nwait] = {peer, lo - s.cy0, hi - s.cy0};
                        nwait++;
                    }
                }
                const int x = L - 1 - k; // exchanges 1 .. L-1
                // Fused form: level k+1's kernel has already pushed my rows (pushed_by_kernel, previous iteration) and this
                // level's kernel waits for the neighbours' flags itself; the copy / wait kernels remain for what does not
                // fit the kernel's two slots (strips thinner than the halo).
                wait_in_kernel = st->fused && nwait >= 1 && nwait <= 2;
                int rc = peer_exchange(st, x, sends, sources, q, launches, on(2 * x) && !pushed_by_kernel, on(2 * x + 1) && !wait_in_kernel);
                if (rc) return rc;
            }
        }
        // what THIS level's kernel pushes: my rows of cum_k that the neighbours' level k-1 needs (exchange L-k)
        LkLevelArgs::PeerPush push[2] = {};
        int npush = 0;
        pushed_by_kernel = false;
        if (peer_mode && st->fused && k >= 1) {
            int cnt = 0, lo, hi;
            for (int peer = 0; peer < pl.world; peer++) {
                if (peer == me) continue;
                const LevelStrip pf = pl.level(k - 1, peer); // the peer's finer level reads rows [cy0, cy1) of cum_k
                if (!overlap(s.y0, s.y1, pf.cy0, pf.cy1, &lo, &hi)) continue;
                if (cnt < 2 && st->peer_arena[peer]) {
                    const int xn = L - k;
                    push[cnt].dst = reinterpret_cast<float *>(st->peer_arena[peer] + st->peer_lay[peer].cum_in[k - 1]) +
                                    ((ptrdiff_t)s.eb0 - pf.cy0) * (ptrdiff_t)s.w * 2;
                    push[cnt].row_lo = lo - s.eb0;
                    push[cnt].row_hi = hi - s.eb0;
                    push[cnt].flag = flag_ptr(st->peer_arena[peer], st->peer_lay[peer], (size_t)xn * pl.world + me);
                    push[cnt].done = flag_ptr(st->arena, st->lay, (size_t)L * pl.world + peer);
                    if (std::find(st->done_from.begin(), st->done_from.end(), peer) == st->done_from.end()) st->done_from.push_back(peer);
                }
                cnt++;
            }
            if (cnt >= 1 && cnt <= 2) {
                npush = cnt;
                pushed_by_kernel = true;
            }
        }
        if (!on(2 * (L - 1 - k) + 1)) continue;
        LkLevelArgs a{};
        a.prev = st->prev[k];
        a.next = st->next[k];
        a.pitch = st->pitch[k];
        a.image_stride = a.pitch * (size_t)(s.eb1 - s.eb0);
        a.w = s.w;
        a.h_local = s.eb1 - s.eb0;
        a.y_off = s.eb0;
        a.h_global = s.h;
        a.out_y0 = s.y0 - s.eb0;
        a.out_y1 = s.y1 - s.eb0;
        a.n_pairs = 1;
        a.win = pl.win;
        a.warp_mode = st->warp_mode;
        a.flow_scale = st->flow_scale;
        a.cum_in = k < L - 1 ? st->cum_in[k] : nullptr;
        a.cum_w = s.w >> 1;
        a.cum_h_global = s.h >> 1;
        a.cum_y_off = s.cy0;
        a.cum_h_local = std::max(s.cy1 - s.cy0, 1);
        a.cum_pair_stride = 0;
        a.flow_out = st->flow[k];
        const bool want_cum = k > 0 || (L > 1 && st->want_total); // every level but the finest feeds the next one
        a.cum_out = want_cum ? st->cum_dst[k] : nullptr;
        a.flow_pair_stride = 0;
        a.reach_overflow = st->overflow;
        a.sm_count = ctx_sm_count(st->ctx);
        a.solve_fast = ctx_solve_fast(st->ctx);
        if (npush > 0 && a.cum_out) {
            a.npush = npush;
            a.push[0] = push[0];
            a.push[1] = push[1];
            a.push_counter = flag_ptr(st->arena, st->lay, (size_t)(L + 1) * pl.world + L + 1 + 2 * (size_t)(L - k));
        }
        if (wait_in_kernel) {
            a.nwait = nwait;
            const int x = L - 1 - k;
            for (int d = 0; d < nwait; d++)
                a.wait[d] = {flag_ptr(st->arena, st->lay, (size_t)x * pl.world + wait_src[d].peer), wait_src[d].crow_lo, wait_src[d].crow_hi};
        }
        a.epoch_src = flag_ptr(st->arena, st->lay, (size_t)(L + 1) * pl.world + L);
        int rc = launch_lk_level(a, q, launches);
        if (rc) return rc;
    }
    if (peer_mode && on(2 * L - 1)) { // the rows my neighbours pushed have been used: they may push the next pair's
        PeerDoneArgs d{};
        for (int src : all_sources) {
            if (d.n == 2 * PEER_MAX_DST) {
                set_error("strips: more than %d ranks push to this one", 2 * PEER_MAX_DST);
                return OFB_ERR_UNSUPPORTED;
            }
            d.flag[d.n++] = flag_ptr(st->peer_arena[src], st->peer_lay[src], (size_t)L * pl.world + me);
        }
        d.pairs = flag_ptr(st->arena, st->lay, (size_t)(L + 1) * pl.world + L);
        peer_done_kernel<<<1, 32, 0, q>>>(d);
        OFB_CUDA_TRY(cudaGetLastError());
        if (launches) ++*launches;
    }
    } catch (const std::bad_alloc &) {
        set_error("strips_run: out of host memory");
        return OFB_ERR_NOMEM;
    }
    return OFB_OK;
}

int ofb_strips_run_device(ofb_strips *st, const uint8_t *prev_own_d, const uint8_t *next_own_d, size_t pitch, void *stream)
{
    return strips_run_impl(st, prev_own_d, next_own_d, pitch, stream, -1);
}

int ofb_strips_run_phase_device(ofb_strips *st, const uint8_t *prev_own_d, const uint8_t *next_own_d, size_t pitch, int phase,
                                void *stream)
{
    if (phase < 0) {
        set_error("strips_run_phase: phase %d", phase);
        return OFB_ERR_INVALID;
    }
    return strips_run_impl(st, prev_own_d, next_own_d, pitch, stream, phase);
}

// Synchronises the stream and reports whether a warp sample reached past the exchanged halo rows (the result is
// then NOT the whole-frame result: raise `reach`).
int ofb_strips_check(ofb_strips *st, void *stream, int *overflow)
{
    if (!st) {
        set_error("strips_check: NULL handle");
        return OFB_ERR_INVALID;
    }
    OFB_STRIPS_GUARD(st);
    int v = 0;
    // sticky flag: everything since the last check; read, then clear (in stream order, so pairs enqueued later start clean)
    OFB_CUDA_TRY(cudaMemcpyAsync(&v, st->overflow, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    OFB_CUDA_TRY(cudaMemsetAsync(st->overflow, 0, sizeof(int), static_cast<cudaStream_t>(stream)));
    OFB_CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (overflow) *overflow = v;
    return OFB_OK;
}

} // extern "C"

// pyramid.cu -- one Gaussian pyramid step: 3x3 binomial centred on (2x, 2y), decimate by 2.
// Replaces g_gauss_pyramid / gauss_pyramid_level (reference OptFlowGpu.cu:1198-1232, 1235-1259):
//   dst(x,y) = (unsigned char)( sum_{p,q} G[p][q] * src(2x-1+q, 2y-1+p) ),  G = [1 2 1;2 4 2;1 2 1]/16,
// taps left of / above the image skipped (no renormalisation).  With integer weights the fp32
// sum of the reference is exact, so (sum w_i p_i) >> 4 is byte-identical.
#include "ofb_common.cuh"

#include <cstdlib>

namespace ofb {

// Planar u8.  One thread = 8 output columns x 2 output rows: five source rows of 16 bytes (one
// 128-bit load each; the byte left of them comes from the neighbouring lane by shuffle), horizontal
// [1 2 1] on packed 16-bit pairs (values stay below 2^12), vertical [1 2 1] shared between the two
// output rows, one 64-bit store per row.  Row strips: dst local row y is global row dst_y0 + y, src
// local row 0 is global row src_y_off.
__device__ __forceinline__ void pyr_hrow(const uint4 q, uint32_t left, uint32_t h[4])
{
    // bytes b0..b15; h[k] (k = 0..7, packed two per word) = b[2k-1] + 2*b[2k] + b[2k+1], b[-1] = left
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t prev_odd = left << 16; // odd byte of the previous pair in the high half
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t ev = __byte_perm(w[i], 0, 0x4240); // b0 | b2 << 16
        const uint32_t od = __byte_perm(w[i], 0, 0x4341); // b1 | b3 << 16
        const uint32_t odl = __funnelshift_l(prev_odd, od, 16); // (previous odd byte) | b1 << 16
        h[i] = odl + 2 * ev + od;
        prev_odd = od;
    }
}

#ifndef PYR_ROWS
#define PYR_ROWS 2 // output rows per thread
#endif
__global__ void __launch_bounds__(256)
pyr_down_planar_kernel(const uint8_t *__restrict__ src, size_t src_pitch, size_t src_stride, int dw, int dh,
                       uint8_t *__restrict__ dst, size_t dst_pitch, size_t dst_stride, int src_y_off, int dst_y0,
                       int src_rows, const uint8_t *__restrict__ src2, uint8_t *__restrict__ dst2, int n_first)
{
    pdl_launch_dependents(); // (programmatic dependent launch, ofb_common.cuh)
    pdl_wait();
    // images n_first .. of the launch form a second batch with its own base pointers (prev and next frames of the
    // pairs in one launch)
    const int zi = (int)blockIdx.z < n_first ? (int)blockIdx.z : (int)blockIdx.z - n_first;
    if ((int)blockIdx.z >= n_first) {
        src = src2;
        dst = dst2;
    }
    const int tx = blockIdx.x * 32 + threadIdx.x;
    const int x0 = 8 * tx;
    constexpr int NR = PYR_ROWS;
    const int y = NR * (blockIdx.y * 8 + threadIdx.y);
    // whole warps stay alive for the shuffle; lanes past the row only skip their loads and stores
    const bool live = x0 < dw && y < dh;
    const uint8_t *s = src + (size_t)zi * src_stride;
    uint32_t h[2 * NR + 1][4];
#pragma unroll
    for (int r = 0; r < 2 * NR + 1; r++) {
        const int sg = 2 * (dst_y0 + y) - 1 + r; // global source row; above the image: skipped
        const int sy = sg - src_y_off;
        const bool ok = live && sg >= 0 && sy < src_rows && (r < 3 || y + (r - 1) / 2 < dh);
        uint4 q = make_uint4(0, 0, 0, 0);
        if (ok) q = __ldg(reinterpret_cast<const uint4 *>(s + (size_t)sy * src_pitch + 2 * x0));
        uint32_t left = __shfl_up_sync(0xffffffffu, q.w >> 24, 1);
        if (threadIdx.x == 0) left = (ok && tx > 0) ? (uint32_t)__ldg(s + (size_t)sy * src_pitch + 2 * x0 - 1) : 0u;
        pyr_hrow(q, left, h[r]);
    }
    if (!live) return;
#pragma unroll
    for (int o = 0; o < NR; o++) {
        if (y + o >= dh) break;
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = ((h[2 * o][i] + 2 * h[2 * o + 1][i] + h[2 * o + 2][i]) >> 4) & 0x00ff00ffu;
        // v[i] holds outputs 2i (low half) and 2i+1 (high half), each already < 256
        const uint32_t lo = __byte_perm(v[0], v[1], 0x6420), hi = __byte_perm(v[2], v[3], 0x6420);
        uint8_t *d = dst + (size_t)zi * dst_stride + (size_t)(y + o) * dst_pitch + x0;
        if (x0 + 7 < dw && (reinterpret_cast<uintptr_t>(d) & 7) == 0) {
            *reinterpret_cast<uint2 *>(d) = make_uint2(lo, hi);
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (x0 + k < dw) d[k] = (uint8_t)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 255u);
        }
    }
}

// ---- whole frames: a thread walks DOWN its 8 output columns, one or two pyramid steps at once --------------------------
// The kernel above spends ~28 instructions per output byte on addresses and predicates (ncu, 1080p: 58 % of the issue
// slots for 54 % of the DRAM rate -- bound by neither).  Here a thread keeps its 16 source columns and slides down `rpt`
// output rows: per output row two new source rows (one 128-bit load each, issued one iteration ahead), their horizontal
// [1 2 1] sums on packed 16-bit pairs, the vertical [1 2 1] with the row kept from the previous step, one 64-bit store.
// TWO: the next pyramid level is formed on the way from the level-1 bytes the thread has just produced (same packed
// arithmetic; the byte left of its columns comes from the neighbouring lane), so level 1 is never read back from memory:
// 2.07 MB read + 0.65 MB written per 1080p frame for two levels instead of 3.24 MB.
// Lane 0 of every warp recomputes the last 8 columns of the warp to its left and stores nothing: it only feeds lane 1's
// left neighbours (source byte, level-1 byte), which keeps the image's left border (taps left of column 0 are skipped =
// zero) and the warp seams out of the inner loop.
struct PyrRollParams {
    const uint8_t *src[2]; // two batches of n_first / (gridDim.z - n_first) images with the same geometry
    uint8_t *d1[2], *d2[2];
    size_t sp, ss, p1, s1, p2, s2; // pitch and image stride of the source, level +1, level +2
    int dw, dh;                    // size of level +1 (level +2 is dw >> 1 x dh >> 1)
    int n_first, rpt;              // rpt: level +1 rows per thread (even)
};

// level-1 bytes b0..b7 packed as v[i] = b[2i] | b[2i+1] << 16: g[k] (k = 0..3, two per word) = b[2k-1] + 2*b[2k] + b[2k+1]
__device__ __forceinline__ void pyr_hrow2(const uint32_t (&v)[4], uint32_t left, uint32_t (&g)[2])
{
    const uint32_t e0 = __byte_perm(v[0], v[1], 0x5410), o0 = __byte_perm(v[0], v[1], 0x7632);
    const uint32_t e1 = __byte_perm(v[2], v[3], 0x5410), o1 = __byte_perm(v[2], v[3], 0x7632);
    const uint32_t ol0 = __byte_perm(left, v[0], 0x7610), ol1 = __byte_perm(v[1], v[2], 0x7632);
    g[0] = ol0 + 2 * e0 + o0;
    g[1] = ol1 + 2 * e1 + o1;
}

#ifndef PYR_ROLL_MINB
#define PYR_ROLL_MINB 8 // 63 registers: 8 CTAs of 128 per SM (measured on B200, 512 pairs of 1080p, two steps: 440 us; 6 CTAs 449, 10 CTAs with spills 471, no bound 495)
#endif
#ifndef PYR_ROLL_UNROLL
#define PYR_ROLL_UNROLL 1
#endif
#define PYR_PRAGMA(x) _Pragma(#x)
#define PYR_UNROLL(n) PYR_PRAGMA(unroll n)
template <bool TWO>
__global__ void __launch_bounds__(128, PYR_ROLL_MINB) pyr_roll_kernel(const __grid_constant__ PyrRollParams P)
{
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int tx = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 31 + lane - 1;
    const int x0 = 8 * tx;
    const int b = (int)blockIdx.z >= P.n_first ? 1 : 0;
    const int zi = (int)blockIdx.z - (b ? P.n_first : 0);
    const bool col = tx >= 0 && x0 < P.dw;
    const bool owner = col && lane != 0;
    const uint8_t *s = P.src[b] + (size_t)zi * P.ss + 2 * x0;
    uint8_t *d1 = P.d1[b] + (size_t)zi * P.s1 + x0;
    uint8_t *d2 = TWO ? P.d2[b] + (size_t)zi * P.s2 + (x0 >> 1) : nullptr;
    const int dw2 = P.dw >> 1;
    const int y0 = blockIdx.y * P.rpt, y1 = min(y0 + P.rpt, P.dh);
    const bool st8 = owner && x0 + 7 < P.dw && ((reinterpret_cast<uintptr_t>(d1) | P.p1) & 7) == 0;
    const bool st4 = TWO && owner && (x0 >> 1) + 3 < dw2 && ((reinterpret_cast<uintptr_t>(d2) | P.p2) & 3) == 0;

    // source row sg (global), zero above the image and in dead lanes; rows below level +1's last row are never asked for
    auto load = [&](int sg) {
        uint4 q = make_uint4(0, 0, 0, 0);
        if (col && sg >= 0) q = __ldg(reinterpret_cast<const uint4 *>(s + (size_t)sg * P.sp));
        return q;
    };
    auto hrow = [&](const uint4 &q, uint32_t (&h)[4]) {
        const uint32_t left = __shfl_up_sync(0xffffffffu, q.w >> 24, 1);
        pyr_hrow(q, lane ? left : 0u, h);
    };
    auto vert = [&](const uint32_t (&a)[4], const uint32_t (&c)[4], const uint32_t (&e)[4], uint32_t (&v)[4]) {
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = ((a[i] + 2 * c[i] + e[i]) >> 4) & 0x00ff00ffu;
    };
    auto store1 = [&](int y, const uint32_t (&v)[4]) {
        const uint32_t lo = __byte_perm(v[0], v[1], 0x6420), hi = __byte_perm(v[2], v[3], 0x6420);
        uint8_t *d = d1 + (size_t)y * P.p1;
        if (st8) {
            *reinterpret_cast<uint2 *>(d) = make_uint2(lo, hi);
        } else if (owner) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (x0 + k < P.dw) d[k] = (uint8_t)(((k < 4 ? lo : hi) >> (8 * (k & 3))) & 255u);
        }
    };

    uint32_t hp[4], gp[2] = {0u, 0u};
    if (TWO) {
        // level +1 row y0 - 1 (not stored: the row group above owns it) starts the second level's vertical sums
        uint32_t ha[4], hb[4], v[4];
        hrow(load(2 * y0 - 3), ha);
        hrow(load(2 * y0 - 2), hb);
        hrow(load(2 * y0 - 1), hp);
        vert(ha, hb, hp, v);
        const uint32_t left = __shfl_up_sync(0xffffffffu, v[3] >> 16, 1);
        if (y0 > 0) pyr_hrow2(v, lane ? left : 0u, gp);
    } else {
        hrow(load(2 * y0 - 1), hp);
    }
    uint4 q[4];
#pragma unroll
    for (int r = 0; r < 4; r++) q[r] = (y0 + (r >> 1) < y1) ? load(2 * y0 + r) : make_uint4(0, 0, 0, 0);
PYR_UNROLL(PYR_ROLL_UNROLL)
    for (int y = y0; y < y1; y += 2) {
        uint4 qn[4];
#pragma unroll
        for (int r = 0; r < 4; r++) qn[r] = (y + 2 + (r >> 1) < y1) ? load(2 * (y + 2) + r) : make_uint4(0, 0, 0, 0);
        uint32_t h0[4], h1[4], h2[4], h3[4], va[4], vb[4];
        hrow(q[0], h0);
        hrow(q[1], h1);
        hrow(q[2], h2);
        hrow(q[3], h3);
        vert(hp, h0, h1, va);
        vert(h1, h2, h3, vb);
#pragma unroll
        for (int i = 0; i < 4; i++) hp[i] = h3[i];
        store1(y, va);
        if (y + 1 < y1) store1(y + 1, vb);
        if (TWO) {
            uint32_t ga[2], gb[2];
            const uint32_t la = __shfl_up_sync(0xffffffffu, va[3] >> 16, 1), lb = __shfl_up_sync(0xffffffffu, vb[3] >> 16, 1);
            pyr_hrow2(va, lane ? la : 0u, ga);
            pyr_hrow2(vb, lane ? lb : 0u, gb);
            if (y + 1 < P.dh) { // level +2 row y / 2 exists
                const uint32_t w0 = ((gp[0] + 2 * ga[0] + gb[0]) >> 4) & 0x00ff00ffu, w1 = ((gp[1] + 2 * ga[1] + gb[1]) >> 4) & 0x00ff00ffu;
                const uint32_t out = __byte_perm(w0, w1, 0x6420);
                uint8_t *d = d2 + (size_t)(y >> 1) * P.p2;
                if (st4) {
                    *reinterpret_cast<uint32_t *>(d) = out;
                } else if (owner) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if ((x0 >> 1) + k < dw2) d[k] = (uint8_t)((out >> (8 * k)) & 255u);
                }
            }
            gp[0] = gb[0];
            gp[1] = gb[1];
        }
#pragma unroll
        for (int r = 0; r < 4; r++) q[r] = qn[r];
    }
}

// Interleaved multi-channel u8 (the reference's 3-channel host layout), one output byte per thread.
// Only the host-pointer drop-in wrapper uses it; the channels are processed independently exactly
// like OptFlowGpu.cu:1224-1226.
__global__ void __launch_bounds__(256)
pyr_down_interleaved_kernel(const uint8_t *__restrict__ src, size_t src_pitch, int dw, int dh, int ch,
                            uint8_t *__restrict__ dst, size_t dst_pitch)
{
    const int xb = blockIdx.x * blockDim.x + threadIdx.x; // byte index within the output row
    const int y = blockIdx.y;
    if (xb >= dw * ch || y >= dh) return;
    const int x = xb / ch, c = xb - x * ch;
    int acc = 0;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int sy = 2 * y - 1 + r;
        if (sy < 0) continue;
        const uint8_t *row = src + (size_t)sy * src_pitch;
        const int a = (x > 0) ? (int)__ldg(row + (size_t)(2 * x - 1) * ch + c) : 0;
        const int b = (int)__ldg(row + (size_t)(2 * x) * ch + c);
        const int e = (int)__ldg(row + (size_t)(2 * x + 1) * ch + c);
        acc += ((r == 1) ? 2 : 1) * (a + 2 * b + e);
    }
    dst[(size_t)y * dst_pitch + xb] = (uint8_t)(acc >> 4);
}

int preload_pyramid()
{
    cudaFuncAttributes fa;
    OFB_CUDA_TRY(cudaFuncGetAttributes(&fa, pyr_down_planar_kernel));
    OFB_CUDA_TRY(cudaFuncGetAttributes(&fa, pyr_roll_kernel<false>));
    OFB_CUDA_TRY(cudaFuncGetAttributes(&fa, pyr_roll_kernel<true>));
    return OFB_OK;
}

int launch_pyr_down_strip(const uint8_t *src, size_t src_pitch, int sw, int src_rows, int src_y_off, uint8_t *dst,
                          size_t dst_pitch, int dst_y0, int dst_y1, cudaStream_t stream, unsigned long long *launches,
                          int n_images, size_t src_stride, size_t dst_stride)
{
    const int dw = sw >> 1, dh = dst_y1 - dst_y0;
    // rows 2*y-1 .. 2*y+1 of every destination row must be in the source strip (or above the image)
    const int first = 2 * dst_y0 - 1 < 0 ? 0 : 2 * dst_y0 - 1, last = 2 * (dst_y1 - 1) + 1;
    if (dw < 1 || dh < 1 || first < src_y_off || last >= src_y_off + src_rows) {
        set_error("pyr_down_strip: destination rows [%d,%d) need source rows [%d,%d], strip holds [%d,%d)", dst_y0, dst_y1,
                  first, last, src_y_off, src_y_off + src_rows);
        return OFB_ERR_INVALID;
    }
    if ((src_pitch & 7) || (dst_pitch & 3) || (reinterpret_cast<uintptr_t>(src) & 7) || (reinterpret_cast<uintptr_t>(dst) & 3)) {
        set_error("pyr_down_strip: planar images need 8-byte aligned source rows and 4-byte aligned destination rows");
        return OFB_ERR_INVALID;
    }
    if ((src_pitch & 15) || (reinterpret_cast<uintptr_t>(src) & 15)) {
        set_error("pyr_down_strip: source rows must be 16-byte aligned");
        return OFB_ERR_INVALID;
    }
    dim3 block(32, 8);
    if (n_images < 1 || (src_stride & 15) || (dst_stride & 3)) {
        set_error("pyr_down_strip: bad image count / strides");
        return OFB_ERR_INVALID;
    }
    dim3 grid((unsigned)((dw + 255) / 256), (unsigned)((dh + 8 * PYR_ROWS - 1) / (8 * PYR_ROWS)), (unsigned)n_images);
    OFB_CUDA_TRY(launch_pdl(pyr_down_planar_kernel, grid, block, 0, stream, src, src_pitch, src_stride, dw, dh, dst, dst_pitch, dst_stride,
                            src_y_off, dst_y0, src_rows, (const uint8_t *)nullptr, (uint8_t *)nullptr, n_images));
    if (launches) ++*launches;
    return OFB_OK;
}

// One (dst2 == nullptr) or two pyramid steps of a batch of whole planar frames in one launch; srcB / dst1B / dst2B: an
// optional second batch of n_images with the same geometry (the other frame of every pair).
int launch_pyr_roll(const uint8_t *src, size_t src_pitch, size_t src_stride, int sw, int sh, uint8_t *dst1, size_t pitch1,
                    size_t stride1, uint8_t *dst2, size_t pitch2, size_t stride2, int n_images, cudaStream_t stream,
                    unsigned long long *launches, const uint8_t *srcB, uint8_t *dst1B, uint8_t *dst2B, int sm_count)
{
    const int dw = sw >> 1, dh = sh >> 1;
    const int nz = srcB ? 2 * n_images : n_images;
    if (dw < 1 || dh < 1 || n_images < 1 || nz > 65535 || (dst2 && ((dw >> 1) < 1 || (dh >> 1) < 1))) {
        set_error("pyr_roll: bad geometry (src %dx%d, %d images)", sw, sh, n_images);
        return OFB_ERR_INVALID;
    }
    if ((src_pitch & 15) || (src_stride & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(srcB) & 15) ||
        (srcB && (!dst1B || (dst2 && !dst2B)))) {
        set_error("pyr_roll: planar source images need 16-byte aligned rows (pitch, base and image stride)");
        return OFB_ERR_INVALID;
    }
    PyrRollParams P;
    P.src[0] = src;
    P.src[1] = srcB;
    P.d1[0] = dst1;
    P.d1[1] = dst1B;
    P.d2[0] = dst2;
    P.d2[1] = dst2B;
    P.sp = src_pitch;
    P.ss = src_stride;
    P.p1 = pitch1;
    P.s1 = stride1;
    P.p2 = pitch2;
    P.s2 = stride2;
    P.dw = dw;
    P.dh = dh;
    P.n_first = n_images;
    // 31 lanes of a warp own 8 columns each; rows per thread: as many as still leave every SM a few dozen warps
    const int warps_x = ((dw + 7) / 8 + 30) / 31;
    const int wpb = warps_x < 4 ? warps_x : 4;
    const int gx = (warps_x + wpb - 1) / wpb;
    const long long want = (long long)(sm_count > 0 ? sm_count : 148) * 32;
    int rpt = 32;
    while (rpt > 2 && (long long)gx * wpb * ((dh + rpt - 1) / rpt) * nz < want) rpt >>= 1;
    P.rpt = rpt;
    dim3 grid((unsigned)gx, (unsigned)((dh + rpt - 1) / rpt), (unsigned)nz), block((unsigned)(32 * wpb));
    if (dst2) OFB_CUDA_TRY(launch_pdl(pyr_roll_kernel<true>, grid, block, 0, stream, P));
    else OFB_CUDA_TRY(launch_pdl(pyr_roll_kernel<false>, grid, block, 0, stream, P));
    if (launches) ++*launches;
    return OFB_OK;
}

int launch_pyr_down(const uint8_t *src, size_t src_pitch, size_t src_stride, int sw, int sh, uint8_t *dst,
                    size_t dst_pitch, size_t dst_stride, int n_images, int channels, cudaStream_t stream,
                    unsigned long long *launches, const uint8_t *src2, uint8_t *dst2)
{
    // src2 / dst2 (planar only): a second batch of n_images with the same geometry in the same launch
    const int dw = sw >> 1, dh = sh >> 1;
    if (src2 && (channels != 1 || !dst2 || (reinterpret_cast<uintptr_t>(src2) & 15))) {
        set_error("pyr_down: the second batch needs planar, 16-byte aligned images");
        return OFB_ERR_INVALID;
    }
    if (dw < 1 || dh < 1 || n_images < 1 || (channels != 1 && channels != 3)) {
        set_error("pyr_down: bad geometry (src %dx%d, %d images, %d channels)", sw, sh, n_images, channels);
        return OFB_ERR_INVALID;
    }
    if (channels == 1) {
        if ((src_pitch & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (src_stride & 15)) {
            set_error("pyr_down: planar source images need 16-byte aligned rows (pitch, base and image stride)");
            return OFB_ERR_INVALID;
        }
        const int nz = src2 ? 2 * n_images : n_images;
        if (nz > 65535) {
            set_error("pyr_down: at most 65535 images per launch");
            return OFB_ERR_INVALID;
        }
        static const bool old_kernel = [] {
            const char *e = getenv("OFB_PYR_OLD"); // developer switch for A/B measurements: the 8 x 2 outputs-per-thread kernel
            return e && e[0] == '1';
        }();
        if (!old_kernel)
            return launch_pyr_roll(src, src_pitch, src_stride, sw, sh, dst, dst_pitch, dst_stride, nullptr, 0, 0, n_images, stream,
                                   launches, src2, dst2, nullptr, 0);
        dim3 block(32, 8);
        dim3 grid((unsigned)((dw + 255) / 256), (unsigned)((dh + 8 * PYR_ROWS - 1) / (8 * PYR_ROWS)), (unsigned)nz);
        OFB_CUDA_TRY(launch_pdl(pyr_down_planar_kernel, grid, block, 0, stream, src, src_pitch, src_stride, dw, dh, dst, dst_pitch,
                                dst_stride, 0, 0, sh, src2, dst2, n_images));
        if (launches) ++*launches;
    } else {
        for (int i = 0; i < n_images; i++) {
            dim3 grid((unsigned)((dw * channels + 255) / 256), (unsigned)dh);
            pyr_down_interleaved_kernel<<<grid, 256, 0, stream>>>(src + (size_t)i * src_stride, src_pitch, dw, dh, channels,
                                                                 dst + (size_t)i * dst_stride, dst_pitch);
            OFB_CUDA_TRY(cudaGetLastError());
            if (launches) ++*launches;
        }
    }
    return OFB_OK;
}

} // namespace ofb

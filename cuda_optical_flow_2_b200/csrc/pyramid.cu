// pyramid.cu -- one Gaussian pyramid step: 3x3 binomial centred on (2x, 2y), decimate by 2.
// Replaces g_gauss_pyramid / gauss_pyramid_level (reference OptFlowGpu.cu:1198-1232, 1235-1259):
//   dst(x,y) = (unsigned char)( sum_{p,q} G[p][q] * src(2x-1+q, 2y-1+p) ),  G = [1 2 1;2 4 2;1 2 1]/16,
// taps left of / above the image skipped (no renormalisation).  With integer weights the fp32
// sum of the reference is exact, so (sum w_i p_i) >> 4 is byte-identical.
#include "ofb_common.cuh"

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

// pyramid.cu -- one Gaussian pyramid step: 3x3 binomial centred on (2x, 2y), decimate by 2.
// Replaces g_gauss_pyramid / gauss_pyramid_level (reference OptFlowGpu.cu:1198-1232, 1235-1259):
//   dst(x,y) = (unsigned char)( sum_{p,q} G[p][q] * src(2x-1+q, 2y-1+p) ),  G = [1 2 1;2 4 2;1 2 1]/16,
// taps left of / above the image skipped (no renormalisation).  With integer weights the fp32
// sum of the reference is exact, so (sum w_i p_i) >> 4 is byte-identical.
#include "ofb_common.cuh"

namespace ofb {

// Planar u8, four outputs per thread: 3 rows x (1 + 8) source bytes, one 32-bit store.
__global__ void __launch_bounds__(256)
pyr_down_planar_kernel(const uint8_t *__restrict__ src, size_t src_pitch, size_t src_stride, int dw, int dh,
                       uint8_t *__restrict__ dst, size_t dst_pitch, size_t dst_stride, int src_y_off, int dst_y0)
{
    // Row strips: dst local row y is global row dst_y0 + y; src local row 0 is global row src_y_off.
    const int x0 = 4 * (blockIdx.x * 32 + threadIdx.x);
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x0 >= dw || y >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * src_stride;
    uint8_t *d = dst + (size_t)blockIdx.z * dst_stride + (size_t)y * dst_pitch + x0;
    const bool wide = (size_t)(2 * x0 + 8) <= src_pitch; // the 8-byte load stays inside this row's pitch
    int acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int sg = 2 * (dst_y0 + y) - 1 + r; // global source row; above the image: skipped
        if (sg < 0) continue;
        const int sy = sg - src_y_off;
        const uint8_t *row = s + (size_t)sy * src_pitch + 2 * x0;
        int v[9];
        v[0] = (x0 > 0) ? (int)__ldg(row - 1) : 0;
        if (wide) {
            const uint2 q = __ldg(reinterpret_cast<const uint2 *>(row));
            v[1] = q.x & 255;
            v[2] = (q.x >> 8) & 255;
            v[3] = (q.x >> 16) & 255;
            v[4] = q.x >> 24;
            v[5] = q.y & 255;
            v[6] = (q.y >> 8) & 255;
            v[7] = (q.y >> 16) & 255;
            v[8] = q.y >> 24;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) v[1 + k] = ((size_t)(2 * x0 + k) < src_pitch) ? (int)__ldg(row + k) : 0;
        }
        const int wr = (r == 1) ? 2 : 1;
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] += wr * (v[2 * k] + 2 * v[2 * k + 1] + v[2 * k + 2]);
    }
    if (x0 + 3 < dw) {
        const uint32_t o = (uint32_t)(acc[0] >> 4) | ((uint32_t)(acc[1] >> 4) << 8) | ((uint32_t)(acc[2] >> 4) << 16) |
                           ((uint32_t)(acc[3] >> 4) << 24);
        *reinterpret_cast<uint32_t *>(d) = o;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (x0 + k < dw) d[k] = (uint8_t)(acc[k] >> 4);
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

int launch_pyr_down_strip(const uint8_t *src, size_t src_pitch, int sw, int src_rows, int src_y_off, uint8_t *dst,
                          size_t dst_pitch, int dst_y0, int dst_y1, cudaStream_t stream, unsigned long long *launches)
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
    dim3 block(32, 8);
    dim3 grid((unsigned)((dw + 127) / 128), (unsigned)((dh + 7) / 8), 1);
    pyr_down_planar_kernel<<<grid, block, 0, stream>>>(src, src_pitch, 0, dw, dh, dst, dst_pitch, 0, src_y_off, dst_y0);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

int launch_pyr_down(const uint8_t *src, size_t src_pitch, size_t src_stride, int sw, int sh, uint8_t *dst,
                    size_t dst_pitch, size_t dst_stride, int n_images, int channels, cudaStream_t stream,
                    unsigned long long *launches)
{
    const int dw = sw >> 1, dh = sh >> 1;
    if (dw < 1 || dh < 1 || n_images < 1 || (channels != 1 && channels != 3)) {
        set_error("pyr_down: bad geometry (src %dx%d, %d images, %d channels)", sw, sh, n_images, channels);
        return OFB_ERR_INVALID;
    }
    if (channels == 1) {
        if ((src_pitch & 7) || (dst_pitch & 3) || (reinterpret_cast<uintptr_t>(src) & 7) ||
            (reinterpret_cast<uintptr_t>(dst) & 3) || (src_stride & 7) || (dst_stride & 3)) {
            set_error("pyr_down: planar images need 8-byte aligned source rows and 4-byte aligned destination rows");
            return OFB_ERR_INVALID;
        }
        if (n_images > 65535) {
            set_error("pyr_down: at most 65535 images per launch");
            return OFB_ERR_INVALID;
        }
        dim3 block(32, 8);
        dim3 grid((unsigned)((dw + 127) / 128), (unsigned)((dh + 7) / 8), (unsigned)n_images);
        pyr_down_planar_kernel<<<grid, block, 0, stream>>>(src, src_pitch, src_stride, dw, dh, dst, dst_pitch, dst_stride, 0, 0);
        OFB_CUDA_TRY(cudaGetLastError());
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

// stages.cu -- the reference's stand-alone stage functions and layout helpers.
// These back the host-pointer drop-in wrappers (gpu::conv_3ch_1ch_tiled_uchar_float,
// gpu::srm_1ch_float, gpu::inverse_matrix_float); the hot path uses the fused kernel instead.
// Each keeps the reference's floating-point operation order, so results are bit-identical to the
// reference kernels for any float input, not only for integer-valued derivatives.
#include "ofb_common.cuh"

namespace ofb {

struct Mask25 {
    float m[25];
};

// g_conv_3ch_1ch_constant_uchar_float, OptFlowGpu.cu:1040-1090: correlation of channel 0 with an
// mw x mh mask, out-of-image taps and zero mask entries skipped, fp32 fma in row-major tap order.
// The mask travels as a kernel argument (the reference's shared __constant__ mask makes its API
// non re-entrant, OptFlowGpu.cu:190,1109).
__global__ void __launch_bounds__(256)
conv_c3_f32_kernel(const uint8_t *__restrict__ src, int w, int h, float *__restrict__ dst, Mask25 mk, int mw, int mh)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int hmw = mw >> 1, hmh = mh >> 1;
    float tmp = 0.0f;
    for (int i = 0; i < mh; i++) {
        const int ty = y - hmh + i;
        if (ty < 0 || ty >= h) continue;
        for (int j = 0; j < mw; j++) {
            const int tx = x - hmw + j;
            if (tx < 0 || tx >= w) continue;
            const float m = mk.m[i * mw + j];
            if (m == 0) continue;
            tmp = __fmaf_rn((float)__ldg(src + ((size_t)ty * w + tx) * 3), m, tmp);
        }
    }
    dst[(size_t)y * w + x] = tmp;
}

int launch_conv_c3_f32(const uint8_t *src_c3, int w, int h, float *dst, const float *mask_host, int mw, int mh,
                       cudaStream_t stream, unsigned long long *launches)
{
    if (w < 1 || h < 1 || mw < 1 || mh < 1 || mw * mh > 25) {
        set_error("conv: bad geometry (w %d h %d mask %dx%d, at most 25 taps)", w, h, mw, mh);
        return OFB_ERR_INVALID;
    }
    Mask25 mk;
    for (int i = 0; i < 25; i++) mk.m[i] = (i < mw * mh) ? mask_host[i] : 0.0f;
    dim3 block(32, 8), grid((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8));
    conv_c3_f32_kernel<<<grid, block, 0, stream>>>(src_c3, w, h, dst, mk, mw, mh);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// g_conv_3ch_1ch_constant, OptFlowGpu.cu:380-423 (what gpu::conv_3ch_1ch_tiled launches, :741-766): the u8 result
// variant used by the debug views.  An INT accumulator: `tmp += src * mask` per tap in row-major order is int ->
// float, one fused multiply-add (FFMA in the reference TU's SASS), truncation back to int, every tap; out-of-image
// taps and zero mask entries are skipped; the result is cast to unsigned char (wraps).
__device__ __forceinline__ int conv_u8_at(const uint8_t *__restrict__ src_c3, int w, int h, const Mask25 &mask, int mw, int mh,
                                          int x, int y)
{
    const int hmw = mw >> 1, hmh = mh >> 1;
    int tmp = 0;
    for (int i = 0; i < mh; i++) {
        const int ty = y - hmh + i;
        if (ty < 0 || ty >= h) continue;
        for (int j = 0; j < mw; j++) {
            const int tx = x - hmw + j;
            if (tx < 0 || tx >= w) continue;
            const float m = mask.m[i * mw + j];
            if (m == 0.0f) continue;
            tmp = __float2int_rz(__fmaf_rn((float)__ldg(src_c3 + ((size_t)ty * w + tx) * 3), m, (float)tmp));
        }
    }
    return tmp;
}

__global__ void __launch_bounds__(256)
conv_c3_u8_kernel(const uint8_t *__restrict__ src_c3, int w, int h, uint8_t *__restrict__ dst, const Mask25 mask, int mw, int mh)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    dst[(size_t)y * w + x] = (uint8_t)conv_u8_at(src_c3, w, h, mask, mw, mh, x, y);
}

int launch_conv_c3_u8(const uint8_t *src_c3, int w, int h, uint8_t *dst, const float *mask_host, int mw, int mh,
                      cudaStream_t stream, unsigned long long *launches)
{
    if (w < 1 || h < 1 || mw < 1 || mh < 1 || mw * mh > 25) {
        set_error("conv: bad geometry (w %d h %d mask %dx%d, at most 25 taps)", w, h, mw, mh);
        return OFB_ERR_INVALID;
    }
    Mask25 mk;
    for (int i = 0; i < 25; i++) mk.m[i] = (i < mw * mh) ? mask_host[i] : 0.0f;
    dim3 block(32, 8), grid((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8));
    conv_c3_u8_kernel<<<grid, block, 0, stream>>>(src_c3, w, h, dst, mk, mw, mh);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// One window of showTest (main.cu:19-92) for a pyramid level of size w x h, written at full resolution
// (w << k) x (h << k): the u8 convolution above with Dx_3x3 / Dy_3x3 on cur, or with Dt_3x3_n on cur minus on prev
// (cpu::sub_arr, unsigned char wrap-around), then utils::cleanup_outliers (>= 240 or < 20 -> 0, else 255) and
// utils::upscale_1ch (nearest).  One thread per OUTPUT pixel (coalesced stores; the 3x3 taps are cache hits).
__global__ void __launch_bounds__(256)
debug_view_kernel(const uint8_t *__restrict__ prev_c3, const uint8_t *__restrict__ cur_c3, int w, int h, int k, int which,
                  const Mask25 mask, uint8_t *__restrict__ out)
{
    const int ox = blockIdx.x * 32 + threadIdx.x, oy = blockIdx.y * 8 + threadIdx.y;
    if (ox >= (w << k) || oy >= (h << k)) return;
    const int x = ox >> k, y = oy >> k;
    uint8_t v = (uint8_t)conv_u8_at(cur_c3, w, h, mask, 3, 3, x, y);
    if (which == 2) v = (uint8_t)(v - (uint8_t)conv_u8_at(prev_c3, w, h, mask, 3, 3, x, y));
    out[(size_t)oy * ((size_t)w << k) + ox] = (v >= 240 || v < 20) ? 0 : 255;
}

int launch_debug_view(const uint8_t *prev_c3, const uint8_t *cur_c3, int w, int h, int k, int which, uint8_t *out,
                      cudaStream_t stream, unsigned long long *launches)
{
    static const float DX[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};                                                 // kernels.cpp:6-10
    static const float DY[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};                                                 // kernels.cpp:15-19
    static const float DTN[9] = {0.0666f, 0.1333f, 0.0666f, 0.1333f, 0.2f, 0.1333f, 0.0666f, 0.1333f, 0.0666f}; // kernels.cpp:25-28
    if (w < 1 || h < 1 || k < 0 || k > 12 || which < 0 || which > 2 || ((size_t)h << k) > 0x7fffffffull / ((size_t)w << k)) {
        set_error("debug_view: bad arguments (w %d h %d level %d view %d)", w, h, k, which);
        return OFB_ERR_INVALID;
    }
    const float *m = which == 0 ? DX : which == 1 ? DY : DTN;
    Mask25 mk;
    for (int i = 0; i < 25; i++) mk.m[i] = i < 9 ? m[i] : 0.0f;
    dim3 block(32, 8), grid((unsigned)(((w << k) + 31) / 32), (unsigned)(((h << k) + 7) / 8));
    debug_view_kernel<<<grid, block, 0, stream>>>(prev_c3, cur_c3, w, h, k, which, mk, out);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// g_srm_1ch_float, OptFlowGpu.cu:1549-1588: windowed sum of a*b, out-of-image taps skipped,
// `tmp += a*b` (one fma per tap) in row-major tap order.  A 32x8 output tile stages its
// (32+ww-1) x (8+wh-1) footprint of both inputs in shared memory; out-of-image positions are
// flagged so that the tap order and skipping match the reference exactly.
constexpr int SRM_BX = 32, SRM_BY = 8, SRM_MAXW = 32; // window up to 32x32
__global__ void __launch_bounds__(SRM_BX *SRM_BY)
srm_f32_kernel(const float *__restrict__ a, const float *__restrict__ b, int w, int h, int ww, int wh,
               float *__restrict__ dst)
{
    extern __shared__ float sm[];
    const int tw = SRM_BX + ww - 1, th = SRM_BY + wh - 1;
    float *sa = sm, *sb = sm + tw * th;
    const int hww = ww >> 1, hwh = wh >> 1;
    const int bx0 = blockIdx.x * SRM_BX - hww, by0 = blockIdx.y * SRM_BY - hwh;
    const int tid = threadIdx.y * SRM_BX + threadIdx.x;
    for (int t = tid; t < tw * th; t += SRM_BX * SRM_BY) {
        const int ty = t / tw, tx = t - ty * tw;
        const int gx = bx0 + tx, gy = by0 + ty;
        const bool in = gx >= 0 && gx < w && gy >= 0 && gy < h;
        sa[t] = in ? __ldg(a + (size_t)gy * w + gx) : 0.0f;
        sb[t] = in ? __ldg(b + (size_t)gy * w + gx) : 0.0f;
    }
    __syncthreads();
    const int x = blockIdx.x * SRM_BX + threadIdx.x, y = blockIdx.y * SRM_BY + threadIdx.y;
    if (x >= w || y >= h) return;
    float tmp = 0.0f;
    for (int p = 0; p < wh; p++) {
        const int gy = y - hwh + p;
        if (gy < 0 || gy >= h) continue;
        const float *ra = sa + (threadIdx.y + p) * tw + threadIdx.x, *rb = sb + (threadIdx.y + p) * tw + threadIdx.x;
        for (int q = 0; q < ww; q++) {
            const int gx = x - hww + q;
            if (gx < 0 || gx >= w) continue;
            tmp = __fmaf_rn(ra[q], rb[q], tmp);
        }
    }
    dst[(size_t)y * w + x] = tmp;
}

int launch_srm_f32(const float *a, const float *b, int w, int h, int ww, int wh, float *dst, cudaStream_t stream,
                   unsigned long long *launches)
{
    if (w < 1 || h < 1 || ww < 1 || wh < 1 || ww > SRM_MAXW || wh > SRM_MAXW) {
        set_error("srm: bad geometry (w %d h %d window %dx%d, at most %dx%d)", w, h, ww, wh, SRM_MAXW, SRM_MAXW);
        return OFB_ERR_INVALID;
    }
    const size_t smem = (size_t)(SRM_BX + ww - 1) * (SRM_BY + wh - 1) * 2 * sizeof(float);
    dim3 block(SRM_BX, SRM_BY), grid((unsigned)((w + SRM_BX - 1) / SRM_BX), (unsigned)((h + SRM_BY - 1) / SRM_BY));
    srm_f32_kernel<<<grid, block, smem, stream>>>(a, b, w, h, ww, wh, dst);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// g_inv_matrix_float, OptFlowGpu.cu:1819-1846, same double-precision operation order (lk_solve
// takes exact integers; this variant takes the reference's float sums).
__global__ void __launch_bounds__(256)
inverse_f32_kernel(const float *__restrict__ sxx, const float *__restrict__ syy, const float *__restrict__ sxy,
                   const float *__restrict__ sxt, const float *__restrict__ syt, float2 *__restrict__ flow, int n)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double a = (double)sxx[i], b = (double)sxy[i], d = (double)syy[i], tx = (double)sxt[i], ty = (double)syt[i];
    const double det = __fma_rn(a, d, -__dmul_rn(b, b));
    const double prefix = 1.0 / det;
    const double ap = __dmul_rn(a, prefix), bp = __dmul_rn(b, prefix), dp = __dmul_rn(d, prefix);
    flow[i] = make_float2((float)__fma_rn(bp, ty, -__dmul_rn(dp, tx)), (float)__fma_rn(bp, tx, -__dmul_rn(ap, ty)));
}

int launch_inverse_f32(const float *sxx, const float *syy, const float *sxy, const float *sxt, const float *syt,
                       float *flow, int n, cudaStream_t stream, unsigned long long *launches)
{
    if (n < 1) {
        set_error("inverse: empty input");
        return OFB_ERR_INVALID;
    }
    inverse_f32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(sxx, syy, sxy, sxt, syt, reinterpret_cast<float2 *>(flow), n);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// Channel 0 of 3-channel interleaved u8 -> planar pitched u8.  VEC: four pixels per thread -- three aligned 32-bit loads
// (12 source bytes), two byte permutes, one 32-bit store -- for widths that are multiples of 4 on 4-byte aligned
// buffers; otherwise one pixel per thread.
template <bool VEC>
__global__ void __launch_bounds__(256)
c3_to_planar_kernel(const uint8_t *__restrict__ src, int w, int h, uint8_t *__restrict__ dst, size_t dst_pitch,
                    size_t dst_stride)
{
    const int x = (blockIdx.x * 256 + threadIdx.x) * (VEC ? 4 : 1), y = blockIdx.y;
    if (x >= w) return;
    const size_t img = blockIdx.z;
    const uint8_t *s = src + (img * (size_t)w * h + (size_t)y * w + x) * 3;
    uint8_t *d = dst + img * dst_stride + (size_t)y * dst_pitch + x;
    if (VEC) {
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(s);
        const uint32_t a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2); // bytes 0..11: channel 0 at 0, 3, 6, 9
        const uint32_t lo = __byte_perm(a, b, 0x0630);                      // [a.b0, a.b3, b.b2, -]
        *reinterpret_cast<uint32_t *>(d) = __byte_perm(lo, c, 0x5210);      // [.., .., .., c.b1]
    } else {
        *d = __ldg(s);
    }
}

int launch_c3_to_planar(const uint8_t *src_c3, int w, int h, int n_images, uint8_t *dst, size_t dst_pitch,
                        size_t dst_stride, cudaStream_t stream, unsigned long long *launches)
{
    if (w < 1 || h < 1 || h > 65535 || n_images < 1 || n_images > 65535) {
        set_error("c3_to_planar: bad geometry (w %d h %d images %d)", w, h, n_images);
        return OFB_ERR_INVALID;
    }
    const bool vec = (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(src_c3) | reinterpret_cast<uintptr_t>(dst) | dst_pitch | dst_stride) & 3) == 0;
    if (vec) {
        dim3 grid((unsigned)((w / 4 + 255) / 256), (unsigned)h, (unsigned)n_images);
        c3_to_planar_kernel<true><<<grid, 256, 0, stream>>>(src_c3, w, h, dst, dst_pitch, dst_stride);
    } else {
        dim3 grid((unsigned)((w + 255) / 256), (unsigned)h, (unsigned)n_images);
        c3_to_planar_kernel<false><<<grid, 256, 0, stream>>>(src_c3, w, h, dst, dst_pitch, dst_stride);
    }
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// cum_k = 2*cum_{k+1}[i>>1, j>>1] + flow_k : the composition rule of main.cu:136-147 applied
// coarse to fine (power-of-two scaling is exact, so this equals the reference's running sum).
__global__ void __launch_bounds__(256)
compose_cum_kernel(const float2 *__restrict__ flow, const float2 *__restrict__ coarser, int w, int h, int cw, int ch,
                   float2 *__restrict__ out)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t pair = blockIdx.z;
    const float2 f = flow[pair * (size_t)w * h + (size_t)y * w + x];
    float2 c = make_float2(0.0f, 0.0f);
    if (coarser) c = __ldg(coarser + pair * (size_t)cw * ch + (size_t)min(y >> 1, ch - 1) * cw + min(x >> 1, cw - 1));
    out[pair * (size_t)w * h + (size_t)y * w + x] = make_float2(2.0f * c.x + f.x, 2.0f * c.y + f.y);
}

int launch_compose_cum(const float *flow_k, const float *cum_coarser, int w, int h, int n_pairs, float *cum_out,
                       cudaStream_t stream, unsigned long long *launches)
{
    if (w < 1 || h < 1 || h > 65535 || n_pairs < 1 || n_pairs > 65535) {
        set_error("compose_cum: bad geometry (w %d h %d pairs %d)", w, h, n_pairs);
        return OFB_ERR_INVALID;
    }
    dim3 grid((unsigned)((w + 255) / 256), (unsigned)h, (unsigned)n_pairs);
    compose_cum_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float2 *>(flow_k),
                                                 reinterpret_cast<const float2 *>(cum_coarser), w, h, w >> 1, h >> 1,
                                                 reinterpret_cast<float2 *>(cum_out));
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

} // namespace ofb

// =============================================================================================
// Pre-processing of the reference's frame loop: grayscale and the bilateral
// pre-filter that main.cu:232-240 applies to every new frame before the pyramid.
// =============================================================================================
namespace ofb {

// g_grayscale_avg_2d, OptFlowGpu.cu:47-60: avg = (c0 + c1 + c2) / 3 (integer division), written to
// all three channels (c3 -> c3, for the drop-in wrapper) or to a planar image (hot path).
__global__ void __launch_bounds__(256)
grayscale_kernel(const uint8_t *__restrict__ src, int w, int h, uint8_t *__restrict__ dst_c3, uint8_t *__restrict__ dst_planar,
                 size_t dst_pitch)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const size_t pos = ((size_t)y * w + x) * 3;
    const int avg = ((int)__ldg(src + pos) + (int)__ldg(src + pos + 1) + (int)__ldg(src + pos + 2)) / 3;
    if (dst_c3) dst_c3[pos] = dst_c3[pos + 1] = dst_c3[pos + 2] = (uint8_t)avg;
    if (dst_planar) dst_planar[(size_t)y * dst_pitch + x] = (uint8_t)avg;
}

int launch_grayscale(const uint8_t *src_c3, int w, int h, uint8_t *dst_c3, uint8_t *dst_planar, size_t dst_pitch,
                     cudaStream_t stream, unsigned long long *launches)
{
    if (w < 1 || h < 1 || h > 65535) {
        set_error("grayscale: bad geometry (w %d h %d)", w, h);
        return OFB_ERR_INVALID;
    }
    dim3 grid((unsigned)((w + 255) / 256), (unsigned)h);
    grayscale_kernel<<<grid, 256, 0, stream>>>(src_c3, w, h, dst_c3, dst_planar, dst_pitch);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

// Range weights of g_bilinear_filter (OptFlowGpu.cu:1984-2048; a bilateral filter despite its name):
//   n_b(k) = 1/(2*pi*sigmaB^2) * pow(e, -0.5*k^2/sigmaB^2),  k = gray(tap) - gray(centre) in -255..255.
// The reference evaluates this per tap; it depends on |k| only, so a 256-entry table built with the
// same double-precision operation order (read from the reference TU's SASS: (k*k*-0.5)/sigmaB2, pow,
// times the hoisted 1/(2*pi*sigmaB2)) gives bit-identical weights at 1/81 of the pow calls.
__global__ void bilateral_lut_kernel(double sigmaB, double *__restrict__ lut)
{
    const int d = threadIdx.x; // 0..255
    const double sigmaB2 = sigmaB * sigmaB;
    const double pref = 1.0 / (2.0 * M_PI * sigmaB2);
    const double k = (double)d;
    const double k2 = k * k;
    lut[d] = pref * pow(M_E, -0.5 * (k2) / sigmaB2);
}

struct SpatialMask {
    double m[100]; // up to 10x10, like the reference's gaus_kernel_10x10_gpu (OptFlowGpu.cu:1982)
};

// One output per thread; gray tile with halo and the range table in shared memory.
//   wsb = fma(n_b, n_s, wsb);  tmp_c = fma(n_s, n_b*src_c, tmp_c)   in row-major tap order,
//   out-of-image taps skipped; dest_c = (unsigned char)(tmp_c / wsb)  -- the reference's order.
// CH = 1: planar, src == gray (the main.cu use).  CH = 3: interleaved src and gray (drop-in wrapper).
constexpr int BIL_BX = 32, BIL_BY = 8;
template <int CH>
__global__ void __launch_bounds__(BIL_BX *BIL_BY)
bilateral_kernel(const uint8_t *__restrict__ src, const uint8_t *__restrict__ gray, size_t pitch, int w, int h, int ww,
                 int wh, const double *__restrict__ lut_g, SpatialMask sm, uint8_t *__restrict__ dst, size_t dst_pitch)
{
    extern __shared__ __align__(16) uint8_t smem_b[];
    double *lut = reinterpret_cast<double *>(smem_b);
    uint8_t *tile = smem_b + 256 * sizeof(double);
    const int tw = BIL_BX + ww - 1, th = BIL_BY + wh - 1;
    const int hww = ww >> 1, hwh = wh >> 1;
    const int bx0 = blockIdx.x * BIL_BX - hww, by0 = blockIdx.y * BIL_BY - hwh;
    const int tid = threadIdx.y * BIL_BX + threadIdx.x;
    for (int t = tid; t < 256; t += BIL_BX * BIL_BY) lut[t] = lut_g[t];
    for (int t = tid; t < tw * th; t += BIL_BX * BIL_BY) {
        const int ty = t / tw, tx = t - ty * tw;
        const int gx = bx0 + tx, gy = by0 + ty;
        tile[t] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? __ldg(gray + (size_t)gy * pitch + (size_t)gx * CH) : 0;
    }
    __syncthreads();
    const int x = blockIdx.x * BIL_BX + threadIdx.x, y = blockIdx.y * BIL_BY + threadIdx.y;
    if (x >= w || y >= h) return;
    const int f_ij = tile[(threadIdx.y + hwh) * tw + threadIdx.x + hww];
    double wsb = 0.0, tmp[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) tmp[c] = 0.0;
    for (int m = 0; m < wh; m++) {
        const int cy = y - hwh + m;
        if (cy < 0 || cy >= h) continue;
        for (int n = 0; n < ww; n++) {
            const int cx = x - hww + n;
            if (cx < 0 || cx >= w) continue;
            const int f_mn = tile[(threadIdx.y + m) * tw + threadIdx.x + n];
            const int k = f_mn - f_ij;
            const double n_b = lut[k < 0 ? -k : k];
            const double n_s = sm.m[m * ww + n];
            wsb = __fma_rn(n_b, n_s, wsb);
            if (CH == 1) {
                tmp[0] = __fma_rn(n_s, __dmul_rn(n_b, (double)f_mn), tmp[0]);
            } else {
                const uint8_t *sp = src + (size_t)cy * pitch + (size_t)cx * CH;
#pragma unroll
                for (int c = 0; c < CH; c++) tmp[c] = __fma_rn(n_s, __dmul_rn(n_b, (double)__ldg(sp + c)), tmp[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CH; c++) dst[(size_t)y * dst_pitch + (size_t)x * CH + c] = (uint8_t)(unsigned int)(tmp[c] / wsb);
}

// utils::generate_gaussian_kernel, OptFlowUtils.cpp:68-114 (host, double, libm pow): the normalised
// spatial Gaussian the reference uploads into its __constant__ table.  Same arithmetic, same order.
static int make_spatial_mask(double sigmaS, int ksize, SpatialMask *out)
{
    if (ksize < 1 || ksize > 10 || (ksize & 1) == 0) {
        set_error("bilateral: spatial window %d not supported (odd, at most 9; the reference's table holds 10x10)", ksize);
        return OFB_ERR_UNSUPPORTED;
    }
    const int hk = ksize >> 1;
    double *g = out->m;
    for (int i = 0; i < hk + 1; i++)
        for (int j = 0; j < hk + 1; j++) {
            const double sigmaS2 = sigmaS * sigmaS;
            const double m = i, n = j;
            const double n2 = n * n, m2 = m * m;
            const double value = 1.0 / (2.0 * M_PI * sigmaS2) * pow(M_E, -0.5 * (n2 + m2) / sigmaS2);
            g[(hk + i) * ksize + hk + j] = value;
            g[(hk - i) * ksize + hk - j] = value;
            g[(hk + i) * ksize + hk - j] = value;
            g[(hk - i) * ksize + hk + j] = value;
        }
    double sum = 0;
    for (int i = 0; i < ksize * ksize; i++) sum += g[i];
    for (int i = 0; i < ksize * ksize; i++) g[i] /= sum;
    return OFB_OK;
}

int launch_bilateral(const uint8_t *src, const uint8_t *gray, size_t pitch, int channels, int w, int h, int ww, int wh,
                     double sigmaS, double sigmaB, double *lut_dev, uint8_t *dst, size_t dst_pitch, cudaStream_t stream,
                     unsigned long long *launches)
{
    if (w < 1 || h < 1 || ww != wh || (channels != 1 && channels != 3)) {
        set_error("bilateral: bad arguments (w %d h %d window %dx%d channels %d; the window must be square)", w, h, ww, wh,
                  channels);
        return OFB_ERR_INVALID;
    }
    SpatialMask sm;
    int rc = make_spatial_mask(sigmaS, ww, &sm);
    if (rc) return rc;
    bilateral_lut_kernel<<<1, 256, 0, stream>>>(sigmaB, lut_dev);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    const size_t smem = 256 * sizeof(double) + (size_t)(BIL_BX + ww - 1) * (BIL_BY + wh - 1);
    dim3 block(BIL_BX, BIL_BY), grid((unsigned)((w + BIL_BX - 1) / BIL_BX), (unsigned)((h + BIL_BY - 1) / BIL_BY));
    if (channels == 1)
        bilateral_kernel<1><<<grid, block, smem, stream>>>(src, gray, pitch, w, h, ww, wh, lut_dev, sm, dst, dst_pitch);
    else
        bilateral_kernel<3><<<grid, block, smem, stream>>>(src, gray, pitch, w, h, ww, wh, lut_dev, sm, dst, dst_pitch);
    OFB_CUDA_TRY(cudaGetLastError());
    if (launches) ++*launches;
    return OFB_OK;
}

} // namespace ofb

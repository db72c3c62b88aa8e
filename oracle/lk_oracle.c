/*
 * lk_oracle.c -- CPU restatement of the reference's dense pyramidal Lucas-Kanade GPU path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (cuda_optical_flow_2_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks every function here against the
 * reference's own code compiled from /root/reference into oracle/_ref/libofref.so (CPU-side
 * functions run in the build container; the reference's GPU stage functions were run on a B200
 * at launch-valid sizes and their outputs are committed under tests/golden/, see
 * oracle/make_golden.py).  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4).
 *
 * All file:line citations are into the reference tree (Kr-Stam/CUDA_Optical_Flow_2).
 * Images here are PLANAR u8 (one channel); the reference stores 3 equal interleaved channels
 * and its LK stages read channel 0 only (OptFlowGpu.cu:1081), its pyramid treats the channels
 * independently (OptFlowGpu.cu:1224-1226), so planar is the same arithmetic on 1/3 of the bytes.
 */
#define _USE_MATH_DEFINES
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_WARP_AS_WRITTEN 0
#define ORC_WARP_NEAREST 1
#define ORC_WARP_BILINEAR 2

#define ORC_SUMS_F32_SEQUENTIAL 0 /* the reference's fp32 running sum, tap order, fma-contracted */
#define ORC_SUMS_EXACT 1          /* exact integer window sums (what the B200 kernel computes) */

/* masks: kernels.cpp:6-10 (Dx_3x3), :15-19 (Dy_3x3), :20-24 (Dt_3x3) */
static const float ORC_DX[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};
static const float ORC_DY[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};
static const float ORC_DT[9] = {1, 2, 1, 2, 3, 2, 1, 2, 1};

/* ---------------------------------------------------------------------------------------------
 * a1. One pyramid level.  OptFlowGpu.cu:1198-1232 (g_gauss_pyramid), mask :1193-1196,
 * identical arithmetic in OptFlowCPU.cpp:112-148 (cpu::downscale_gaussian).
 *   dst(x,y) = (unsigned char)( sum_{p,q in 0..2} G[p][q] * src(2x-1+q, 2y-1+p) ),
 *   G = [1 2 1; 2 4 2; 1 2 1]/16, taps with cx<0, cy<0, cx>=2*dw or cy>=2*dh skipped, no
 *   renormalisation.  fp32 accumulation of k/16 multiples below 256 is exact, so the float
 *   running sum equals the integer (sum w_i p_i)/16 and the cast truncates.
 * dw = sw>>1, dh = sh>>1 (main.cu:101-102 allocates levels that way).  The reference indexes
 * the source with stride 2*dw (:1221), which equals sw only for even sw; this restatement uses
 * the true stride sw, i.e. the intended image, for odd sw (no BASELINE config has one).
 * ------------------------------------------------------------------------------------------- */
void orc_pyr_down_u8(const uint8_t *src, int sw, int sh, uint8_t *dst)
{
    const int dw = sw >> 1, dh = sh >> 1;
    static const float G[9] = {0.0625f, 0.125f, 0.0625f, 0.125f, 0.25f, 0.125f, 0.0625f, 0.125f, 0.0625f};
    for (int y = 0; y < dh; y++) {
        for (int x = 0; x < dw; x++) {
            float tmp = 0.0f;
            const int sy0 = (y << 1) - 1, sx0 = (x << 1) - 1;
            for (int p = 0; p < 3; p++) {
                for (int q = 0; q < 3; q++) {
                    const int cx = sx0 + q, cy = sy0 + p;
                    if (cx >= 0 && cx < dw * 2 && cy >= 0 && cy < dh * 2)
                        tmp += G[p * 3 + q] * (float)src[(size_t)cy * sw + cx];
                }
            }
            dst[(size_t)y * dw + x] = (unsigned char)tmp;
        }
    }
}

/* gpu::gauss_pyramid, OptFlowGpu.cu:1262-1271: level k from level k-1, k = 1..levels-1;
 * pyramid[k] is (w>>k) x (h>>k), caller-allocated (main.cu:95-104). */
void orc_gauss_pyramid_u8(uint8_t **pyramid, int w, int h, int levels)
{
    for (int k = 1; k < levels; k++)
        orc_pyr_down_u8(pyramid[k - 1], w >> (k - 1), h >> (k - 1), pyramid[k]);
}

/* ---------------------------------------------------------------------------------------------
 * a3. Derivative "convolution" (a correlation; the mask is not flipped).
 * OptFlowGpu.cu:1040-1090 (g_conv_3ch_1ch_constant_uchar_float): zero padding by skipping
 * out-of-image taps (:1064-1073), zero mask entries skipped (:1077), fp32 running sum in
 * row-major tap order, nvcc contracts `tmp += (float)src * mask` into one fma.
 * ------------------------------------------------------------------------------------------- */
void orc_conv_u8_f32(const uint8_t *src, int w, int h, const float *mask, int mw, int mh, float *dst)
{
    const int hmw = mw >> 1, hmh = mh >> 1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float tmp = 0.0f;
            for (int i = 0; i < mh; i++) {
                const int ty = y - hmh + i;
                if (ty < 0 || ty >= h) continue;
                for (int j = 0; j < mw; j++) {
                    const int tx = x - hmw + j;
                    if (tx < 0 || tx >= w) continue;
                    const float m = mask[i * mw + j];
                    if (m == 0) continue;
                    tmp = fmaf((float)src[(size_t)ty * w + tx], m, tmp);
                }
            }
            dst[(size_t)y * w + x] = tmp;
        }
    }
}

/* a4. utils::arr_sub_float, OptFlowUtils.hpp:21-31: dest[i] = arr1[i] - arr2[i]. */
void orc_sub_f32(const float *a, const float *b, size_t n, float *dst)
{
    for (size_t i = 0; i < n; i++) dst[i] = a[i] - b[i];
}

/* ---------------------------------------------------------------------------------------------
 * a5. Window product-sum.  OptFlowGpu.cu:1549-1588 (g_srm_1ch_float):
 *   S(x,y) = sum_{p<wh} sum_{q<ww} a(x-ww/2+q, y-wh/2+p) * b(same); out-of-image taps skipped;
 *   fp32 running sum `tmp += a*b` in row-major tap order (one fma per tap after contraction).
 * ------------------------------------------------------------------------------------------- */
void orc_srm_f32(const float *a, const float *b, int w, int h, int ww, int wh, float *dst)
{
    const int hww = ww >> 1, hwh = wh >> 1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            float tmp = 0.0f;
            for (int p = 0; p < wh; p++) {
                const int ty = y - hwh + p;
                if (ty < 0 || ty >= h) continue;
                for (int q = 0; q < ww; q++) {
                    const int tx = x - hww + q;
                    if (tx < 0 || tx >= w) continue;
                    const size_t k = (size_t)ty * w + tx;
                    tmp = fmaf(a[k], b[k], tmp);
                }
            }
            dst[(size_t)y * w + x] = tmp;
        }
    }
}

/* Same sums with exact integer arithmetic (inputs are exact small integers, a3).  Uses a
 * summed-area table in int64 so that big frames stay cheap; the result is the mathematically
 * exact sum, which is what order-independent int32 accumulation on the GPU produces. */
void orc_srm_exact_i64(const float *a, const float *b, int w, int h, int ww, int wh, int64_t *dst)
{
    const int hww = ww >> 1, hwh = wh >> 1;
    const size_t W1 = (size_t)w + 1;
    int64_t *sat = (int64_t *)calloc(W1 * ((size_t)h + 1), sizeof(int64_t));
    for (int y = 0; y < h; y++) {
        int64_t row = 0;
        for (int x = 0; x < w; x++) {
            const size_t k = (size_t)y * w + x;
            row += (int64_t)a[k] * (int64_t)b[k];
            sat[(size_t)(y + 1) * W1 + (x + 1)] = sat[(size_t)y * W1 + (x + 1)] + row;
        }
    }
    for (int y = 0; y < h; y++) {
        int y0 = y - hwh, y1 = y - hwh + wh; /* [y0, y1) */
        if (y0 < 0) y0 = 0;
        if (y1 > h) y1 = h;
        for (int x = 0; x < w; x++) {
            int x0 = x - hww, x1 = x - hww + ww;
            if (x0 < 0) x0 = 0;
            if (x1 > w) x1 = w;
            dst[(size_t)y * w + x] = sat[(size_t)y1 * W1 + x1] - sat[(size_t)y0 * W1 + x1] -
                                     sat[(size_t)y1 * W1 + x0] + sat[(size_t)y0 * W1 + x0];
        }
    }
    free(sat);
}

/* ---------------------------------------------------------------------------------------------
 * a6. Per-pixel 2x2 solve.  OptFlowGpu.cu:1819-1846 (g_inv_matrix_float), all in double:
 *   a = SIx2, b = c = SIxIy, d = SIy2; prefix = 1/(a*d - b*c); a,b,c,d *= prefix;
 *   u = (float)(-d*SIxIt + b*SIyIt);  v = (float)(c*SIxIt - a*SIyIt);   flow = (u, v) interleaved.
 * No determinant threshold: det == 0 yields inf/NaN (Q3).  The fma contraction below is the
 * one nvcc 12.9 emits for that source on sm_100a (cuobjdump -sass of the reference TU):
 *   det = fma(a, d, -(b*b)); u = fma(b', SIyIt, -(d'*SIxIt)); v = fma(b', SIxIt, -(a'*SIyIt)).
 * ------------------------------------------------------------------------------------------- */
static inline void orc_solve_one(double a, double b, double d, double sxt, double syt, float *u, float *v)
{
    const double bb = b * b;
    const double det = fma(a, d, -bb);
    const double prefix = 1.0 / det;
    const double ap = a * prefix, bp = b * prefix, dp = d * prefix;
    *u = (float)fma(bp, syt, -(dp * sxt));
    *v = (float)fma(bp, sxt, -(ap * syt));
}

void orc_solve_f32(const float *sxx, const float *syy, const float *sxy, const float *sxt, const float *syt,
                   size_t n, float *flow)
{
    for (size_t i = 0; i < n; i++)
        orc_solve_one((double)sxx[i], (double)sxy[i], (double)syy[i], (double)sxt[i], (double)syt[i],
                      &flow[2 * i], &flow[2 * i + 1]);
}

void orc_solve_i64(const int64_t *sxx, const int64_t *syy, const int64_t *sxy, const int64_t *sxt,
                   const int64_t *syt, size_t n, float *flow)
{
    for (size_t i = 0; i < n; i++)
        orc_solve_one((double)sxx[i], (double)sxy[i], (double)syy[i], (double)sxt[i], (double)syt[i],
                      &flow[2 * i], &flow[2 * i + 1]);
}

/* ---------------------------------------------------------------------------------------------
 * a2 / a10. Cumulative coarser flow and the warp ("shift next back").
 * cpu::shift_back_pyramid, OptFlowCPU.cpp:241-282, called from OptFlowGpu.cu:1918-1922.
 *   for k = maxLevel-1 .. level+1:  off = k-level;  (u,v) += (float)(1<<off) * flow_k[pos_k]
 *   new_x = (int)(j + u); new_y = (int)(i + v); out-of-image targets are skipped (:270-273).
 * Modes:
 *   AS_WRITTEN: pos_k = 0 for every pixel, because `i * (1 >> offset)` is always 0 (:260-261).
 *   NEAREST:    pos_k = (i>>off, j>>off), the indexing main.cu:141-143 uses (the intended one);
 *               indices are clamped to the level (only matters for odd sizes).
 *   BILINEAR:   same (u,v) as NEAREST, sampled bilinearly in 8.8 fixed point (new work asked for
 *               by BASELINE.json north_star; no reference counterpart).  The flow is rounded to
 *               1/256 px once, so that every pixel of a 2x2 block (which shares one coarser flow
 *               vector) has the same integer offset and the same two weights:
 *                 skipped unless |u| < 32768 and |v| < 32768 (also rejects NaN)
 *                 U = rint(u*256), V = rint(v*256)  (round half to even); X = j*256 + U, Y = i*256 + V
 *                 skipped unless 0 <= X <= (w-1)*256 and 0 <= Y <= (h-1)*256
 *                 x0 = X>>8, wx = X&255, x1 = min(x0+1, w-1), same for y
 *                 val = ((256-wy)*((256-wx)*p00 + wx*p01) + wy*((256-wx)*p10 + wx*p11) + 32768) >> 16
 * Skipped pixels keep the unwarped next pixel.  (The reference pre-fills with a memcpy of only
 * w*h of the 3*w*h bytes (:247), so there skipped pixels beyond the first third are
 * uninitialised heap; the pin test masks those.)  flow_scale multiplies (u,v) before the add
 * (1.0 = the reference, Q5).
 * ------------------------------------------------------------------------------------------- */
static void orc_cum_flow(int i, int j, int W0, int H0, int level, int maxLevel, float *const *flowPyr, int mode,
                         float *pu, float *pv)
{
    float u = 0.0f, v = 0.0f;
    for (int k = maxLevel - 1; k > level; k--) {
        const int off = k - level;
        const int wk = W0 >> k, hk = H0 >> k;
        int ti = 0, tj = 0;
        if (mode != ORC_WARP_AS_WRITTEN) {
            ti = i >> off;
            tj = j >> off;
            if (ti > hk - 1) ti = hk - 1;
            if (tj > wk - 1) tj = wk - 1;
        }
        const size_t pos = (size_t)ti * wk + tj;
        const float mult = (float)(1 << off);
        u += mult * flowPyr[k][pos * 2];
        v += mult * flowPyr[k][pos * 2 + 1];
    }
    *pu = u;
    *pv = v;
}

/* W0,H0 are the level-0 dimensions; the image passed in is level `level`: (W0>>level) x (H0>>level). */
void orc_warp_u8(const uint8_t *next, int W0, int H0, int level, int maxLevel, float *const *flowPyr, int mode,
                 float flow_scale, uint8_t *dst)
{
    const int w = W0 >> level, h = H0 >> level;
    memcpy(dst, next, (size_t)w * h);
    if (level == maxLevel - 1) return; /* OptFlowGpu.cu:1918: the coarsest level is not warped */
    for (int i = 0; i < h; i++) {
        for (int j = 0; j < w; j++) {
            float u, v;
            orc_cum_flow(i, j, W0, H0, level, maxLevel, flowPyr, mode, &u, &v);
            if (flow_scale != 1.0f) {
                u *= flow_scale;
                v *= flow_scale;
            }
            const float fx = (float)j + u, fy = (float)i + v;
            if (mode == ORC_WARP_BILINEAR) {
                if (!(fabsf(u) < 32768.0f && fabsf(v) < 32768.0f)) continue;
                const int U = (int)rintf(u * 256.0f), V = (int)rintf(v * 256.0f);
                const int X = j * 256 + U, Y = i * 256 + V;
                if (X < 0 || X > (w - 1) * 256 || Y < 0 || Y > (h - 1) * 256) continue;
                const int x0 = X >> 8, y0 = Y >> 8, wx = X & 255, wy = Y & 255;
                const int x1 = x0 + 1 < w ? x0 + 1 : w - 1, y1 = y0 + 1 < h ? y0 + 1 : h - 1;
                const int p00 = next[(size_t)y0 * w + x0], p01 = next[(size_t)y0 * w + x1];
                const int p10 = next[(size_t)y1 * w + x0], p11 = next[(size_t)y1 * w + x1];
                const int top = (256 - wx) * p00 + wx * p01, bot = (256 - wx) * p10 + wx * p11;
                dst[(size_t)i * w + j] = (uint8_t)(((256 - wy) * top + wy * bot + 32768) >> 16);
            } else {
                /* `int new_pos_x = j + u;` -- float add, truncation toward zero.  Values outside
                 * int range (or NaN) are undefined in C++; x86 yields INT_MIN => "skipped". */
                if (!(fx > -2147483648.0f && fx < 2147483648.0f && fy > -2147483648.0f && fy < 2147483648.0f))
                    continue;
                const int nx = (int)fx, ny = (int)fy;
                if (nx >= w || nx < 0 || ny >= h || ny < 0) continue;
                dst[(size_t)i * w + j] = next[(size_t)ny * w + nx];
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * a7. One level of gpu::calc_opt_flow, OptFlowGpu.cu:1909-1979, on an already-warped next:
 *   Ix = conv(prev, Dx) :1929-1930, Iy = conv(prev, Dy) :1932-1933,
 *   It = conv(next, Dt) - conv(prev, Dt) :1935-1940, five window sums :1947-1960 (win x win;
 *   the entry point hard-codes 19 at :1944-1945, the stage function takes it as an argument),
 *   solve :1964.  flow is (u,v) interleaved, w*h*2 floats.
 * sums_mode picks the fp32-sequential (reference rounding) or exact window sums; they agree
 * bit-for-bit whenever every partial sum stays below 2^24.
 * ------------------------------------------------------------------------------------------- */
void orc_lk_level(const uint8_t *prev, const uint8_t *next, int w, int h, int win, int sums_mode, float *flow)
{
    const size_t n = (size_t)w * h;
    float *Ix = (float *)malloc(n * sizeof(float)), *Iy = (float *)malloc(n * sizeof(float));
    float *It1 = (float *)malloc(n * sizeof(float)), *It2 = (float *)malloc(n * sizeof(float));
    orc_conv_u8_f32(prev, w, h, ORC_DX, 3, 3, Ix);
    orc_conv_u8_f32(prev, w, h, ORC_DY, 3, 3, Iy);
    orc_conv_u8_f32(prev, w, h, ORC_DT, 3, 3, It1);
    orc_conv_u8_f32(next, w, h, ORC_DT, 3, 3, It2);
    float *It = It1;
    orc_sub_f32(It2, It1, n, It);
    if (sums_mode == ORC_SUMS_F32_SEQUENTIAL) {
        float *s[5];
        for (int k = 0; k < 5; k++) s[k] = (float *)malloc(n * sizeof(float));
        orc_srm_f32(Ix, Ix, w, h, win, win, s[0]);
        orc_srm_f32(Iy, Iy, w, h, win, win, s[1]);
        orc_srm_f32(Ix, Iy, w, h, win, win, s[2]);
        orc_srm_f32(Ix, It, w, h, win, win, s[3]);
        orc_srm_f32(Iy, It, w, h, win, win, s[4]);
        orc_solve_f32(s[0], s[1], s[2], s[3], s[4], n, flow);
        for (int k = 0; k < 5; k++) free(s[k]);
    } else {
        int64_t *s[5];
        for (int k = 0; k < 5; k++) s[k] = (int64_t *)malloc(n * sizeof(int64_t));
        orc_srm_exact_i64(Ix, Ix, w, h, win, win, s[0]);
        orc_srm_exact_i64(Iy, Iy, w, h, win, win, s[1]);
        orc_srm_exact_i64(Ix, Iy, w, h, win, win, s[2]);
        orc_srm_exact_i64(Ix, It, w, h, win, win, s[3]);
        orc_srm_exact_i64(Iy, It, w, h, win, win, s[4]);
        orc_solve_i64(s[0], s[1], s[2], s[3], s[4], n, flow);
        for (int k = 0; k < 5; k++) free(s[k]);
    }
    free(Ix);
    free(Iy);
    free(It1);
    free(It2);
}

/* Same, but also hands back the five window sums as doubles (for stage-level parity tests). */
void orc_lk_level_sums(const uint8_t *prev, const uint8_t *next, int w, int h, int win, int64_t *const *sums5)
{
    const size_t n = (size_t)w * h;
    float *Ix = (float *)malloc(n * sizeof(float)), *Iy = (float *)malloc(n * sizeof(float));
    float *It1 = (float *)malloc(n * sizeof(float)), *It2 = (float *)malloc(n * sizeof(float));
    orc_conv_u8_f32(prev, w, h, ORC_DX, 3, 3, Ix);
    orc_conv_u8_f32(prev, w, h, ORC_DY, 3, 3, Iy);
    orc_conv_u8_f32(prev, w, h, ORC_DT, 3, 3, It1);
    orc_conv_u8_f32(next, w, h, ORC_DT, 3, 3, It2);
    orc_sub_f32(It2, It1, n, It1);
    orc_srm_exact_i64(Ix, Ix, w, h, win, win, sums5[0]);
    orc_srm_exact_i64(Iy, Iy, w, h, win, win, sums5[1]);
    orc_srm_exact_i64(Ix, Iy, w, h, win, win, sums5[2]);
    orc_srm_exact_i64(Ix, It1, w, h, win, win, sums5[3]);
    orc_srm_exact_i64(Iy, It1, w, h, win, win, sums5[4]);
    free(Ix);
    free(Iy);
    free(It1);
    free(It2);
}

/* ---------------------------------------------------------------------------------------------
 * a9. The driver loop of main.cu:250-262 for one frame pair given level-0 planar frames:
 *   pyramids of both frames, then for k = levels-1 .. 0: warp next_k by the coarser flows,
 *   one LK level; flow_pyr[k] receives the RESIDUAL flow of level k (a7), as in the reference.
 * flow_pyr[k] must hold (W0>>k)*(H0>>k)*2 floats.  cum_pyr (optional, may be NULL) receives the
 * a10 composition main.cu:136-147: cum_k(i,j) = sum_{m>=k} 2^(m-k) * flow_m[i>>(m-k), j>>(m-k)],
 * accumulated in fp32 coarse-to-fine exactly like shift_back_pyramid does.
 * Returns 0, or -1 on bad arguments.
 * ------------------------------------------------------------------------------------------- */
int orc_flow_pair(const uint8_t *prev0, const uint8_t *next0, int W0, int H0, int levels, int win, int warp_mode,
                  int sums_mode, float flow_scale, float *const *flow_pyr, float *const *cum_pyr)
{
    if (levels < 1 || levels > 16 || (W0 >> (levels - 1)) < 1 || (H0 >> (levels - 1)) < 1) return -1;
    uint8_t *pp[16], *pn[16];
    for (int k = 0; k < levels; k++) {
        const size_t n = (size_t)(W0 >> k) * (H0 >> k);
        pp[k] = (uint8_t *)malloc(n);
        pn[k] = (uint8_t *)malloc(n);
    }
    memcpy(pp[0], prev0, (size_t)W0 * H0);
    memcpy(pn[0], next0, (size_t)W0 * H0);
    orc_gauss_pyramid_u8(pp, W0, H0, levels);
    orc_gauss_pyramid_u8(pn, W0, H0, levels);
    for (int k = levels - 1; k >= 0; k--) {
        const int w = W0 >> k, h = H0 >> k;
        uint8_t *shifted = (uint8_t *)malloc((size_t)w * h);
        orc_warp_u8(pn[k], W0, H0, k, levels, flow_pyr, warp_mode, flow_scale, shifted);
        orc_lk_level(pp[k], shifted, w, h, win, sums_mode, flow_pyr[k]);
        free(shifted);
    }
    if (cum_pyr) {
        for (int k = 0; k < levels; k++) {
            const int w = W0 >> k, h = H0 >> k;
            if (!cum_pyr[k]) continue;
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++) {
                    float u, v;
                    /* sum over m = levels-1 .. k inclusive: treat `k-1` as the current level */
                    float uu = 0.0f, vv = 0.0f;
                    for (int m = levels - 1; m >= k; m--) {
                        const int off = m - k;
                        int ti = i >> off, tj = j >> off;
                        const int wm = W0 >> m, hm = H0 >> m;
                        if (ti > hm - 1) ti = hm - 1;
                        if (tj > wm - 1) tj = wm - 1;
                        const size_t pos = (size_t)ti * wm + tj;
                        uu += (float)(1 << off) * flow_pyr[m][pos * 2];
                        vv += (float)(1 << off) * flow_pyr[m][pos * 2 + 1];
                    }
                    u = uu;
                    v = vv;
                    cum_pyr[k][((size_t)i * w + j) * 2] = u;
                    cum_pyr[k][((size_t)i * w + j) * 2 + 1] = v;
                }
        }
    }
    for (int k = 0; k < levels; k++) {
        free(pp[k]);
        free(pn[k]);
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * 8f rows 1-2: the pre-processing of main.cu:232-240.
 * Grayscale, OptFlowGpu.cu:47-60 (g_grayscale_avg_2d): (c0 + c1 + c2) / 3, integer division, written to
 * all three channels.  3-channel interleaved in and out.
 * ------------------------------------------------------------------------------------------- */
void orc_grayscale_c3(const uint8_t *src, int w, int h, uint8_t *dst)
{
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const int avg = (src[3 * i] + src[3 * i + 1] + src[3 * i + 2]) / 3;
        dst[3 * i] = dst[3 * i + 1] = dst[3 * i + 2] = (uint8_t)avg;
    }
}

/* utils::generate_gaussian_kernel, OptFlowUtils.cpp:68-114: value(i,j) = 1/(2 pi s^2) * e^(-(i^2+j^2)/(2 s^2))
 * mirrored into the four quadrants, then divided by the sum of the whole mask. */
void orc_gaussian_kernel(double sigmaS, int ksize, double *g)
{
    const int hk = ksize >> 1;
    for (int i = 0; i < hk + 1; i++)
        for (int j = 0; j < hk + 1; j++) {
            const double s2 = sigmaS * sigmaS, m = i, n = j;
            const double value = 1.0 / (2.0 * M_PI * s2) * pow(M_E, -0.5 * (n * n + m * m) / s2);
            g[(hk + i) * ksize + hk + j] = g[(hk - i) * ksize + hk - j] = value;
            g[(hk + i) * ksize + hk - j] = g[(hk - i) * ksize + hk + j] = value;
        }
    double sum = 0;
    for (int i = 0; i < ksize * ksize; i++) sum += g[i];
    for (int i = 0; i < ksize * ksize; i++) g[i] /= sum;
}

/* Bilateral pre-filter, OptFlowGpu.cu:1984-2048 (g_bilinear_filter; same arithmetic in
 * OptFlowCPU.cpp:401-465): for every pixel, over the ww x wh window (out-of-image taps skipped, row-major)
 *   n_b = 1/(2 pi sB^2) * e^(-0.5 (gray(tap) - gray(centre))^2 / sB^2),  n_s = spatial mask
 *   wsb += n_b*n_s;  tmp_c += src_c(tap) * n_b * n_s;   dest_c = (unsigned char)(tmp_c / wsb)
 * all in double, range weights from channel 0 of `gray`.  The fma contraction below is the one nvcc emits
 * for the GPU kernel (wsb = fma(n_b, n_s, wsb); tmp = fma(n_s, n_b*src, tmp)).  libm's pow and CUDA's pow
 * may differ in the last bit, so against the GPU this restatement is exact up to rare 1-level flips where
 * tmp/wsb lands within an ulp of an integer; the tests state that tolerance. */
void orc_bilateral_c3(const uint8_t *src, const uint8_t *gray, int w, int h, int ww, int wh, double sigmaS, double sigmaB,
                      uint8_t *dst)
{
    double *mask = (double *)malloc((size_t)ww * wh * sizeof(double));
    orc_gaussian_kernel(sigmaS, ww, mask);
    const int hww = ww >> 1, hwh = wh >> 1;
    const double sB2 = sigmaB * sigmaB;
    const double pref = 1.0 / (2.0 * M_PI * sB2);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const double f_ij = gray[((size_t)y * w + x) * 3];
            double wsb = 0, tmp[3] = {0, 0, 0};
            for (int m = 0; m < wh; m++) {
                const int cy = y - hwh + m;
                if (cy < 0 || cy >= h) continue;
                for (int n = 0; n < ww; n++) {
                    const int cx = x - hww + n;
                    if (cx < 0 || cx >= w) continue;
                    const size_t cp = ((size_t)cy * w + cx) * 3;
                    const double k = (double)gray[cp] - f_ij;
                    const double n_b = pref * pow(M_E, -0.5 * (k * k) / sB2);
                    const double n_s = mask[m * ww + n];
                    wsb = fma(n_b, n_s, wsb);
                    for (int c = 0; c < 3; c++) tmp[c] = fma(n_s, n_b * (double)src[cp + c], tmp[c]);
                }
            }
            for (int c = 0; c < 3; c++) dst[((size_t)y * w + x) * 3 + c] = (uint8_t)(unsigned int)(tmp[c] / wsb);
        }
    free(mask);
}

/* ---------------------------------------------------------------------------------------------
 * Deterministic synthetic frames: 8-bit value noise from an LCG
 * (s = s*1664525 + 1013904223, value = (s>>8) % 256) on a coarse grid of `cell` pixels,
 * bilinearly interpolated, sampled at (x - dx, y - dy), truncated to u8.  Textured enough that
 * det != 0 almost everywhere.  Shared by tests, bench and the golden-vector generator so that
 * no image files need to travel.
 * ------------------------------------------------------------------------------------------- */
void orc_make_frame(uint8_t *img, int w, int h, float dx, float dy, int cell, uint32_t seed)
{
    const int gw = w / cell + 4, gh = h / cell + 4;
    float *g = (float *)malloc((size_t)gw * gh * sizeof(float));
    uint32_t s = seed;
    for (int i = 0; i < gw * gh; i++) {
        s = s * 1664525u + 1013904223u;
        g[i] = (float)((s >> 8) % 256u);
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float fx = ((float)x - dx) / (float)cell + 1.5f, fy = ((float)y - dy) / (float)cell + 1.5f;
            int ix = (int)floorf(fx), iy = (int)floorf(fy);
            if (ix < 0) ix = 0;
            if (ix > gw - 2) ix = gw - 2;
            if (iy < 0) iy = 0;
            if (iy > gh - 2) iy = gh - 2;
            const float ax = fx - (float)ix, ay = fy - (float)iy;
            const float top = (1.0f - ax) * g[iy * gw + ix] + ax * g[iy * gw + ix + 1];
            const float bot = (1.0f - ax) * g[(iy + 1) * gw + ix] + ax * g[(iy + 1) * gw + ix + 1];
            float v = (1.0f - ay) * top + ay * bot;
            if (v < 0.0f) v = 0.0f;
            if (v > 255.0f) v = 255.0f;
            img[(size_t)y * w + x] = (uint8_t)v;
        }
    free(g);
}


/* ---- debug derivative views, main.cu:19-92 ------------------------------------------------
 * orc_conv_u8: gpu::conv_3ch_1ch_tiled, which launches g_conv_3ch_1ch_constant (OptFlowGpu.cu:741-766 ->
 * :380-423): per pixel an INT accumulator, `tmp += src[ch0] * mask` per tap in row-major order -- i.e. int ->
 * float, one fused multiply-add (nvcc contracts it: FFMA in the reference TU's sm_100a SASS), truncation back to
 * int, every tap -- out-of-image taps and zero mask entries skipped, result cast to unsigned char (wraps).
 * src: planar u8 (channel 0 of the reference's 3-channel image). */
void orc_conv_u8(const uint8_t *src, int w, int h, const float *mask, int mw, int mh, uint8_t *dst)
{
    const int hmw = mw >> 1, hmh = mh >> 1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int tmp = 0;
            for (int i = 0; i < mh; i++) {
                const int ty = y - hmh + i;
                if (ty < 0 || ty >= h) continue;
                for (int j = 0; j < mw; j++) {
                    const int tx = x - hmw + j;
                    if (tx < 0 || tx >= w) continue;
                    const float m = mask[i * mw + j];
                    if (m == 0.0f) continue;
                    tmp = (int)fmaf((float)src[(size_t)ty * w + tx], m, (float)tmp);
                }
            }
            dst[(size_t)y * w + x] = (uint8_t)tmp;
        }
}

/* orc_debug_view: one window of showTest (main.cu:19-92) for pyramid level k of size w x h.
 * which 0: Dx_3x3 on cur (:44-56); 1: Dy_3x3 on cur (:76-88); 2: Dt_3x3_n on cur minus on prev with unsigned
 * char wrap-around (cpu::sub_arr, OptFlowCPU.cpp:11-17; :57-75).  Then utils::cleanup_outliers
 * (OptFlowUtils.cpp:5-20: >= 240 or < 20 -> 0, else 255) and utils::upscale_1ch by 2^k (:44-61, nearest).
 * out: (w << k) x (h << k). */
static const float ORC_DTN[9] = {0.0666f, 0.1333f, 0.0666f, 0.1333f, 0.2f, 0.1333f, 0.0666f, 0.1333f, 0.0666f}; /* :25-28 */
int orc_debug_view(const uint8_t *prev, const uint8_t *cur, int w, int h, int k, int which, uint8_t *out)
{
    if (w < 1 || h < 1 || k < 0 || k > 12 || which < 0 || which > 2) return 1;
    const size_t n = (size_t)w * h;
    uint8_t *a = (uint8_t *)malloc(n), *b = (uint8_t *)malloc(n);
    if (!a || !b) {
        free(a);
        free(b);
        return 2;
    }
    if (which == 2) {
        orc_conv_u8(cur, w, h, ORC_DTN, 3, 3, a);
        orc_conv_u8(prev, w, h, ORC_DTN, 3, 3, b);
        for (size_t i = 0; i < n; i++) a[i] = (uint8_t)(a[i] - b[i]);
    } else {
        orc_conv_u8(cur, w, h, which == 0 ? ORC_DX : ORC_DY, 3, 3, a);
    }
    const int s = 1 << k;
    const size_t ow = (size_t)w << k;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t v = a[(size_t)y * w + x];
            const uint8_t c = (v >= 240 || v < 20) ? 0 : 255;
            for (int p = 0; p < s; p++)
                for (int q = 0; q < s; q++) out[((size_t)y * s + p) * ow + (size_t)x * s + q] = c;
        }
    free(a);
    free(b);
    return 0;
}

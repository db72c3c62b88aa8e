// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference, compiled from the sources
// where they lie under /root/reference (oracle/Makefile target `ref`).  TEST INFRASTRUCTURE ONLY:
// the resulting oracle/_ref/libofref.so is used to pin oracle/lk_oracle.c and as the CPU timing
// baseline (bench.py cpu_baseline / --impl reference).  No reference source is copied here; this
// file only forwards to the reference's own declarations (OptFlowGpu.cuh, OptFlowCpu.hpp,
// OptFlowUtils.hpp, kernels.hpp).
#include "OptFlowCpu.hpp"
#include "OptFlowGpu.cuh"
#include "OptFlowUtils.hpp"
#include "kernels.hpp"

#include <cstdlib>
#include <cstring>

extern "C" {

// ---- masks (kernels.cpp) ----
const float *ref_mask_dx() { return Dx_3x3; }
const float *ref_mask_dy() { return Dy_3x3; }
const float *ref_mask_dt() { return Dt_3x3; }
const float *ref_mask_gauss() { return GAUS_KERNEL_3x3; }

// ---- CPU-side functions of the reference (run anywhere) ----
void ref_cpu_gauss_pyramid(unsigned char **pyramid, int w, int h, int levels)
{
    cpu::gauss_pyramid(pyramid, w, h, levels, GAUS_KERNEL_3x3, 3, 3); // main.cu:251 (commented CPU twin)
}
void ref_cpu_downscale_gaussian(unsigned char *src, int w, int h, unsigned char *dest)
{
    cpu::downscale_gaussian(src, w, h, dest, GAUS_KERNEL_3x3, 3, 3);
}
void ref_cpu_shift_back_pyramid(const unsigned char *src, int w, int h, int level, int maxLevel,
                                float **optFlowPyramid, unsigned char *dest)
{
    cpu::shift_back_pyramid(src, w, h, level, maxLevel, optFlowPyramid, dest);
}
void ref_utils_arr_sub_float(float *a, float *b, int n, float *dest) { utils::arr_sub_float(a, b, n, dest); }
void ref_cpu_calc_optical_flow(const unsigned char *prev, unsigned char *next, int w, int h, float **optFlowPyramid,
                               int level, int maxLevel)
{
    cpu::calc_optical_flow(prev, next, w, h, optFlowPyramid, level, maxLevel);
}

// The reference's CPU driver sequence for one pair (main.cu:251,261 commented CPU calls):
// pyramids of both frames, then calc_optical_flow for k = levels-1 .. 0.  Frames are 3-channel
// interleaved u8.  Scratch pyramids are allocated here like main.cu:95-104 does.
void ref_cpu_flow_pair(const unsigned char *prev0, const unsigned char *next0, int w, int h, int levels,
                       float **flow_pyr)
{
    unsigned char **pp = (unsigned char **)malloc(levels * sizeof(unsigned char *));
    unsigned char **pn = (unsigned char **)malloc(levels * sizeof(unsigned char *));
    for (int k = 0; k < levels; k++) {
        size_t n = (size_t)(w >> k) * (h >> k) * 3;
        pp[k] = (unsigned char *)malloc(n);
        pn[k] = (unsigned char *)malloc(n);
    }
    memcpy(pp[0], prev0, (size_t)w * h * 3);
    memcpy(pn[0], next0, (size_t)w * h * 3);
    cpu::gauss_pyramid(pp, w, h, levels, GAUS_KERNEL_3x3, 3, 3);
    cpu::gauss_pyramid(pn, w, h, levels, GAUS_KERNEL_3x3, 3, 3);
    for (int k = levels - 1; k >= 0; k--)
        cpu::calc_optical_flow(pp[k], pn[k], w >> k, h >> k, flow_pyr, k, levels);
    for (int k = 0; k < levels; k++) {
        free(pp[k]);
        free(pn[k]);
    }
    free(pp);
    free(pn);
}

void ref_cpu_grayscale_avg(const unsigned char *src, unsigned char *dest, int w, int h)
{
    cpu::grayscale_avg_cpu(src, dest, w, h);
}
void ref_cpu_bilinear_filter_3ch(unsigned char *src, unsigned char *gray, unsigned char *dest, int w, int h, int ww, int wh,
                                 double sigmaS, double sigmaB)
{
    cpu::bilinear_filter_3ch(src, gray, dest, w, h, ww, wh, sigmaS, sigmaB);
}
void ref_utils_generate_gaussian_kernel(double sigmaS, int kernel_size, double *dest)
{
    utils::generate_gaussian_kernel(sigmaS, kernel_size, dest);
}
void ref_gpu_grayscale_avg(const unsigned char *src, unsigned char *dest, int h, int w) { gpu::grayscale_avg(src, dest, h, w); }
void ref_gpu_bilinear_filter(unsigned char *src, unsigned char *gray, unsigned char *dest, int w, int h, int ww, int wh,
                             double sigmaS, double sigmaB)
{
    gpu::bilinear_filter(src, gray, dest, w, h, ww, wh, sigmaS, sigmaB);
}

// ---- GPU-side functions of the reference (need a GPU; launch-valid only when w,h are multiples
// of 32 and (w/32)*(h/32) <= 1024) ----
// the debug derivative views (main.cu:19-92)
const float *ref_mask_dt_n() { return Dt_3x3_n; }
void ref_gpu_conv_3ch_1ch_tiled(const unsigned char *src, int w, int h, unsigned char *dest, const float *mask, int mw, int mh)
{
    gpu::conv_3ch_1ch_tiled(src, w, h, dest, mask, mw, mh);
}
void ref_cpu_conv_3ch_to_1ch(const unsigned char *src, int w, int h, unsigned char *dest, const float *mask, int mw, int mh)
{
    cpu::conv_3ch_to_1ch(src, w, h, dest, mask, mw, mh);
}
void ref_cpu_sub_arr(unsigned char *a, unsigned char *b, int n, unsigned char *dest) { cpu::sub_arr(a, b, n, dest); }
void ref_utils_cleanup_outliers(unsigned char *src, int w, int h) { utils::cleanup_outliers(src, w, h); }
void ref_utils_upscale_1ch(unsigned char *src, int w, int h, int n, unsigned char *dest) { utils::upscale_1ch(src, w, h, n, dest); }

void ref_gpu_gauss_pyramid(unsigned char **pyramid, int w, int h, int levels)
{
    gpu::gauss_pyramid(pyramid, w, h, levels, GAUS_KERNEL_3x3, 3, 3); // main.cu:250
}
void ref_gpu_conv_3ch_1ch_tiled_uchar_float(const unsigned char *src, int w, int h, float *dest, const float *mask,
                                            int mw, int mh)
{
    gpu::conv_3ch_1ch_tiled_uchar_float(src, w, h, dest, mask, mw, mh);
}
void ref_gpu_srm_1ch_float(const float *a, const float *b, int w, int h, int ww, int wh, float *dest)
{
    gpu::srm_1ch_float(a, b, w, h, ww, wh, dest);
}
void ref_gpu_inverse_matrix_float(float *sIx2, float *sIy2, float *sIxIy, float *sIxIt, float *sIyIt,
                                  float **optFlowPyramid, int level, int w, int h)
{
    gpu::inverse_matrix_float(sIx2, sIy2, sIxIy, sIxIt, sIyIt, optFlowPyramid, level, w, h);
}
void ref_gpu_calc_opt_flow(const unsigned char *prev, unsigned char *next, int w, int h, float **optFlowPyramid,
                           int level, int maxLevel)
{
    gpu::calc_opt_flow(prev, next, w, h, optFlowPyramid, level, maxLevel);
}

// One level of gpu::calc_opt_flow with the window exposed: the same call sequence as
// OptFlowGpu.cu:1929-1964 composed from the reference's own stage functions (the entry point
// hard-codes 19x19 at :1944-1945; the stage functions take the window as an argument).
void ref_gpu_lk_level_win(const unsigned char *prev, const unsigned char *next, int w, int h, int win,
                          float **optFlowPyramid, int level)
{
    size_t n = (size_t)w * h;
    float *Ix = (float *)malloc(n * 4), *Iy = (float *)malloc(n * 4), *It1 = (float *)malloc(n * 4),
          *It2 = (float *)malloc(n * 4);
    float *s[5];
    for (int k = 0; k < 5; k++) s[k] = (float *)malloc(n * 4);
    gpu::conv_3ch_1ch_tiled_uchar_float(prev, w, h, Ix, Dx_3x3, 3, 3);
    gpu::conv_3ch_1ch_tiled_uchar_float(prev, w, h, Iy, Dy_3x3, 3, 3);
    gpu::conv_3ch_1ch_tiled_uchar_float(prev, w, h, It1, Dt_3x3, 3, 3);
    gpu::conv_3ch_1ch_tiled_uchar_float(next, w, h, It2, Dt_3x3, 3, 3);
    utils::arr_sub_float(It2, It1, (int)n, It1);
    gpu::srm_1ch_float(Ix, Ix, w, h, win, win, s[0]);
    gpu::srm_1ch_float(Iy, Iy, w, h, win, win, s[1]);
    gpu::srm_1ch_float(Ix, Iy, w, h, win, win, s[2]);
    gpu::srm_1ch_float(Ix, It1, w, h, win, win, s[3]);
    gpu::srm_1ch_float(Iy, It1, w, h, win, win, s[4]);
    gpu::inverse_matrix_float(s[0], s[1], s[2], s[3], s[4], optFlowPyramid, level, w, h);
    free(Ix);
    free(Iy);
    free(It1);
    free(It2);
    for (int k = 0; k < 5; k++) free(s[k]);
}

} // extern "C"

"""ctypes doorway to the CPU oracle (liblkoracle.so) and the compiled reference (_ref/libofref.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liblkoracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libofref.so")

WARP_AS_WRITTEN, WARP_NEAREST, WARP_BILINEAR = 0, 1, 2
SUMS_F32_SEQUENTIAL, SUMS_EXACT = 0, 1

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and the reference (only where /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liblkoracle.so"])
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _lib = C.CDLL(ORACLE_SO)
    return _lib


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _ptr_array(arrs, t):
    return (t * len(arrs))(*[_p(a, t) for a in arrs])


# ------------------------------------------------------------------ oracle (planar u8) wrappers
def make_frame(w: int, h: int, dx: float = 0.0, dy: float = 0.0, cell: int = 4, seed: int = 1234) -> np.ndarray:
    img = np.empty((h, w), np.uint8)
    lib().orc_make_frame(_p(img, _u8p), w, h, C.c_float(dx), C.c_float(dy), cell, C.c_uint32(seed))
    return img


def pyr_down(src: np.ndarray) -> np.ndarray:
    sh, sw = src.shape
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((sh >> 1, sw >> 1), np.uint8)
    lib().orc_pyr_down_u8(_p(src, _u8p), sw, sh, _p(dst, _u8p))
    return dst


def gauss_pyramid(img0: np.ndarray, levels: int) -> list[np.ndarray]:
    pyr = [np.ascontiguousarray(img0, np.uint8)]
    for _ in range(1, levels):
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def conv(src: np.ndarray, mask: np.ndarray) -> np.ndarray:
    h, w = src.shape
    src = np.ascontiguousarray(src, np.uint8)
    mask = np.ascontiguousarray(mask, np.float32)
    mh, mw = mask.shape
    dst = np.empty((h, w), np.float32)
    lib().orc_conv_u8_f32(_p(src, _u8p), w, h, _p(mask, _f32p), mw, mh, _p(dst, _f32p))
    return dst


DX = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32)
DY = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], np.float32)
DT = np.array([[1, 2, 1], [2, 3, 2], [1, 2, 1]], np.float32)


def srm_f32(a: np.ndarray, b: np.ndarray, ww: int, wh: int) -> np.ndarray:
    h, w = a.shape
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    dst = np.empty((h, w), np.float32)
    lib().orc_srm_f32(_p(a, _f32p), _p(b, _f32p), w, h, ww, wh, _p(dst, _f32p))
    return dst


def srm_exact(a: np.ndarray, b: np.ndarray, ww: int, wh: int) -> np.ndarray:
    h, w = a.shape
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    dst = np.empty((h, w), np.int64)
    lib().orc_srm_exact_i64(_p(a, _f32p), _p(b, _f32p), w, h, ww, wh, _p(dst, _i64p))
    return dst


def solve_f32(sxx, syy, sxy, sxt, syt) -> np.ndarray:
    h, w = sxx.shape
    arrs = [np.ascontiguousarray(x, np.float32) for x in (sxx, syy, sxy, sxt, syt)]
    flow = np.empty((h, w, 2), np.float32)
    lib().orc_solve_f32(*[_p(x, _f32p) for x in arrs], C.c_size_t(h * w), _p(flow, _f32p))
    return flow


def warp(next_k: np.ndarray, W0: int, H0: int, level: int, max_level: int, flow_pyr: list[np.ndarray], mode: int,
         flow_scale: float = 1.0) -> np.ndarray:
    next_k = np.ascontiguousarray(next_k, np.uint8)
    dst = np.empty_like(next_k)
    fl = [np.ascontiguousarray(f, np.float32) for f in flow_pyr]
    lib().orc_warp_u8(_p(next_k, _u8p), W0, H0, level, max_level, _ptr_array(fl, _f32p), mode, C.c_float(flow_scale),
                      _p(dst, _u8p))
    return dst


def lk_level(prev: np.ndarray, nxt: np.ndarray, win: int, sums_mode: int = SUMS_EXACT) -> np.ndarray:
    h, w = prev.shape
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    flow = np.empty((h, w, 2), np.float32)
    lib().orc_lk_level(_p(prev, _u8p), _p(nxt, _u8p), w, h, win, sums_mode, _p(flow, _f32p))
    return flow


def lk_level_sums(prev: np.ndarray, nxt: np.ndarray, win: int) -> list[np.ndarray]:
    h, w = prev.shape
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    sums = [np.empty((h, w), np.int64) for _ in range(5)]
    lib().orc_lk_level_sums(_p(prev, _u8p), _p(nxt, _u8p), w, h, win, _ptr_array(sums, _i64p))
    return sums


def flow_pair(prev0: np.ndarray, next0: np.ndarray, levels: int, win: int, warp_mode: int = WARP_BILINEAR,
              sums_mode: int = SUMS_EXACT, flow_scale: float = 1.0, want_cum: bool = False):
    """Residual flow pyramid (list, level 0 first) and optionally the cumulative pyramid."""
    H0, W0 = prev0.shape
    prev0 = np.ascontiguousarray(prev0, np.uint8)
    next0 = np.ascontiguousarray(next0, np.uint8)
    flows = [np.zeros((H0 >> k, W0 >> k, 2), np.float32) for k in range(levels)]
    cums = [np.zeros((H0 >> k, W0 >> k, 2), np.float32) for k in range(levels)] if want_cum else None
    cum_arg = _ptr_array(cums, _f32p) if want_cum else None
    rc = lib().orc_flow_pair(_p(prev0, _u8p), _p(next0, _u8p), W0, H0, levels, win, warp_mode, sums_mode,
                             C.c_float(flow_scale), _ptr_array(flows, _f32p), cum_arg)
    if rc != 0:
        raise ValueError("orc_flow_pair: bad arguments")
    return (flows, cums) if want_cum else flows


def grayscale_c3(src_c3: np.ndarray) -> np.ndarray:
    h, w, _ = src_c3.shape
    src_c3 = np.ascontiguousarray(src_c3, np.uint8)
    dst = np.empty_like(src_c3)
    lib().orc_grayscale_c3(_p(src_c3, _u8p), w, h, _p(dst, _u8p))
    return dst


def gaussian_kernel(sigma_s: float, ksize: int) -> np.ndarray:
    g = np.empty((ksize, ksize), np.float64)
    lib().orc_gaussian_kernel(C.c_double(sigma_s), ksize, g.ctypes.data_as(C.POINTER(C.c_double)))
    return g


def bilateral_c3(src_c3: np.ndarray, gray_c3: np.ndarray, ww: int, wh: int, sigma_s: float, sigma_b: float) -> np.ndarray:
    h, w, _ = src_c3.shape
    src_c3 = np.ascontiguousarray(src_c3, np.uint8)
    gray_c3 = np.ascontiguousarray(gray_c3, np.uint8)
    dst = np.empty_like(src_c3)
    lib().orc_bilateral_c3(_p(src_c3, _u8p), _p(gray_c3, _u8p), w, h, ww, wh, C.c_double(sigma_s), C.c_double(sigma_b),
                           _p(dst, _u8p))
    return dst


def make_bgr_frame(w: int, h: int, dx: float = 0.0, dy: float = 0.0, cell: int = 8, seed: int = 1) -> np.ndarray:
    """Three different value-noise channels (a colour frame for the grayscale / frame-loop tests)."""
    return np.ascontiguousarray(np.stack([make_frame(w, h, dx, dy, cell, seed + 17 * c) for c in range(3)], axis=2))


# ------------------------------------------------------------------ helpers shared by tests
def to_c3(img: np.ndarray) -> np.ndarray:
    """Planar gray -> the reference's 3-equal-channel interleaved layout (OptFlowGpu.cu:58-59)."""
    return np.ascontiguousarray(np.repeat(img[:, :, None], 3, axis=2))


# ------------------------------------------------------------------ compiled-reference wrappers
def ref_cpu_gauss_pyramid_c3(img0_c3: np.ndarray, levels: int) -> list[np.ndarray]:
    h, w, _ = img0_c3.shape
    pyr = [np.ascontiguousarray(img0_c3, np.uint8)] + [np.zeros((h >> k, w >> k, 3), np.uint8) for k in range(1, levels)]
    ref().ref_cpu_gauss_pyramid(_ptr_array(pyr, _u8p), w, h, levels)
    return pyr


def ref_cpu_shift_back_c3(next_c3: np.ndarray, level: int, max_level: int, flow_pyr: list[np.ndarray],
                          prefill: int = 0xAB) -> np.ndarray:
    h, w, _ = next_c3.shape
    dst = np.full((h, w, 3), prefill, np.uint8)
    fl = [np.ascontiguousarray(f, np.float32) for f in flow_pyr]
    ref().ref_cpu_shift_back_pyramid(_p(np.ascontiguousarray(next_c3), _u8p), w, h, level, max_level,
                                     _ptr_array(fl, _f32p), _p(dst, _u8p))
    return dst


def ref_arr_sub(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32).copy()
    b = np.ascontiguousarray(b, np.float32).copy()
    dst = np.empty_like(a)
    ref().ref_utils_arr_sub_float(_p(a, _f32p), _p(b, _f32p), a.size, _p(dst, _f32p))
    return dst


def ref_cpu_flow_pair_c3(prev_c3: np.ndarray, next_c3: np.ndarray, levels: int) -> list[np.ndarray]:
    h, w, _ = prev_c3.shape
    flows = [np.zeros((h >> k, w >> k, 2), np.float32) for k in range(levels)]
    ref().ref_cpu_flow_pair(_p(np.ascontiguousarray(prev_c3), _u8p), _p(np.ascontiguousarray(next_c3), _u8p), w, h,
                            levels, _ptr_array(flows, _f32p))
    return flows


def ref_cpu_grayscale(src_c3: np.ndarray) -> np.ndarray:
    h, w, _ = src_c3.shape
    dst = np.empty_like(src_c3)
    ref().ref_cpu_grayscale_avg(_p(np.ascontiguousarray(src_c3), _u8p), _p(dst, _u8p), w, h)
    return dst


def ref_gaussian_kernel(sigma_s: float, ksize: int) -> np.ndarray:
    g = np.empty((ksize, ksize), np.float64)
    ref().ref_utils_generate_gaussian_kernel(C.c_double(sigma_s), ksize, g.ctypes.data_as(C.POINTER(C.c_double)))
    return g


def ref_cpu_bilateral(src_c3, gray_c3, ww, wh, sigma_s, sigma_b) -> np.ndarray:
    h, w, _ = src_c3.shape
    dst = np.empty_like(src_c3)
    ref().ref_cpu_bilinear_filter_3ch(_p(np.ascontiguousarray(src_c3).copy(), _u8p), _p(np.ascontiguousarray(gray_c3).copy(), _u8p),
                                      _p(dst, _u8p), w, h, ww, wh, C.c_double(sigma_s), C.c_double(sigma_b))
    return dst


def ref_gpu_grayscale(src_c3: np.ndarray) -> np.ndarray:
    h, w, _ = src_c3.shape
    dst = np.empty_like(src_c3)
    ref().ref_gpu_grayscale_avg(_p(np.ascontiguousarray(src_c3), _u8p), _p(dst, _u8p), h, w)
    return dst


def ref_gpu_bilateral(src_c3, gray_c3, ww, wh, sigma_s, sigma_b) -> np.ndarray:
    h, w, _ = src_c3.shape
    dst = np.zeros_like(src_c3)
    ref().ref_gpu_bilinear_filter(_p(np.ascontiguousarray(src_c3).copy(), _u8p), _p(np.ascontiguousarray(gray_c3).copy(), _u8p),
                                  _p(dst, _u8p), w, h, ww, wh, C.c_double(sigma_s), C.c_double(sigma_b))
    return dst


def ref_gpu_conv(src_c3: np.ndarray, mask: np.ndarray) -> np.ndarray:
    h, w, _ = src_c3.shape
    mask = np.ascontiguousarray(mask, np.float32)
    dst = np.empty((h, w), np.float32)
    ref().ref_gpu_conv_3ch_1ch_tiled_uchar_float(_p(np.ascontiguousarray(src_c3), _u8p), w, h, _p(dst, _f32p),
                                                 _p(mask, _f32p), mask.shape[1], mask.shape[0])
    return dst


def ref_gpu_srm(a: np.ndarray, b: np.ndarray, ww: int, wh: int) -> np.ndarray:
    h, w = a.shape
    dst = np.empty((h, w), np.float32)
    ref().ref_gpu_srm_1ch_float(_p(np.ascontiguousarray(a, np.float32), _f32p),
                                _p(np.ascontiguousarray(b, np.float32), _f32p), w, h, ww, wh, _p(dst, _f32p))
    return dst


def ref_gpu_inverse(sxx, syy, sxy, sxt, syt) -> np.ndarray:
    h, w = sxx.shape
    arrs = [np.ascontiguousarray(x, np.float32).copy() for x in (sxx, syy, sxy, sxt, syt)]
    flow = np.full((h, w, 2), -7.0, np.float32)
    pyr = _ptr_array([flow], _f32p)
    ref().ref_gpu_inverse_matrix_float(*[_p(x, _f32p) for x in arrs], pyr, 0, w, h)
    return flow


def ref_gpu_lk_level_win(prev_c3: np.ndarray, next_c3: np.ndarray, win: int) -> np.ndarray:
    h, w, _ = prev_c3.shape
    flow = np.full((h, w, 2), -7.0, np.float32)
    pyr = _ptr_array([flow], _f32p)
    ref().ref_gpu_lk_level_win(_p(np.ascontiguousarray(prev_c3), _u8p), _p(np.ascontiguousarray(next_c3), _u8p), w, h,
                               win, pyr, 0)
    return flow


def ref_gpu_gauss_pyramid_c3(img0_c3: np.ndarray, levels: int) -> list[np.ndarray]:
    h, w, _ = img0_c3.shape
    pyr = [np.ascontiguousarray(img0_c3, np.uint8)] + [np.zeros((h >> k, w >> k, 3), np.uint8) for k in range(1, levels)]
    ref().ref_gpu_gauss_pyramid(_ptr_array(pyr, _u8p), w, h, levels)
    return pyr


# ---- flow composition and arrows, literal restatement of main.cu:114-174 (numpy) ----------------
def compose_total(flows, level: int = 0) -> np.ndarray:
    """main.cu:136-147: for every pixel of `level`, u = v = 0 (float); for k = levels-1 .. level:
    u += (double)(1 << (k-level)) * flow_k[(i >> (k-level)) * (w >> (k-level)) + (j >> (k-level))], stored back to
    float each time.  Indices past a coarser level (odd sizes; the reference reads out of bounds there) are
    clamped to its last row / column."""
    levels = len(flows)
    h, w = flows[level].shape[:2]
    ii, jj = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    tot = np.zeros((h, w, 2), np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        for k in range(levels - 1, level - 1, -1):
            sc = k - level
            fk = flows[k]
            yi = np.minimum(ii >> sc, fk.shape[0] - 1)
            xi = np.minimum(jj >> sc, fk.shape[1] - 1)
            tot = (tot.astype(np.float64) + np.float64(1 << sc) * fk[yi, xi].astype(np.float64)).astype(np.float32)
    return tot


def flow_arrows(flows, level: int, arrow_res: int) -> np.ndarray:
    """main.cu:124-171 without the drawing: (n, 4) int32 (x0, y0, x1, y1)."""
    h, w = flows[level].shape[:2]
    step = w // arrow_res
    tot = compose_total(flows, level)
    out = []
    for i in range(0, h, step):
        for j in range(0, w, step):
            u, v = np.float32(tot[i, j, 0]), np.float32(tot[i, j, 1])
            if u > step:
                u = np.float32(step)
            elif u < -step:
                u = np.float32(-step)
            if v > step:
                v = np.float32(step)
            elif v < -step:
                v = np.float32(-step)
            if np.isnan(u) or np.isnan(v):  # (int)NaN is undefined in the reference; x86 yields INT_MIN -> dropped
                continue
            ni, nj = int(np.float32(v + np.float32(i))), int(np.float32(u + np.float32(j)))
            if ni < 0 or nj < 0:
                continue
            out.append((j, i, nj, ni))
    return np.array(out, np.int32).reshape(-1, 4)


def read_flo(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        tag = np.frombuffer(f.read(4), np.float32)[0]
        w, h = np.frombuffer(f.read(8), np.int32)
        assert tag == np.float32(202021.25)
        return np.frombuffer(f.read(), np.float32).reshape(h, w, 2).copy()


# ---- debug derivative views (main.cu:19-92) ------------------------------------------------------
VIEW_X, VIEW_Y, VIEW_T = 0, 1, 2


def conv_u8(src: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """gpu::conv_3ch_1ch_tiled (OptFlowGpu.cu:741-766): u8 planar in, u8 out, int accumulator truncated every tap."""
    h, w = src.shape
    src = np.ascontiguousarray(src, np.uint8)
    mask = np.ascontiguousarray(mask, np.float32)
    dst = np.empty((h, w), np.uint8)
    lib().orc_conv_u8(_p(src, _u8p), w, h, _p(mask, _f32p), mask.shape[1], mask.shape[0], _p(dst, _u8p))
    return dst


def debug_view(prev: np.ndarray, cur: np.ndarray, k: int, which: int) -> np.ndarray:
    """One window of showTest for level k images (planar u8): thresholded derivative, upscaled by 2^k."""
    h, w = cur.shape
    prev = np.ascontiguousarray(prev, np.uint8)
    cur = np.ascontiguousarray(cur, np.uint8)
    out = np.empty((h << k, w << k), np.uint8)
    rc = lib().orc_debug_view(_p(prev, _u8p), _p(cur, _u8p), w, h, k, which, _p(out, _u8p))
    if rc != 0:
        raise ValueError("orc_debug_view: bad arguments")
    return out


def _c3(img: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.repeat(img[:, :, None], 3, axis=2))


def ref_mask_dt_n() -> np.ndarray:
    r = ref()
    r.ref_mask_dt_n.restype = C.POINTER(C.c_float)
    return np.ctypeslib.as_array(r.ref_mask_dt_n(), shape=(9,)).reshape(3, 3).copy()


def ref_cpu_conv_u8(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """cpu::conv_3ch_to_1ch (OptFlowCPU.cpp:75-110), the host twin of the u8 convolution (no fused multiply-add)."""
    h, w = img.shape
    src, mask = _c3(img), np.ascontiguousarray(mask, np.float32)
    dst = np.empty((h, w), np.uint8)
    ref().ref_cpu_conv_3ch_to_1ch(_p(src, _u8p), w, h, _p(dst, _u8p), _p(mask, _f32p), mask.shape[1], mask.shape[0])
    return dst


def ref_gpu_conv_u8(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """gpu::conv_3ch_1ch_tiled itself (needs a GPU)."""
    h, w = img.shape
    src, mask = _c3(img), np.ascontiguousarray(mask, np.float32)
    dst = np.empty((h, w), np.uint8)
    ref().ref_gpu_conv_3ch_1ch_tiled(_p(src, _u8p), w, h, _p(dst, _u8p), _p(mask, _f32p), mask.shape[1], mask.shape[0])
    return dst


def ref_debug_view(prev: np.ndarray, cur: np.ndarray, k: int, which: int, conv=None) -> np.ndarray:
    """showTest composed from the reference's own functions: conv (GPU by default), cpu::sub_arr,
    utils::cleanup_outliers, utils::upscale_1ch."""
    conv = conv or ref_gpu_conv_u8
    h, w = cur.shape
    masks = {VIEW_X: np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32),
             VIEW_Y: np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], np.float32), VIEW_T: ref_mask_dt_n()}
    a = conv(cur, masks[which])
    if which == VIEW_T:
        b = conv(prev, masks[which])
        ref().ref_cpu_sub_arr(_p(a, _u8p), _p(b, _u8p), w * h, _p(b, _u8p))  # main.cu:65: dest = second operand
        a = b
    a = np.ascontiguousarray(a)
    ref().ref_utils_cleanup_outliers(_p(a, _u8p), w, h)
    out = np.empty((h << k, w << k), np.uint8)
    ref().ref_utils_upscale_1ch(_p(a, _u8p), w, h, k, _p(out, _u8p))
    return out

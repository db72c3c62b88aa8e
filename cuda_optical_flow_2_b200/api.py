"""Host-side mirror of the reference's `namespace gpu` interface for the pyramidal LK path.

Function names, argument order and meaning follow OptFlowGpu.cuh (reference): `gauss_pyramid` (:21),
`calc_opt_flow` (:33), `conv_3ch_1ch_tiled_uchar_float` (:17), `srm_1ch_float` (:25),
`inverse_matrix_float` (:31).  Host images are numpy u8 arrays in the reference layout
(h, w, 3 interleaved channels); flow is float32 (h, w, 2).  Everything computes on the GPU through
the C ABI in include/ofb200.h -- there is no CPU path here.

The device-resident entry points (`flow_pairs_device`, `lk_level_device`, `pyr_down_device`) take
torch CUDA tensors only as owners of device memory; torch does none of the arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import SOLVE_EXACT, SOLVE_FAST, WARP_AS_WRITTEN, WARP_BILINEAR, WARP_NEAREST, OfbError, OfbParams  # noqa: F401

REFERENCE_WINDOW = 19  # OptFlowGpu.cu:1944-1945


def _u8(a: np.ndarray) -> np.ndarray:
    if a.dtype != np.uint8 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("expected a C-contiguous uint8 array")
    return a


def _f32(a: np.ndarray) -> np.ndarray:
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("expected a C-contiguous float32 array")
    return a


def _ptrs(arrs: Sequence[np.ndarray], t):
    return (t * len(arrs))(*[a.ctypes.data_as(t) for a in arrs])


def align_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


class Context:
    """One per GPU (ofb_ctx): owns the device workspace reused across calls."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        L.check(self._lib.ofb_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.ofb_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self) -> int:
        v = C.c_int()
        L.check(self._lib.ofb_ctx_sm_count(self._h, C.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        v = C.c_ulonglong()
        L.check(self._lib.ofb_ctx_launch_count(self._h, C.byref(v)))
        return v.value

    @property
    def solve(self) -> int:
        """SOLVE_EXACT (default; the reference's double-precision operation order, bit-identical flow) or SOLVE_FAST
        (exact 64-bit integer determinant and numerators, one float reciprocal; |d| <= 1e-4 px + 1e-5 |ref| per level)."""
        v = C.c_int()
        L.check(self._lib.ofb_ctx_get_solve(self._h, C.byref(v)))
        return v.value

    @solve.setter
    def solve(self, mode: int) -> None:
        L.check(self._lib.ofb_ctx_set_solve(self._h, int(mode)))

    @property
    def host_threads(self) -> int:
        """Host threads that extract channel 0 of 3-channel host images in `flow_pairs_host` (0: the device does it
        after uploading all three channels)."""
        v = C.c_int(0)
        L.check(self._lib.ofb_ctx_get_host_threads(self._h, C.byref(v)))
        return v.value

    @host_threads.setter
    def host_threads(self, n: int) -> None:
        L.check(self._lib.ofb_ctx_set_host_threads(self._h, int(n)))

    def reserve_pairs(self, w: int, h: int, levels: int, win: int, n_pairs: int) -> None:
        """Size the context workspace for flow_pairs_device calls of this shape now (a later growth would invalidate
        CUDA graphs captured from earlier calls)."""
        p = OfbParams(w, h, levels, win, WARP_BILINEAR, 1.0, n_pairs)
        L.check(self._lib.ofb_ctx_reserve_pairs(self._h, C.byref(p)))

    def profile_enable(self, on: bool = True) -> None:
        """CUDA-event timing around each fused-LK launch (tag = level) and the pyramid build."""
        L.check(self._lib.ofb_ctx_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, tag: int) -> tuple[float, int]:
        """(summed milliseconds, number of records) for a tag: a pyramid level or PROFILE_PYRAMID."""
        ms, n = C.c_double(), C.c_ulonglong()
        L.check(self._lib.ofb_ctx_profile_read(self._h, int(tag), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---------------------------------------------------------------- reference-named host API
    def gauss_pyramid(self, pyramid: Sequence[np.ndarray], w: int, h: int, levels: int) -> None:
        """gpu::gauss_pyramid (OptFlowGpu.cu:1262): fills pyramid[1..levels-1] from pyramid[0] in place.
        pyramid[k] is a (h>>k, w>>k, 3) uint8 array."""
        for k in range(levels):
            if pyramid[k].shape != (h >> k, w >> k, 3):
                raise ValueError(f"pyramid[{k}] has shape {pyramid[k].shape}, expected {(h >> k, w >> k, 3)}")
        arrs = [_u8(p) for p in pyramid[:levels]]
        L.check(self._lib.ofb_gauss_pyramid_host_u8c3(self._h, _ptrs(arrs, L.u8p), w, h, levels))

    def calc_opt_flow(self, prev: np.ndarray, next: np.ndarray, w: int, h: int, optFlowPyramid: Sequence[np.ndarray],
                      level: int, maxLevel: int, win: int = REFERENCE_WINDOW, warp_mode: int = WARP_AS_WRITTEN,
                      flow_scale: float = 1.0) -> None:
        """gpu::calc_opt_flow (OptFlowGpu.cu:1909): one level; reads optFlowPyramid[k > level], writes
        optFlowPyramid[level] in place.  Defaults reproduce the reference (window 19, its warp as written)."""
        if prev.shape != (h, w, 3) or next.shape != (h, w, 3):
            raise ValueError("prev/next must be (h, w, 3) uint8")
        arrs = [_f32(f) for f in optFlowPyramid[:maxLevel]]
        for k in range(level, maxLevel):
            exp = (h >> (k - level), w >> (k - level), 2)
            if arrs[k].shape != exp:
                raise ValueError(f"optFlowPyramid[{k}] has shape {arrs[k].shape}, expected {exp}")
        ptrs = (L.f32p * maxLevel)(*[a.ctypes.data_as(L.f32p) for a in arrs])
        L.check(self._lib.ofb_calc_opt_flow_host_u8c3(self._h, _u8(prev).ctypes.data_as(L.u8p),
                                                      _u8(next).ctypes.data_as(L.u8p), w, h, ptrs, level, maxLevel, win,
                                                      warp_mode, C.c_float(flow_scale)))

    def conv_3ch_1ch_tiled_uchar_float(self, src: np.ndarray, w: int, h: int, mask: np.ndarray, mw: int, mh: int) -> np.ndarray:
        """gpu::conv_3ch_1ch_tiled_uchar_float (OptFlowGpu.cu:1100)."""
        if src.shape != (h, w, 3):
            raise ValueError("src must be (h, w, 3) uint8")
        mask = np.ascontiguousarray(mask, np.float32).reshape(-1)
        if mask.size != mw * mh:
            raise ValueError("mask size does not match mw*mh")
        dst = np.empty((h, w), np.float32)
        L.check(self._lib.ofb_conv_3ch_1ch_u8_f32_host(self._h, _u8(src).ctypes.data_as(L.u8p), w, h,
                                                       dst.ctypes.data_as(L.f32p), mask.ctypes.data_as(L.f32p), mw, mh))
        return dst

    def srm_1ch_float(self, arr1: np.ndarray, arr2: np.ndarray, w: int, h: int, ww: int, wh: int) -> np.ndarray:
        """gpu::srm_1ch_float (OptFlowGpu.cu:1597)."""
        if arr1.shape != (h, w) or arr2.shape != (h, w):
            raise ValueError("arr1/arr2 must be (h, w) float32")
        dst = np.empty((h, w), np.float32)
        L.check(self._lib.ofb_srm_1ch_f32_host(self._h, _f32(arr1).ctypes.data_as(L.f32p), _f32(arr2).ctypes.data_as(L.f32p),
                                               w, h, ww, wh, dst.ctypes.data_as(L.f32p)))
        return dst

    def inverse_matrix_float(self, sumIx2, sumIy2, sumIxIy, sumIxIt, sumIyIt, optFlowPyramid: Sequence[np.ndarray],
                             level: int, w: int, h: int) -> None:
        """gpu::inverse_matrix_float (OptFlowGpu.cu:1858): writes optFlowPyramid[level] in place."""
        sums = [_f32(s) for s in (sumIx2, sumIy2, sumIxIy, sumIxIt, sumIyIt)]
        for s in sums:
            if s.shape != (h, w):
                raise ValueError("sums must be (h, w) float32")
        arrs = [_f32(f) for f in optFlowPyramid[:level + 1]]
        if arrs[level].shape != (h, w, 2):
            raise ValueError("optFlowPyramid[level] must be (h, w, 2) float32")
        ptrs = (L.f32p * (level + 1))(*[a.ctypes.data_as(L.f32p) for a in arrs])
        L.check(self._lib.ofb_inverse_matrix_f32_host(self._h, *[s.ctypes.data_as(L.f32p) for s in sums], ptrs, level, w, h))

    # ---- debug derivative views: showTest (main.cu:19-92) without the windows ----
    def conv_3ch_1ch_tiled(self, src: np.ndarray, w: int, h: int, mask: np.ndarray, mw: int, mh: int) -> np.ndarray:
        """gpu::conv_3ch_1ch_tiled: (h, w, 3) u8 in, (h, w) u8 out."""
        mask = np.ascontiguousarray(mask, np.float32).reshape(-1)
        dst = np.empty((h, w), np.uint8)
        L.check(self._lib.ofb_conv_3ch_1ch_u8_u8_host(self._h, _u8(src).ctypes.data_as(L.u8p), w, h, dst.ctypes.data_as(L.u8p),
                                                      mask.ctypes.data_as(L.f32p), mw, mh))
        return dst

    def debug_view(self, prev_level: Optional[np.ndarray], cur_level: np.ndarray, w: int, h: int, level: int, which: int) -> np.ndarray:
        """One window of showTest: thresholded derivative of a pyramid level, upscaled by 2^level."""
        out = np.empty((h << level, w << level), np.uint8)
        pp = _u8(prev_level).ctypes.data_as(L.u8p) if prev_level is not None else None
        L.check(self._lib.ofb_debug_view_host_u8c3(self._h, pp, _u8(cur_level).ctypes.data_as(L.u8p), w, h, level, which,
                                                   out.ctypes.data_as(L.u8p)))
        return out

    # ---- flow composition and export: the headless part of visualizeFlowField (main.cu:114-174) ----
    def compose_flow(self, flow_pyramid: Sequence[np.ndarray], w: int, h: int, levels: int, level: int = 0) -> np.ndarray:
        """Total flow of `level` from the residual pyramid (composition rule main.cu:136-147)."""
        flows = [_f32(f) for f in flow_pyramid]
        out = np.empty((h >> level, w >> level, 2), np.float32)
        L.check(self._lib.ofb_compose_flow_host(self._h, _ptrs(flows, L.f32p), w, h, levels, level, out.ctypes.data_as(L.f32p)))
        return out

    def flow_arrows(self, flow_pyramid: Sequence[np.ndarray], w: int, h: int, levels: int, level: int, arrow_res: int) -> np.ndarray:
        """(n, 4) int32 array of (x0, y0, x1, y1): the arrows main.cu:125-171 draws."""
        flows = [_f32(f) for f in flow_pyramid]
        # one arrow per grid point at most: the grid step is (w >> level) // arrow_res in both directions (main.cu:125-133)
        step = max((w >> level) // max(arrow_res, 1), 1)
        cap = -(-(w >> level) // step) * -(-(h >> level) // step)
        buf = np.empty((cap, 4), np.int32)
        n = C.c_int(0)
        L.check(self._lib.ofb_flow_arrows_host(self._h, _ptrs(flows, L.f32p), w, h, levels, level, arrow_res,
                                               buf.ctypes.data_as(L.i32p), cap, C.byref(n)))
        if n.value > cap:  # (the C side reports the full count and truncates the list: ask again with room for all)
            cap = n.value
            buf = np.empty((cap, 4), np.int32)
            L.check(self._lib.ofb_flow_arrows_host(self._h, _ptrs(flows, L.f32p), w, h, levels, level, arrow_res,
                                                   buf.ctypes.data_as(L.i32p), cap, C.byref(n)))
        return buf[: min(n.value, cap)].copy()

    def grayscale_avg(self, src: np.ndarray, h: int, w: int) -> np.ndarray:
        """gpu::grayscale_avg(src, dest, h, w) (OptFlowGpu.cu:75; height before width like the reference)."""
        if src.shape != (h, w, 3):
            raise ValueError("src must be (h, w, 3) uint8")
        dst = np.empty_like(src)
        L.check(self._lib.ofb_grayscale_avg_host_u8c3(self._h, _u8(src).ctypes.data_as(L.u8p), dst.ctypes.data_as(L.u8p), h, w))
        return dst

    def bilinear_filter(self, src: np.ndarray, gray: np.ndarray, w: int, h: int, ww: int, wh: int, sigmaS: float,
                        sigmaB: float) -> np.ndarray:
        """gpu::bilinear_filter (OptFlowGpu.cu:2050): the bilateral pre-filter of main.cu:240."""
        if src.shape != (h, w, 3) or gray.shape != (h, w, 3):
            raise ValueError("src/gray must be (h, w, 3) uint8")
        dst = np.empty_like(src)
        L.check(self._lib.ofb_bilinear_filter_host_u8c3(self._h, _u8(src).ctypes.data_as(L.u8p), _u8(gray).ctypes.data_as(L.u8p),
                                                        dst.ctypes.data_as(L.u8p), w, h, ww, wh, float(sigmaS), float(sigmaB)))
        return dst

    def bilateral_planar_device(self, gray, w: int, ww: int, sigmaS: float, sigmaB: float, dst=None, stream: int = 0):
        """Bilateral filter of a planar (h, pitch) uint8 CUDA tensor (src == gray)."""
        import torch

        h, pitch = gray.shape
        if dst is None:
            dst = torch.zeros_like(gray)
        L.check(self._lib.ofb_bilateral_planar_device(self._h, gray.data_ptr(), pitch, w, h, ww, ww, float(sigmaS), float(sigmaB),
                                                      dst.data_ptr(), dst.stride(0), C.c_void_p(stream)))
        return dst

    def open_stream(self, w: int, h: int, levels: int, win: int, warp_mode: int = WARP_BILINEAR, flow_scale: float = 1.0,
                    bil_win: int = 0, bil_sigma_s: float = 2.0, bil_sigma_b: float = 10.0) -> "FrameStream":
        """Frame sequence in the role of main.cu:222-275 (defaults of main.cu:236-240 for the pre-filter)."""
        return FrameStream(self, w, h, levels, win, warp_mode, flow_scale, bil_win, bil_sigma_s, bil_sigma_b)

    # ---------------------------------------------------------------- whole-pair host API
    def flow_pairs_host(self, prev: np.ndarray, next: np.ndarray, levels: int, win: int,
                        warp_mode: int = WARP_BILINEAR, flow_scale: float = 1.0,
                        out: Optional[Sequence[np.ndarray]] = None) -> list[np.ndarray]:
        """The loop of main.cu:246-262 for a batch with host buffers.  prev/next: (n, h, w) gray or
        (n, h, w, 3) reference layout (a single pair may omit n).  Returns the residual flow of every
        level, level 0 first, each (n, h>>k, w>>k, 2)."""
        if prev.ndim == 4:
            if prev.shape[3] != 3:
                raise ValueError("4-d input must be (n, h, w, 3)")
            channels = 3
            n, h, w = prev.shape[0], prev.shape[1], prev.shape[2]
        elif prev.ndim == 3:
            channels = 1
            n, h, w = prev.shape
        elif prev.ndim == 2:
            channels = 1
            n, (h, w) = 1, prev.shape
        else:
            raise ValueError("prev must be (n,h,w), (h,w) or (n,h,w,3)")
        if next.shape != prev.shape:
            raise ValueError("prev and next differ in shape")
        p = OfbParams(w, h, levels, win, warp_mode, flow_scale, n)
        if out is None:
            out = [np.empty((n, h >> k, w >> k, 2), np.float32) for k in range(levels)]
        ptrs = (L.f32p * levels)(*[_f32(o).ctypes.data_as(L.f32p) for o in out])
        L.check(self._lib.ofb_flow_pairs_host(self._h, C.byref(p), _u8(prev).ctypes.data_as(L.u8p),
                                              _u8(next).ctypes.data_as(L.u8p), channels, ptrs))
        return list(out)

    def total_flow_pairs_host(self, prev: np.ndarray, next: np.ndarray, levels: int, win: int,
                              warp_mode: int = WARP_BILINEAR, flow_scale: float = 1.0,
                              out: Optional[np.ndarray] = None) -> np.ndarray:
        """As flow_pairs_host, but only the level-0 composition of main.cu:136-147 (the total flow, (n, h, w, 2))
        leaves the device: 8 bytes per pixel over PCIe instead of 10.5."""
        if prev.ndim == 4:
            if prev.shape[3] != 3:
                raise ValueError("4-d input must be (n, h, w, 3)")
            channels, (n, h, w) = 3, prev.shape[:3]
        elif prev.ndim == 3:
            channels, (n, h, w) = 1, prev.shape
        else:
            raise ValueError("prev must be (n,h,w) or (n,h,w,3)")
        if next.shape != prev.shape:
            raise ValueError("prev and next differ in shape")
        p = OfbParams(w, h, levels, win, warp_mode, flow_scale, n)
        if out is None:
            out = np.empty((n, h, w, 2), np.float32)
        L.check(self._lib.ofb_flow_pairs_host_ex(self._h, C.byref(p), _u8(prev).ctypes.data_as(L.u8p),
                                                 _u8(next).ctypes.data_as(L.u8p), channels, None,
                                                 _f32(out).ctypes.data_as(L.f32p)))
        return out

    # ---------------------------------------------------------------- device-resident API
    def flow_pairs_device(self, prev, next, w: int, levels: int, win: int, warp_mode: int = WARP_BILINEAR,
                          flow_scale: float = 1.0, flows=None, total_flow=None, stream: int = 0):
        """prev/next: torch.uint8 CUDA tensors (n, h, pitch) with pitch % 16 == 0 and pitch >= w.
        Returns the list of residual-flow tensors (n, h>>k, w>>k, 2); pass `flows` to reuse buffers."""
        import torch

        if prev.dtype != torch.uint8 or prev.dim() != 3 or not prev.is_contiguous() or prev.shape != next.shape:
            raise ValueError("prev/next must be contiguous uint8 (n, h, pitch) CUDA tensors of equal shape")
        n, h, pitch = prev.shape
        if flows is None:
            flows = [torch.empty((n, h >> k, w >> k, 2), dtype=torch.float32, device=prev.device) for k in range(levels)]
        p = OfbParams(w, h, levels, win, warp_mode, flow_scale, n)
        ptrs = (C.c_void_p * levels)(*[f.data_ptr() for f in flows])
        L.check(self._lib.ofb_flow_pairs_device(self._h, C.byref(p), prev.data_ptr(), next.data_ptr(), pitch, pitch * h,
                                                ptrs, total_flow.data_ptr() if total_flow is not None else None,
                                                C.c_void_p(stream)))
        return flows

    def pyr_down_device(self, src, sw: int, dst=None, stream: int = 0):
        """src: torch.uint8 (n, sh, pitch) -> dst (n, sh>>1, pitch_d)."""
        import torch

        n, sh, sp = src.shape
        dw, dh = sw >> 1, sh >> 1
        if dst is None:
            dst = torch.zeros((n, dh, align_up(dw, 64)), dtype=torch.uint8, device=src.device)
        dp = dst.shape[2]
        L.check(self._lib.ofb_pyr_down_device(self._h, src.data_ptr(), sp, sp * sh, sw, sh, dst.data_ptr(), dp, dp * dh, n,
                                              C.c_void_p(stream)))
        return dst

    def pyr_down2_device(self, src, sw: int, stream: int = 0):
        """Two pyramid steps in one launch: src (n, sh, pitch) -> (dst1 (n, sh>>1, pitch1), dst2 (n, sh>>2, pitch2))."""
        import torch

        n, sh, sp = src.shape
        d1 = torch.zeros((n, sh >> 1, align_up(sw >> 1, 64)), dtype=torch.uint8, device=src.device)
        d2 = torch.zeros((n, sh >> 2, align_up(sw >> 2, 64)), dtype=torch.uint8, device=src.device)
        L.check(self._lib.ofb_pyr_down2_device(self._h, src.data_ptr(), sp, sp * sh, sw, sh, d1.data_ptr(), d1.shape[2],
                                               d1.shape[2] * d1.shape[1], d2.data_ptr(), d2.shape[2], d2.shape[2] * d2.shape[1],
                                               n, C.c_void_p(stream)))
        return d1, d2

    def pyr_down_strip_device(self, src, sw: int, src_y_off: int, dst, dst_y0: int, dst_y1: int, stream: int = 0):
        """Row-strip pyramid step: src (rows, pitch) holds global rows from src_y_off; dst (dst_y1-dst_y0, pitch_d)."""
        L.check(self._lib.ofb_pyr_down_strip_device(self._h, src.data_ptr(), src.stride(0), sw, src.shape[0], src_y_off,
                                                    dst.data_ptr(), dst.stride(0), dst_y0, dst_y1, C.c_void_p(stream)))
        return dst

    def lk_level_device(self, prev, next, w: int, win: int, warp_mode: int = WARP_BILINEAR, flow_scale: float = 1.0,
                        cum_in=None, flow_out=None, cum_out=None, stream: int = 0):
        """One fused LK level on (n, h, pitch) uint8 tensors; cum_in is the (n, h>>1, w>>1, 2) cumulative
        flow of the coarser level or None for the coarsest level."""
        import torch

        n, h, pitch = prev.shape
        if flow_out is None:
            flow_out = torch.empty((n, h, w, 2), dtype=torch.float32, device=prev.device)
        L.check(self._lib.ofb_lk_level_device(self._h, prev.data_ptr(), next.data_ptr(), pitch, pitch * h, w, h, n, win,
                                              warp_mode, C.c_float(flow_scale),
                                              cum_in.data_ptr() if cum_in is not None else None, flow_out.data_ptr(),
                                              cum_out.data_ptr() if cum_out is not None else None, C.c_void_p(stream)))
        return flow_out

    def lk_level_strip_device(self, prev, next, w: int, y_off: int, h_global: int, out_y0: int, out_y1: int, win: int,
                              warp_mode: int, flow_scale: float, cum_in, cum_y_off: int, flow_out, cum_out=None,
                              overflow_flag=None, stream: int = 0):
        """Row-strip variant: prev/next are (h_local, pitch) uint8 holding global rows [y_off, y_off+h_local)."""
        h_local, pitch = prev.shape
        cum_h_local = cum_in.shape[0] if cum_in is not None else 0
        L.check(self._lib.ofb_lk_level_strip_device(
            self._h, prev.data_ptr(), next.data_ptr(), pitch, w, h_local, y_off, h_global, out_y0, out_y1, win, warp_mode,
            C.c_float(flow_scale), cum_in.data_ptr() if cum_in is not None else None, cum_y_off, cum_h_local,
            flow_out.data_ptr(), cum_out.data_ptr() if cum_out is not None else None,
            overflow_flag.data_ptr() if overflow_flag is not None else None, C.c_void_p(stream)))
        return flow_out


class FrameStream:
    """ofb_stream: push BGR frames one by one; every push after the first returns the flow against the
    previous frame, whose pyramid stays on the device."""

    def __init__(self, ctx: Context, w, h, levels, win, warp_mode, flow_scale, bil_win, bil_sigma_s, bil_sigma_b):
        self._ctx, self._lib = ctx, ctx._lib
        self.w, self.h, self.levels = w, h, levels
        p = OfbParams(w, h, levels, win, warp_mode, flow_scale, 1)
        hnd = C.c_void_p()
        L.check(self._lib.ofb_stream_create(ctx._h, C.byref(p), bil_win, float(bil_sigma_s), float(bil_sigma_b), C.byref(hnd)))
        self._h = hnd

    def push(self, frame_bgr: np.ndarray, want_total: bool = False):
        """Returns None for the first frame, else (flows, total) with flows[k] of shape (h>>k, w>>k, 2)."""
        if frame_bgr.shape != (self.h, self.w, 3):
            raise ValueError("frame must be (h, w, 3) uint8")
        flows = [np.empty((self.h >> k, self.w >> k, 2), np.float32) for k in range(self.levels)]
        total = np.empty((self.h, self.w, 2), np.float32) if want_total else None
        ptrs = (L.f32p * self.levels)(*[f.ctypes.data_as(L.f32p) for f in flows])
        has = C.c_int(0)
        L.check(self._lib.ofb_stream_push_bgr_host(self._h, _u8(frame_bgr).ctypes.data_as(L.u8p), ptrs,
                                                   total.ctypes.data_as(L.f32p) if want_total else None, C.byref(has)))
        return (flows, total) if has.value else None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ofb_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_flo(path: str, flow: np.ndarray) -> None:
    """Middlebury .flo file of an (h, w, 2) float32 flow field."""
    f = _f32(flow)
    L.check(L.load().ofb_write_flo(os.fsencode(path), f.ctypes.data_as(L.f32p), f.shape[1], f.shape[0]))


def planar_to_device(imgs: np.ndarray, device="cuda:0"):
    """(n, h, w) uint8 numpy -> torch (n, h, pitch) with pitch = align_up(w, 64); padding is zero."""
    import torch

    n, h, w = imgs.shape
    pitch = align_up(w, 64)
    t = torch.zeros((n, h, pitch), dtype=torch.uint8, device=device)
    t[:, :, :w] = torch.from_numpy(np.ascontiguousarray(imgs)).to(device)
    return t

"""Generate tests/golden/ref_gpu_b200.npz by running the UNMODIFIED reference's GPU functions on a B200.

TEST INFRASTRUCTURE ONLY.  Run on the GPU box (oracle/_ref/libofref.so travels there prebuilt; the
reference sources do not):

    gpurun -- 'python oracle/make_golden.py gpurun_out/ref_gpu_b200.npz'

then copy the file to tests/golden/.  Inputs are the seeded synthetic frames of orc_make_frame, so
only the reference's OUTPUTS are stored.  Sizes are launch-valid for the reference's swapped
<<<block, grid>>> launches: width and height multiples of 32 with
(w/32)*(h/32) <= 1024, at every level that is used.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

# (name, w, h, dx, dy, cell, seed)
SINGLE = [("a", 64, 64, 1.0, 0.5, 4, 1234), ("b", 96, 64, -0.75, 1.25, 8, 77), ("c", 128, 96, 2.0, -1.0, 2, 4242)]
WINDOWS = [5, 9, 15, 19]
MULTI = ("m", 256, 128, 3, 3.0, 1.5, 8, 99)  # name, w, h, levels, dx, dy, cell, seed


def main(out_path: str) -> None:
    g = {}
    for name, w, h, dx, dy, cell, seed in SINGLE:
        prev = O.make_frame(w, h, 0, 0, cell, seed)
        nxt = O.make_frame(w, h, dx, dy, cell, seed)
        pc3, nc3 = O.to_c3(prev), O.to_c3(nxt)
        ix = O.ref_gpu_conv(pc3, O.DX)
        iy = O.ref_gpu_conv(pc3, O.DY)
        it1 = O.ref_gpu_conv(pc3, O.DT)
        it2 = O.ref_gpu_conv(nc3, O.DT)
        it = O.ref_arr_sub(it2, it1)
        g[f"{name}_ix"], g[f"{name}_iy"], g[f"{name}_it1"], g[f"{name}_it2"] = ix, iy, it1, it2
        for win in WINDOWS:
            sums = [O.ref_gpu_srm(a, b, win, win) for a, b in ((ix, ix), (iy, iy), (ix, iy), (ix, it), (iy, it))]
            g[f"{name}_sums_w{win}"] = np.stack(sums)
            inv = O.ref_gpu_inverse(*sums)
            g[f"{name}_flow_w{win}"] = O.ref_gpu_lk_level_win(pc3, nc3, win)
            # the composed level is the same call sequence: store it once
            assert np.array_equal(inv.view(np.uint32), g[f"{name}_flow_w{win}"].view(np.uint32))
        # the entry point itself, single level (window hard-coded 19, OptFlowGpu.cu:1944-1945)
        flow = np.full((h, w, 2), -7.0, np.float32)
        pyr = (O._f32p * 1)(flow.ctypes.data_as(O._f32p))
        O.ref().ref_gpu_calc_opt_flow(pc3.ctypes.data_as(O._u8p), nc3.ctypes.data_as(O._u8p), w, h, pyr, 0, 1)
        g[f"{name}_entry_flow"] = flow
    # pyramid + the main.cu:256-262 loop through gpu::calc_opt_flow (warp as written, window 19)
    name, w, h, levels, dx, dy, cell, seed = MULTI
    prev = O.make_frame(w, h, 0, 0, cell, seed)
    nxt = O.make_frame(w, h, dx, dy, cell, seed)
    pp = O.ref_gpu_gauss_pyramid_c3(O.to_c3(prev), levels)
    pn = O.ref_gpu_gauss_pyramid_c3(O.to_c3(nxt), levels)
    for k in range(1, levels):
        g[f"{name}_pyr_prev_l{k}"] = pp[k][:, :, 0].copy()
        g[f"{name}_pyr_next_l{k}"] = pn[k][:, :, 0].copy()
        assert np.array_equal(pp[k][:, :, 0], pp[k][:, :, 1]) and np.array_equal(pp[k][:, :, 0], pp[k][:, :, 2])
    flows = [np.full((h >> k, w >> k, 2), -7.0, np.float32) for k in range(levels)]
    fptr = (O._f32p * levels)(*[f.ctypes.data_as(O._f32p) for f in flows])
    for k in range(levels - 1, -1, -1):
        O.ref().ref_gpu_calc_opt_flow(pp[k].ctypes.data_as(O._u8p), pn[k].ctypes.data_as(O._u8p), w >> k, h >> k, fptr, k,
                                      levels)
    for k in range(levels):
        g[f"{name}_loop_flow_l{k}"] = flows[k]
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **g)
    print(f"wrote {out_path}: {len(g)} arrays, {sum(v.nbytes for v in g.values())} bytes raw")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_gpu_b200.npz")

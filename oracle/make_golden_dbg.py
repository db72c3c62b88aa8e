"""Generate tests/golden/ref_gpu_dbg_b200.npz: the reference's u8 convolution (gpu::conv_3ch_1ch_tiled) and the
debug views of showTest (main.cu:19-92) composed from the reference's own functions, run on a B200.
TEST INFRASTRUCTURE ONLY.   gpurun -- 'python oracle/make_golden_dbg.py gpurun_out/ref_gpu_dbg_b200.npz'
Inputs are seeded synthetic frames (oracle.make_frame), so only outputs are stored."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

# (name, w, h, seed, level k)
CASES = [("a", 96, 64, 5, 0), ("b", 160, 120, 9, 1), ("c", 101, 77, 13, 2)]


def main(out_path):
    g = {}
    dtn = O.ref_mask_dt_n()
    g["mask_dt_n"] = dtn
    for name, w, h, seed, k in CASES:
        prev, cur = O.make_frame(w, h, 0, 0, 4, seed), O.make_frame(w, h, 1.5, -0.5, 4, seed)
        g[f"{name}_conv_dtn"] = O.ref_gpu_conv_u8(cur, dtn)
        g[f"{name}_conv_dx"] = O.ref_gpu_conv_u8(cur, np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], np.float32))
        for which in (O.VIEW_X, O.VIEW_Y, O.VIEW_T):
            g[f"{name}_view{which}"] = O.ref_debug_view(prev, cur, k, which)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **g)
    print(f"wrote {out_path}: {len(g)} arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_gpu_dbg_b200.npz")

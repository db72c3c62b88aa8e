"""Generate tests/golden/ref_gpu_pre_b200.npz: the reference's GPU grayscale and bilateral pre-filter
(gpu::grayscale_avg, gpu::bilinear_filter) run on a B200 at launch-valid sizes.
TEST INFRASTRUCTURE ONLY.   gpurun -- 'python oracle/make_golden_pre.py gpurun_out/ref_gpu_pre_b200.npz'
Inputs are seeded synthetic frames (oracle.make_bgr_frame), so only outputs are stored."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

# (name, w, h, seed, window, sigmaS, sigmaB)
CASES = [("p", 64, 64, 3, 9, 2.0, 10.0), ("q", 96, 64, 8, 5, 1.5, 25.0), ("r", 128, 96, 21, 9, 2.0, 10.0)]


def main(out_path):
    g = {}
    for name, w, h, seed, win, ss, sb in CASES:
        bgr = O.make_bgr_frame(w, h, 0, 0, 4, seed)
        gray = O.ref_gpu_grayscale(bgr)
        g[f"{name}_gray"] = gray[:, :, 0].copy()
        assert np.array_equal(gray[:, :, 0], gray[:, :, 1]) and np.array_equal(gray[:, :, 0], gray[:, :, 2])
        g[f"{name}_bil_gray"] = O.ref_gpu_bilateral(gray, gray, win, win, ss, sb)[:, :, 0].copy()  # main.cu:240 use
        g[f"{name}_bil_color"] = O.ref_gpu_bilateral(bgr, gray, win, win, ss, sb)                   # src != gray
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **g)
    print(f"wrote {out_path}: {len(g)} arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_gpu_pre_b200.npz")

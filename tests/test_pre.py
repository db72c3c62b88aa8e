"""grayscale, the bilateral pre-filter ("bilinear_filter") and the frame loop with
previous-pyramid reuse.  CPU tests pin the oracle to the reference (its CPU twins in-process, its GPU kernels
through tests/golden/ref_gpu_pre_b200.npz); GPU tests compare the CUDA path with both.

Tolerance of the bilateral filter: everything is double precision with the reference's operation order, so the
CUDA path equals the reference's GPU kernel byte for byte; the CPU oracle uses libm's pow instead of CUDA's,
which may differ in the last bit, so against the oracle at most 1 grey level on at most 1e-4 of the pixels."""
import os

import numpy as np
import pytest

from oracle.make_golden_pre import CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_gpu_pre_b200.npz")
HAVE_REF = os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libofref.so"))


def near(a, b, frac=1e-4):
    d = np.abs(a.astype(int) - b.astype(int))
    return d.max() <= 1 and (d != 0).mean() <= frac


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


# ------------------------------------------------------------------------------------------- CPU pins
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libofref.so not built")
def test_oracle_matches_reference_cpu_twins(oracle):
    bgr = oracle.make_bgr_frame(96, 64, 0, 0, 4, 11)
    gray = oracle.grayscale_c3(bgr)
    assert np.array_equal(gray, oracle.ref_cpu_grayscale(bgr))
    for k, s in ((9, 2.0), (5, 1.0), (3, 0.7)):
        assert np.array_equal(oracle.gaussian_kernel(s, k), oracle.ref_gaussian_kernel(s, k))
    assert np.array_equal(oracle.bilateral_c3(gray, gray, 9, 9, 2.0, 10.0), oracle.ref_cpu_bilateral(gray, gray, 9, 9, 2.0, 10.0))
    assert np.array_equal(oracle.bilateral_c3(bgr, gray, 5, 5, 1.5, 25.0), oracle.ref_cpu_bilateral(bgr, gray, 5, 5, 1.5, 25.0))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_gpu_golden(oracle, gold, case):
    name, w, h, seed, win, ss, sb = case
    bgr = oracle.make_bgr_frame(w, h, 0, 0, 4, seed)
    gray = oracle.grayscale_c3(bgr)
    assert np.array_equal(gray[:, :, 0], gold[f"{name}_gray"])
    assert near(oracle.bilateral_c3(gray, gray, win, win, ss, sb)[:, :, 0], gold[f"{name}_bil_gray"])
    assert near(oracle.bilateral_c3(bgr, gray, win, win, ss, sb), gold[f"{name}_bil_color"])


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_path_equals_reference_gpu_golden(ctx, oracle, gold, case):
    """Byte for byte against what the reference's own kernels produced on a B200."""
    import torch

    from cuda_optical_flow_2_b200 import planar_to_device

    name, w, h, seed, win, ss, sb = case
    bgr = oracle.make_bgr_frame(w, h, 0, 0, 4, seed)
    gray = ctx.grayscale_avg(bgr, h, w)
    assert np.array_equal(gray[:, :, 0], gold[f"{name}_gray"]) and np.array_equal(gray[:, :, 0], gray[:, :, 2])
    assert np.array_equal(ctx.bilinear_filter(gray, gray, w, h, win, win, ss, sb)[:, :, 0], gold[f"{name}_bil_gray"])
    assert np.array_equal(ctx.bilinear_filter(bgr, gray, w, h, win, win, ss, sb), gold[f"{name}_bil_color"])
    dev = planar_to_device(np.ascontiguousarray(gray[:, :, 0])[None])[0]
    out = ctx.bilateral_planar_device(dev, w, win, ss, sb)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy()[:, :w], gold[f"{name}_bil_gray"])


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,win", [(640, 480, 9), (333, 201, 5), (1920, 1080, 9)])
def test_bilateral_vs_oracle_any_size(ctx, oracle, w, h, win):
    bgr = oracle.make_bgr_frame(w, h, 0, 0, 8, 5)
    gray = oracle.grayscale_c3(bgr)
    assert np.array_equal(ctx.grayscale_avg(bgr, h, w), gray)
    got = ctx.bilinear_filter(gray, gray, w, h, win, win, 2.0, 10.0)
    assert near(got, oracle.bilateral_c3(gray, gray, win, win, 2.0, 10.0))


@pytest.mark.gpu
@pytest.mark.parametrize("bil", [0, 9])
def test_frame_stream_equals_pairwise_pipeline(ctx, oracle, bil):
    """main.cu:222-275 headless: every pushed frame is solved against the previous one, whose pyramid is
    reused; the result must equal running each consecutive pair from scratch."""
    w, h, levels, win = 320, 240, 3, 9
    frames = [oracle.make_bgr_frame(w, h, 1.5 * i, -0.75 * i, 8, 50) for i in range(4)]
    st = ctx.open_stream(w, h, levels, win, warp_mode=2, bil_win=bil, bil_sigma_s=2.0, bil_sigma_b=10.0)
    assert st.push(frames[0]) is None
    grays = []
    for f in frames:
        g = ctx.grayscale_avg(f, h, w)
        if bil:
            g = ctx.bilinear_filter(g, g, w, h, bil, bil, 2.0, 10.0)
        grays.append(np.ascontiguousarray(g[:, :, 0]))
    assert np.array_equal(ctx.grayscale_avg(frames[1], h, w), oracle.grayscale_c3(frames[1]))
    for i in range(1, 4):
        flows, total = st.push(frames[i], want_total=True)
        ref, cums = oracle.flow_pair(grays[i - 1], grays[i], levels, win, 2, oracle.SUMS_EXACT, 1.0, want_cum=True)
        for k in range(levels):
            m = ~np.isnan(ref[k])
            assert np.array_equal(np.isnan(flows[k]), np.isnan(ref[k])) and np.array_equal(flows[k][m], ref[k][m]), (i, k)
        m = ~np.isnan(cums[0])
        assert np.array_equal(total[m], cums[0][m])
    st.close()

"""flow composition, arrow sampling and .flo export (the headless part of
visualizeFlowField, main.cu:114-174) against the literal numpy restatement in oracle/oracle.py."""
import numpy as np
import pytest


def _pyramid(w, h, levels, seed=7):
    rng = np.random.default_rng(seed)
    fl = [(rng.standard_normal((h >> k, w >> k, 2)) * (3.0 + k)).astype(np.float32) for k in range(levels)]
    fl[0][3, 5] = np.nan  # the reference has no determinant threshold: NaN/inf travel through the composition
    fl[levels - 1][1, 1, 0] = np.inf
    return fl


def test_oracle_composition_is_the_coarse_to_fine_horner_form(oracle):
    """sum_k 2^(k-l) f_k accumulated as the reference does equals 2*cum_{k+1} + f_k applied coarse to fine in
    float (power-of-two scaling commutes with rounding): the form the fused kernel's cumulative output uses."""
    fl = _pyramid(96, 64, 4, seed=3)
    for level in range(4):
        cum = fl[3].copy()
        for k in range(2, level - 1, -1):
            up = np.repeat(np.repeat(cum, 2, axis=0), 2, axis=1)[: fl[k].shape[0], : fl[k].shape[1]]
            with np.errstate(invalid="ignore"):
                cum = (np.float32(2.0) * up + fl[k]).astype(np.float32)
        ref = oracle.compose_total(fl, level)
        assert np.array_equal(np.isnan(cum), np.isnan(ref))
        m = ~np.isnan(ref)
        assert np.array_equal(cum[m], ref[m])


def test_flo_roundtrip(tmp_path, oracle):
    from cuda_optical_flow_2_b200 import write_flo

    f = _pyramid(37, 21, 1)[0]
    path = str(tmp_path / "a.flo")
    write_flo(path, f)
    g = oracle.read_flo(path)
    assert g.shape == f.shape and np.array_equal(np.isnan(f), np.isnan(g))
    assert np.array_equal(f[~np.isnan(f)], g[~np.isnan(g)])


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,levels", [(640, 480, 4), (101, 77, 3), (64, 48, 1)])
def test_compose_flow_host_identical_to_reference_rule(ctx, oracle, w, h, levels):
    fl = _pyramid(w, h, levels)
    for level in range(levels):
        got = ctx.compose_flow(fl, w, h, levels, level)
        ref = oracle.compose_total(fl, level)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), f"level {level}"
        m = ~np.isnan(ref)
        assert np.array_equal(got[m], ref[m]), f"level {level}"


@pytest.mark.gpu
@pytest.mark.parametrize("level,arrow_res", [(0, 50), (1, 20), (2, 7)])
def test_flow_arrows_equal_reference_rule(ctx, oracle, level, arrow_res):
    w, h, levels = 640, 480, 3
    fl = _pyramid(w, h, levels, seed=11)
    got = ctx.flow_arrows(fl, w, h, levels, level, arrow_res)
    ref = oracle.flow_arrows(fl, level, arrow_res)
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.gpu
def test_flow_arrows_dense_grid(ctx, oracle):
    """arrow_res above half the width: the grid step is 1 and every pixel is a grid point (more arrows than
    arrow_res + 2 per row, which an earlier capacity guess assumed)."""
    w, h, levels = 100, 100, 1
    fl = _pyramid(w, h, levels, seed=5)
    got = ctx.flow_arrows(fl, w, h, levels, 0, 60)
    ref = oracle.flow_arrows(fl, 0, 60)
    assert got.shape == ref.shape and np.array_equal(got, ref) and got.shape[0] > 62 * 62


@pytest.mark.gpu
def test_compose_flow_equals_fused_total_flow(ctx, oracle):
    """The device path's total flow (cumulative output of the level-0 kernel) is the same composition."""
    import torch

    from cuda_optical_flow_2_b200 import WARP_BILINEAR, planar_to_device

    w, h, levels, win = 320, 240, 3, 9
    prev = oracle.make_frame(w, h, 0.0, 0.0, 4, 99)
    nxt = oracle.make_frame(w, h, 1.5, -0.75, 4, 99)
    total = torch.empty((1, h, w, 2), dtype=torch.float32, device="cuda:0")
    flows = ctx.flow_pairs_device(planar_to_device(prev[None]), planar_to_device(nxt[None]), w, levels, win,
                                  warp_mode=WARP_BILINEAR, total_flow=total)
    torch.cuda.synchronize()
    res = [np.ascontiguousarray(f[0].cpu().numpy()) for f in flows]
    got = ctx.compose_flow(res, w, h, levels, 0)
    ref = total[0].cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    m = ~np.isnan(ref)
    assert np.array_equal(got[m], ref[m])

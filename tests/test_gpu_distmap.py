"""Coarse distance map and activation candidate filter on the device against the oracle (CoarseTracker.cpp:1216-1366,
FullSystem.cpp:838-901): the field and every verdict are integer work and must match exactly."""
import numpy as np
import pytest
import oracle_py as O
import oracle_distmap_py as D
from conftest import load_pkg

pytestmark = pytest.mark.gpu


def _pair(w, h):
    pkg = load_pkg()
    K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
    return pkg.Context(w, h, K, 0.1), O.Oracle(w, h, K, 0.1)


@pytest.mark.parametrize("w,h,n_pts", [(128, 96, 60), (640, 480, 1500), (640, 480, 12), (1232, 368, 4000), (640, 480, 0)])
def test_make_distance_map(w, h, n_pts):
    ctx, orc = _pair(w, h)
    dm = D.DistMap(orc)
    inp = D.make_inputs(orc, 1, n_pts=n_pts)
    mo = dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    mg = ctx.distmap_make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    assert np.array_equal(mo, mg)
    ctx.close()


def test_add_into_dist_final():
    ctx, orc = _pair(320, 240)
    dm = D.DistMap(orc)
    inp = D.make_inputs(orc, 2, n_pts=40)
    dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    ctx.distmap_make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    rng = np.random.default_rng(0)
    for _ in range(5):   # one cell at a time: identical including the border cells
        uv = [[int(rng.integers(1, dm.w1)), int(rng.integers(1, dm.h1))]]
        assert np.array_equal(dm.add(uv), ctx.distmap_add(uv))
    uv = np.stack([rng.integers(1, dm.w1, 50), rng.integers(1, dm.h1, 50)], 1)
    mo, mg = dm.add(uv), ctx.distmap_add(uv)
    assert np.array_equal(mo[1:-1, 1:-1], mg[1:-1, 1:-1])   # interior: independent of the insertion order
    ctx.close()


@pytest.mark.parametrize("w,h,n_pts,n_cand", [(128, 96, 100, 3000), (640, 480, 1500, 8000), (640, 480, 50, 20000), (1232, 368, 3000, 12000)])
@pytest.mark.parametrize("mad", [0.0, 0.3, 1.0, 2.0, 4.0])
def test_activation_filter(w, h, n_pts, n_cand, mad):
    ctx, orc = _pair(w, h)
    dm = D.DistMap(orc)
    inp = D.make_inputs(orc, 7, n_pts=n_pts, n_cand=n_cand)
    dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    ctx.distmap_make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    vo, mo = dm.filter(inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
    vg, mg, rounds = ctx.activation_filter(inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
    assert np.array_equal(vo, vg), (int((vo != vg).sum()), np.nonzero(vo != vg)[0][:10], rounds)
    assert (vg == 1).sum() > 0
    assert np.array_equal(mo[1:-1, 1:-1], mg[1:-1, 1:-1])
    ctx.close()


def test_filter_on_traced_points():
    """the real producer chain: selected pixels -> immature points -> stereo trace -> candidate filter"""
    import synth
    pkg = load_pkg()
    w, h = 640, 480
    K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
    ctx, orc = pkg.Context(w, h, K, synth.BASELINE), O.Oracle(w, h, K, synth.BASELINE)
    img, _ = synth.render(synth.make_scene(), synth.camera_pose(0), w, h, K)
    g = ctx.frame_create(); ctx.make_images(g, img)
    _, n = ctx.make_maps(g, 2000.0, want_map=False)
    uv, ty = ctx.selector_points()
    pts, ok = ctx.immature_init(g, uv)
    pts, ty = pts[ok], ty[ok]
    rng = np.random.default_rng(0)
    pts["idepth_min"] = rng.uniform(0.1, 0.5, len(pts)); pts["idepth_max"] = pts["idepth_min"] + 0.05
    pts["lastTraceStatus"] = 0; pts["lastTracePixelInterval"] = 1.0; pts["quality"] = 5.0
    inp = D.make_inputs(orc, 11, n_hosts=1, n_pts=300)
    host = np.zeros(len(pts), np.int32)
    dm = D.DistMap(orc)
    dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    ctx.distmap_make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    vo, _ = dm.filter(inp["KRKi"], inp["Kt"], [0], host, pts, ty, 2.0)
    vg, _, rounds = ctx.activation_filter(inp["KRKi"], inp["Kt"], [0], host, pts, ty, 2.0)
    assert np.array_equal(vo, vg)
    assert 0 < (vg == 1).sum() < len(pts)
    ctx.close()

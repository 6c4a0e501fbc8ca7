"""Pixel selector on the device against the oracle — bit-exact maps, counts, thresholds, potential adaptation
(PixelSelector2.cpp:84-536; call site FullSystem.cpp:1599-1625)."""
import numpy as np
import pytest
import oracle_py as O
import oracle_select_py as S
import synth
from conftest import load_pkg

pytestmark = pytest.mark.gpu


def _pair(w, h, seed, kind):
    pkg = load_pkg()
    K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
    rng = np.random.default_rng(seed)
    if kind == "scene":
        img, _ = synth.render(synth.make_scene(), synth.camera_pose(seed), w, h, K)
    elif kind == "u8":
        img = np.kron(rng.integers(0, 255, (h // 4, w // 4)), np.ones((4, 4))).astype(np.float32)
        img += rng.integers(0, 3, (h, w)).astype(np.float32)
    elif kind == "flat":
        img = (rng.random((h, w)) * 6 + 100).astype(np.float32)
        img[h // 3:h // 2, w // 4:w // 2] += 60
    elif kind == "noise":
        img = (rng.random((h, w)) * 255).astype(np.float32)
    ctx = pkg.Context(w, h, K, 0.1)
    orc = O.Oracle(w, h, K, 0.1)
    g, o = ctx.frame_create(), orc.frame_new()
    ctx.make_images(g, img)
    orc.make_images(o, img)
    return ctx, orc, g, o


def test_random_pattern():
    ctx, orc, g, o = _pair(160, 128, 0, "scene")
    assert np.array_equal(ctx.selector_random_pattern(), S.Selector(orc).random_pattern())
    ctx.close()


@pytest.mark.parametrize("kind,w,h", [("scene", 640, 480), ("u8", 640, 480), ("flat", 320, 240), ("noise", 168, 136), ("scene", 1240, 376)])
def test_make_hists(kind, w, h):
    ctx, orc, g, o = _pair(w, h, 1, kind)
    ths_o, sm_o = S.Selector(orc).make_hists(o)
    ths_g, sm_g = ctx.selector_make_hists(g)
    assert np.array_equal(ths_o, ths_g)
    assert np.array_equal(sm_o, sm_g)
    ctx.close()


@pytest.mark.parametrize("kind,w,h", [("scene", 640, 480), ("u8", 640, 480), ("flat", 320, 240), ("noise", 168, 136), ("u8", 168, 136),
                                      ("scene", 1240, 376)])
@pytest.mark.parametrize("pot", [1, 2, 3, 4, 7])
def test_select_bit_exact(kind, w, h, pot):
    ctx, orc, g, o = _pair(w, h, 2, kind)
    sel = S.Selector(orc)
    sel.make_hists(o)
    ctx.selector_make_hists(g)
    for thf in (1.0, 2.0):
        mo, no = sel.select(o, pot, thf)
        mg, ng = ctx.selector_select(g, pot, thf)
        assert np.array_equal(no, ng), (no, ng)
        assert np.array_equal(mo, mg)
        uv, ty = ctx.selector_points()
        ys, xs = np.nonzero(mo)
        assert np.array_equal(uv[:, 0], xs.astype(np.float32)) and np.array_equal(uv[:, 1], ys.astype(np.float32))
        assert np.array_equal(ty, mo[ys, xs])
    ctx.close()


@pytest.mark.parametrize("kind,w,h", [("scene", 640, 480), ("u8", 640, 480), ("flat", 640, 480), ("noise", 320, 240)])
def test_make_maps_sequence(kind, w, h):
    """the call pattern of makeNewTraces over consecutive keyframes: currentPotential carries over"""
    pkg = load_pkg()
    ctx, orc, g, o = _pair(w, h, 3, kind)
    sel = S.Selector(orc)
    K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
    for it, density in enumerate([3000.0, 3000.0, 600.0, 1500.0, 12000.0, 50.0]):
        if kind == "scene":
            img, _ = synth.render(synth.make_scene(), synth.camera_pose(it), w, h, K)
            ctx.make_images(g, img)
            orc.make_images(o, img)
            sel.forget_hist()   # a new FrameHessian in the reference; the device keys on the slot's generation
        mo, no = sel.make_maps(o, density)
        mg, ng = ctx.make_maps(g, density)
        assert no == ng, (it, no, ng)
        assert np.array_equal(mo, mg)
        assert sel.potential() == ctx.selector_potential()
        uv, ty = ctx.selector_points()
        assert len(ty) == ng
        ys, xs = np.nonzero(mo)
        assert np.array_equal(uv[:, 0], xs.astype(np.float32)) and np.array_equal(uv[:, 1], ys.astype(np.float32))
    ctx.close()


def test_selected_points_feed_immature_init():
    """makeNewTraces: every selected pixel becomes an ImmaturePoint of its host frame (FullSystem.cpp:1609-1621)"""
    import oracle_trace_py as T
    ctx, orc, g, o = _pair(640, 480, 4, "scene")
    sel = S.Selector(orc)
    mo, no = sel.make_maps(o, 2000.0)
    mg, ng = ctx.make_maps(g, 2000.0, want_map=False)
    uv, ty = ctx.selector_points()
    assert ng == no and len(uv) == no
    pg, okg = ctx.immature_init(g, uv)
    po, oko = T.immature_init(orc, o, uv)
    assert np.array_equal(okg, oko)
    assert np.array_equal(pg["color"], po["color"]) and np.array_equal(pg["energyTH"][okg], po["energyTH"][oko])
    ctx.close()

"""A1 parity on the GPU: FrameHessian::makeImages (HessianBlocks.cpp:141-203) and the bilinear gathers
(globalFuncs.h:73-86,160-184) through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest
import oracle_py as O
import synth

pytestmark = pytest.mark.gpu


def _compare_pyramids(ctx, orc, fg, fo):
    for lvl in range(orc.levels):
        dI_g, ag_g = ctx.frame_download(fg, lvl)
        dI_o, ag_o = orc.frame_get(fo, lvl)
        # intensity: all rows; gradients/absSquaredGrad: rows 1..h-2 (rows 0,h-1 are uninitialised in the reference)
        assert np.array_equal(dI_g[..., 0], dI_o[..., 0], equal_nan=True), f"I level {lvl}"
        assert np.array_equal(dI_g[1:-1, :, 1:], dI_o[1:-1, :, 1:]), f"gradient level {lvl}"
        assert np.array_equal(ag_g[1:-1], ag_o[1:-1]), f"absSquaredGrad level {lvl}"


def test_make_images_bit_exact_kitti_shape(pkg, frames):
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE)
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    assert ctx.levels == orc.levels == 5
    for k in (0, 1):
        fg, fo = ctx.frame_create(), orc.frame_new()
        ctx.make_images(fg, frames[k][0])
        orc.make_images(fo, frames[k][0])
        _compare_pyramids(ctx, orc, fg, fo)
    ctx.close()


@pytest.mark.parametrize("shape", [(1920, 1088), (640, 480), (96, 64), (1241, 376)])
def test_make_images_bit_exact_other_shapes(pkg, shape):
    w, h = shape
    rng = np.random.default_rng(w * 7 + h)
    img = rng.uniform(0, 255, (h, w)).astype(np.float32)
    K = (500.0, 500.0, w / 2 - 0.5, h / 2 - 0.5)
    ctx = pkg.Context(w, h, K)
    orc = O.Oracle(w, h, K)
    assert ctx.levels == orc.levels
    for l in range(orc.levels):
        assert ctx.level_size(l) == orc.level_size(l)
        Kg, Kig = ctx.level_K(l)
        Ko, Kio = orc.level_K(l)
        assert np.array_equal(Kg, Ko) and np.array_equal(Kig, Kio)
    fg, fo = ctx.frame_create(), orc.frame_new()
    ctx.make_images(fg, img)
    orc.make_images(fo, img)
    _compare_pyramids(ctx, orc, fg, fo)
    ctx.close()


def test_make_images_nonfinite_pixels(pkg):
    """Non-finite gradients are zeroed (HessianBlocks.cpp:186-187); NaN intensities propagate into the pyramid."""
    w, h = 128, 96
    rng = np.random.default_rng(11)
    img = rng.uniform(0, 255, (h, w)).astype(np.float32)
    img[10, 17] = np.nan
    img[40, 0] = np.inf
    img[41, w - 1] = -np.inf
    K = (100.0, 100.0, 63.5, 47.5)
    ctx, orc = pkg.Context(w, h, K), O.Oracle(w, h, K)
    fg, fo = ctx.frame_create(), orc.frame_new()
    ctx.make_images(fg, img)
    orc.make_images(fo, img)
    _compare_pyramids(ctx, orc, fg, fo)
    dI, ag = ctx.frame_download(fg, 0)
    assert dI[10, 16, 1] == 0 and dI[10, 18, 1] == 0 and np.isfinite(ag[1:-1]).all()
    ctx.close()


def test_frame_slots_are_reused(pkg, frames):
    ctx = pkg.Context(synth.W, synth.H, synth.K4)
    a = ctx.frame_create()
    ctx.make_images(a, frames[0][0])
    ctx.frame_release(a)
    b = ctx.frame_create()
    assert a == b
    with pytest.raises(pkg.SdsoError):
        ctx.frame_download(b, 0)  # released + recreated: not valid until makeImages runs again
    ctx.close()


def test_interp33_bit_exact(pkg, frames):
    ctx = pkg.Context(synth.W, synth.H, synth.K4)
    orc = O.Oracle(synth.W, synth.H, synth.K4)
    fg, fo = ctx.frame_create(), orc.frame_new()
    ctx.make_images(fg, frames[0][0])
    orc.make_images(fo, frames[0][0])
    rng = np.random.default_rng(2)
    for lvl in (0, 2, 4):
        w, h = orc.level_size(lvl)
        n = 4000
        xy = np.stack([rng.uniform(1.2, w - 3.2, n), rng.uniform(1.2, h - 3.2, n)], 1).astype(np.float32)
        xy[:50] = np.floor(xy[:50])          # integer coordinates
        xy[50:100] = np.floor(xy[50:100]) + 0.5  # half-pixel
        for bilin in (False, True):
            g = ctx.interp33(fg, lvl, xy, bilin)
            o = orc.interp33(fo, lvl, xy, bilin)
            assert np.array_equal(g, o), (lvl, bilin)
    ctx.close()

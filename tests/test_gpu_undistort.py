"""Undistort::undistort (photometric un-mapping + bilinear remap) on the device against the oracle, bit-exact, and the fused
raw -> rectified -> makeImages path (Undistort.cpp:222-260, 398-489; HessianBlocks.cpp:141)."""
import numpy as np
import pytest
import oracle_py as O
import oracle_undistort_py as U
from conftest import load_pkg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w_org,h_org,w,h", [(1241, 376, 1232, 368), (752, 480, 640, 480)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_undistort_bit_exact(w_org, h_org, w, h, mode):
    pkg = load_pkg()
    rng = np.random.default_rng(mode)
    K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
    ctx, orc = pkg.Context(w, h, K, 0.1), O.Oracle(w, h, K, 0.1)
    raw = rng.integers(0, 256, (h_org, w_org), dtype=np.uint8)
    rx, ry = U.radial_remap(w, h, w_org, h_org)
    G = (np.linspace(0, 255, 256) ** 1.1 / 255 ** 0.1).astype(np.float32) if mode else None
    vig = rng.uniform(1.0, 1.6, (h_org, w_org)).astype(np.float32) if mode == 2 else None
    pc = mode if mode else 0
    ctx.undistort_setup(w_org, h_org, rx, ry, G, vig, photometric_calibration=pc)
    g = ctx.frame_create()
    out_g, e_g = ctx.undistort(raw, exposure=0.013, frame=g)
    out_o, e_o = U.undistort(raw, w, h, rx, ry, G, vig, photometric_calibration=pc, exposure=0.013)
    assert np.array_equal(out_g, out_o)
    assert e_g == e_o
    # the frame built on the device from the rectified image equals makeImages of the oracle's rectified image
    o = orc.frame_new()
    orc.make_images(o, out_o)
    for lvl in range(orc.levels):
        a, aa = ctx.frame_download(g, lvl)
        b, bb = orc.frame_get(o, lvl)
        assert np.array_equal(a[1:-1], b[1:-1]) and np.array_equal(aa[1:-1], bb[1:-1])
    # exposure <= 0 falls back to the plain factor (processFrame :231)
    out_g2, e2 = ctx.undistort(raw, exposure=0.0, factor=0.5)
    out_o2, _ = U.undistort(raw, w, h, rx, ry, G, vig, photometric_calibration=pc, exposure=0.0, factor=0.5)
    assert np.array_equal(out_g2, out_o2)
    ctx.close()


def test_undistort_rejects_bad_tables():
    pkg = load_pkg()
    ctx = pkg.Context(640, 480, (380.0, 380.0, 319.5, 239.5), 0.1)
    rx = np.full((480, 640), 751.5, np.float32)   # the 2x2 taps would leave a 752-wide raw image
    ry = np.full((480, 640), 10.0, np.float32)
    with pytest.raises(Exception):
        ctx.undistort_setup(752, 480, rx, ry)
    with pytest.raises(Exception):
        ctx.undistort(np.zeros((480, 752), np.uint8))
    ctx.close()

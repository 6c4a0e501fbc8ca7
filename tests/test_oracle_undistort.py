"""Input preparation / trajectory rows — CPU checks: the product's host-only row formatter against the C++ stream format,
and the oracle's undistortion against a numpy restatement."""
import numpy as np
import oracle_undistort_py as U
from conftest import load_pkg


def test_trajectory_row_format():
    pkg = load_pkg()
    rng = np.random.default_rng(0)
    cases = [np.eye(4)[:3].reshape(12), rng.normal(0, 1, 12), rng.normal(0, 1e-7, 12), rng.normal(0, 1e6, 12),
             np.array([1, 0, 0, 1e-5, 0, 1, 0, 123456789.123456789, 0, 0, 1, -0.1])]
    for T in cases:
        a, b = pkg.trajectory_row(T), U.trajectory_row(T)
        assert a == b, (a, b)
        assert len(a.split()) == 12 and a.endswith("\n")
        assert np.allclose(np.array(a.split(), float), T, rtol=1e-14, atol=0)


def test_oracle_undistort_against_numpy():
    rng = np.random.default_rng(1)
    w_org, h_org, w, h = 200, 120, 160, 96
    raw = rng.integers(0, 256, (h_org, w_org), dtype=np.uint8)
    rx, ry = U.radial_remap(w, h, w_org, h_org)
    assert (rx < 0).any() and (rx >= 0).any()
    G = (np.linspace(0, 255, 256) ** 1.1 / 255 ** 0.1).astype(np.float32)
    vig = rng.uniform(1.0, 1.6, (h_org, w_org)).astype(np.float32)
    for mode, g, v in ((0, None, None), (1, G, None), (2, G, vig)):
        out, e = U.undistort(raw, w, h, rx, ry, g, v, photometric_calibration=max(mode, 1) if g is not None else 0, exposure=0.02)
        f = np.float32
        img = raw.astype(f) if g is None else G[raw]
        if mode == 2:
            img = img * vig
        ok = rx >= 0
        xi, yi = rx.astype(np.int32), ry.astype(np.int32)
        xi[~ok] = 0; yi[~ok] = 0
        dx, dy = (rx - xi).astype(f), (ry - yi).astype(f)
        dxdy = dx * dy
        ref = dxdy * img[yi + 1, xi + 1] + (dy - dxdy) * img[yi + 1, xi] + (dx - dxdy) * img[yi, xi + 1] + (f(1) - dx - dy + dxdy) * img[yi, xi]
        ref[~ok] = 0
        assert np.array_equal(out, ref.astype(f)), mode
        assert e == np.float32(0.02)

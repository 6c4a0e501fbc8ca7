"""Committed golden fixtures of the rows either side of the residual path (tests/golden/frontend_v1.npz, made by
tests/golden/make_golden_frontend.py from the oracle — see its docstring for provenance): the oracle must keep reproducing
them (CPU) and the device path must match them (GPU)."""
import os
import sys

import numpy as np
import pytest
import oracle_py as O
import oracle_select_py as S
import oracle_distmap_py as D
import oracle_undistort_py as U
import oracle_trace_py as T

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_frontend as MG   # noqa: E402  (deterministic input formulas only)

G = np.load(os.path.join(HERE, "golden", "frontend_v1.npz"))
H, W = G["image"].shape
K4 = tuple(float(x) for x in G["K4"])
DENS = [float(d) for d in G["maps_density"]]


def _inputs():
    return dict(KRKi=G["dm_KRKi"], Kt=G["dm_Kt"], pt_host=G["dm_pt_host"], pt_uvid=G["dm_pt_uvid"], cand_host=G["dm_cand_host"],
                my_type=G["dm_my_type"], flagged=G["dm_flagged"], pts=G["dm_pts"].view(T.DTYPE))


def _undistort_inputs():
    raw = G["und_raw"]
    h_org, w_org = raw.shape
    rx, ry = U.radial_remap(W, H, w_org, h_org)
    assert np.array_equal(rx[::4, ::4], G["und_rx"]) and np.array_equal(ry[::4, ::4], G["und_ry"])
    return raw, rx, ry, G["und_G"], MG.vignette(w_org, h_org)


def test_oracle_reproduces_golden():
    orc = O.Oracle(W, H, K4, 0.5)
    f = orc.frame_new()
    orc.make_images(f, G["image"].astype(np.float32))
    sel = S.Selector(orc)
    ths, sm = sel.make_hists(f)
    assert np.array_equal(ths, G["ths"]) and np.array_equal(sm, G["ths_smoothed"])
    for pot in (1, 3, 5):
        m, n = sel.select(f, pot, 1.0)
        assert np.array_equal(m.astype(np.uint8), G[f"select_pot{pot}"]) and np.array_equal(n, G[f"select_n_pot{pot}"])
    sel.potential(3)
    for i, d in enumerate(DENS):
        m, num = sel.make_maps(f, d)
        assert np.array_equal(m.astype(np.uint8), G[f"maps_{i}"]) and num == int(G["maps_num"][i]) and sel.potential() == int(G["maps_potential"][i])
    inp = _inputs()
    dm = D.DistMap(orc)
    assert np.array_equal(dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"]).astype(np.int16), G["dm_map"])
    for mad in (0.5, 2.0):
        dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
        v, m = dm.filter(inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
        assert np.array_equal(v.astype(np.int8), G[f"dm_verdict_{mad}"]) and np.array_equal(m.astype(np.int16), G[f"dm_after_{mad}"])
    raw, rx, ry, g, vig = _undistort_inputs()
    und, e = U.undistort(raw, W, H, rx, ry, g, vig, photometric_calibration=2, exposure=0.02)
    assert np.array_equal(und[::3, ::5], G["und_out_sub"]) and np.isclose(und.astype(np.float64).sum(), float(G["und_out_sum"]), rtol=1e-12)


@pytest.mark.gpu
def test_device_matches_golden(pkg):
    ctx = pkg.Context(W, H, K4, 0.5)
    f = ctx.frame_create()
    ctx.make_images(f, G["image"].astype(np.float32))
    ths, sm = ctx.selector_make_hists(f)
    assert np.array_equal(ths, G["ths"]) and np.array_equal(sm, G["ths_smoothed"])
    for pot in (1, 3, 5):
        m, n = ctx.selector_select(f, pot, 1.0)
        assert np.array_equal(m.astype(np.uint8), G[f"select_pot{pot}"]) and np.array_equal(n, G[f"select_n_pot{pot}"])
    ctx.selector_potential(3)
    for i, d in enumerate(DENS):
        m, num = ctx.make_maps(f, d)
        assert np.array_equal(m.astype(np.uint8), G[f"maps_{i}"]) and num == int(G["maps_num"][i]) and ctx.selector_potential() == int(G["maps_potential"][i])
    inp = _inputs()
    assert np.array_equal(ctx.distmap_make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"]).astype(np.int16), G["dm_map"])
    for mad in (0.5, 2.0):
        ctx.distmap_make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"], want_map=False)
        v, m, _ = ctx.activation_filter(inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
        assert np.array_equal(v.astype(np.int8), G[f"dm_verdict_{mad}"])
        assert np.array_equal(m.astype(np.int16)[1:-1, 1:-1], G[f"dm_after_{mad}"][1:-1, 1:-1])
    raw, rx, ry, g, vig = _undistort_inputs()
    ctx.undistort_setup(raw.shape[1], raw.shape[0], rx, ry, g, vig, photometric_calibration=2)
    und, e = ctx.undistort(raw, exposure=0.02)
    assert np.array_equal(und[::3, ::5], G["und_out_sub"]) and np.isclose(und.astype(np.float64).sum(), float(G["und_out_sum"]), rtol=1e-12)
    assert e == float(G["und_exposure"])
    ctx.close()

"""Committed golden fixtures (tests/golden/hotpath_v1.npz, made by tests/golden/make_golden.py from the oracle — see its
docstring for provenance): the oracle must keep reproducing them (CPU), and the device path must match them (GPU)."""
import os

import numpy as np
import pytest
import oracle_py as O
import oracle_trace_py as OT

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.npz"))
W, H = G["left"].shape[1], G["left"].shape[0]
K4 = tuple(float(x) for x in G["K4"])
BASELINE = float(G["baseline"])
K33 = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)


def check(api, is_device):
    """api: an object with the shared method names of oracle_py.Oracle / sdso_b200.Context."""
    left, right = G["left"].astype(np.float32), G["right"].astype(np.float32)
    fl, fr = (api.frame_create(), api.frame_create()) if is_device else (api.frame_new(), api.frame_new())
    api.make_images(fl, left); api.make_images(fr, right)
    assert api.levels == int(G["levels"])
    get = api.frame_download if is_device else api.frame_get
    for lvl in range(api.levels):
        dI, ag = get(fl, lvl)
        assert np.array_equal(dI[1:-1:5, ::7], G[f"dI_l{lvl}_sub"]), lvl
        assert np.array_equal(ag[1:-1:5, ::7], G[f"ag_l{lvl}_sub"]), lvl
        assert np.allclose(dI[1:-1].astype(np.float64).sum(axis=(0, 1)), G[f"dI_l{lvl}_sum"], rtol=1e-12)
    assert np.array_equal(api.interp33(fl, 0, G["interp_xy"]), G["interp33"])
    assert np.array_equal(api.interp33(fl, 0, G["interp_xy"], bilin=True), G["interp33bilin"])
    dt = OT.DTYPE
    pts, ok = (api.immature_init(fl, G["imm_uv"]) if is_device else OT.immature_init(api, fl, G["imm_uv"]))
    assert np.array_equal(ok, G["imm_ok"])
    assert pts.tobytes() == G["imm_init"].view(dt).tobytes() or np.array_equal(pts["color"], G["imm_init"].view(dt)["color"])
    st = api.trace_stereo(fr, K33, True, pts) if is_device else OT.trace_stereo(api, fr, K33, True, pts)
    gold = G["stereo_pts"].view(dt)
    assert np.array_equal(st, G["stereo_status"])
    assert np.array_equal(pts["bestIdx"], gold["bestIdx"]) and np.array_equal(pts["numSteps"], gold["numSteps"])
    for f in ("idepth_stereo", "idepth_min_stereo", "idepth_max_stereo", "lastTraceUV", "quality"):
        assert np.allclose(pts[f], gold[f], rtol=1e-4, atol=1e-6, equal_nan=True), f
    # the pair is fronto-parallel with a 6 px disparity: the search must have found it
    assert np.median(np.abs((gold["u_stereo"] - gold["lastTraceUV"][:, 0]) - 6.0)) < 0.3
    api.tracker_set_ref(fl, G["splats"], (0.0, 0.0))
    for lvl in range(api.levels):
        pc = np.stack(api.tracker_get_pc(lvl), 1)
        assert pc.shape[0] == int(G[f"pc_n_l{lvl}"]) and np.array_equal(pc, G[f"pc_l{lvl}"]), lvl
    r = api.calc_res_gs(fr, 0, G["gs_T"], (0.0, 0.0), 20.0)
    assert r["warped_n"] == int(G["gs_warped_n"]) and np.array_equal(r["warped"], G["gs_warped"])
    assert np.allclose(r["rs"], G["gs_rs"], rtol=1e-4) and np.allclose(r["H"], G["gs_H"], rtol=1e-4, atol=1e-4 * np.abs(G["gs_H"]).max())
    assert np.allclose(r["b"], G["gs_b"], rtol=1e-4, atol=1e-4 * np.abs(G["gs_b"]).max())
    tr = api.track(fr, np.eye(4)[:3], (0.0, 0.0), api.levels - 1, [np.nan] * 5, 0)
    assert bool(tr["ok"]) == bool(G["track_ok"])
    assert np.abs(tr["T"] - G["track_T"]).max() < 1e-4
    assert np.allclose(tr["lastResiduals"], G["track_lastRes"], rtol=1e-3, atol=5e-4, equal_nan=True)  # RMS of near-zero residuals (grey levels)
    assert abs(tr["T"][0, 3] + BASELINE) < 2e-3   # the motion between the two views is the baseline


def test_oracle_reproduces_golden():
    check(O.Oracle(W, H, K4, BASELINE), False)


@pytest.mark.gpu
def test_device_matches_golden(pkg):
    ctx = pkg.Context(W, H, K4, BASELINE)
    check(ctx, True)
    ctx.close()

"""CPU tests of the restated epipolar search (oracle/trace.cpp): static stereo recovers the true inverse depth of the
synthetic scene, temporal tracing narrows the interval around it, and every ImmaturePointStatus branch is reachable."""
import numpy as np
import pytest
import oracle_py as O
import oracle_trace_py as OT
import synth
import trace_synth as TS


@pytest.fixture(scope="module")
def setup(frames):
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    fl, fr, f1 = orc.frame_new(), orc.frame_new(), orc.frame_new()
    orc.make_images(fl, frames[0][0]); orc.make_images(fr, frames["r0"][0]); orc.make_images(f1, frames[1][0])
    rng = np.random.default_rng(2)
    uv = TS.candidate_pixels(frames[0][0], 600, rng)
    pts, ok = OT.immature_init(orc, fl, uv)
    assert ok.all()
    return orc, (fl, fr, f1), uv, pts


def test_constructor_fields(setup, frames):
    orc, fids, uv, pts = setup
    assert np.all(pts["energyTH"] == 8 * 144) and np.all(pts["quality"] == 10000)
    assert np.all((pts["weights"] > 0) & (pts["weights"] <= 1))
    # integer positions: the BiLin colour is the pixel value itself
    img = frames[0][0]
    assert np.array_equal(pts["color"][:, 4], img[uv[:, 1].astype(int), uv[:, 0].astype(int)])
    assert np.all(pts["lastTraceStatus"] == 5) and np.all(np.isnan(pts["idepth_max"]))


def test_static_stereo_recovers_depth(setup, frames):
    orc, (fl, fr, f1), uv, pts0 = setup
    pts = pts0.copy()
    st = OT.trace_stereo(orc, fr, TS.K33(), True, pts)
    good = st == 0
    assert good.mean() > 0.5
    true_id = 1.0 / frames[0][1][uv[:, 1].astype(int), uv[:, 0].astype(int)]
    err = np.abs(pts["idepth_stereo"][good] - true_id[good]) / true_id[good]
    assert np.median(err) < 0.02
    inside = (pts["idepth_min_stereo"][good] <= true_id[good] * 1.02) & (pts["idepth_max_stereo"][good] >= true_id[good] * 0.98)
    assert inside.mean() > 0.8
    assert np.all(pts["numSteps"][good] == 45)  # 1.9999 + (1232+368)*0.027 = 45.19 -> 45 steps for an open interval


def test_temporal_trace_narrows_interval(setup, frames):
    orc, (fl, fr, f1), uv, pts0 = setup
    pts = pts0.copy()
    st = OT.trace_stereo(orc, fr, TS.K33(), True, pts)
    good = st == 0
    pts["idepth_min"] = np.where(good, pts["idepth_min_stereo"], pts["idepth_min"])
    pts["idepth_max"] = np.where(good, pts["idepth_max_stereo"], pts["idepth_max"])
    KRKi, Kt = TS.krki_kt(synth.camera_pose(0), synth.camera_pose(1))
    before = (pts["idepth_max"] - pts["idepth_min"]).copy()
    st2 = OT.trace_on(orc, f1, KRKi, Kt, (1.0, 0.0), pts)
    assert set(np.unique(st2)) <= {0, 1, 2, 3, 4}
    g = (st2 == 0) & good
    assert g.sum() > 50
    true_id = 1.0 / frames[0][1][uv[:, 1].astype(int), uv[:, 0].astype(int)]
    mid = 0.5 * (pts["idepth_min"][g] + pts["idepth_max"][g])
    assert np.median(np.abs(mid - true_id[g]) / true_id[g]) < 0.05
    assert np.all(pts["lastTracePixelInterval"][g] > 0)


def test_every_status_branch_is_hit(setup):
    orc, (fl, fr, f1), uv, pts0 = setup
    rng = np.random.default_rng(0)
    pts = TS.adversarial(pts0[:64].copy(), rng)
    KRKi, Kt = TS.krki_kt(synth.camera_pose(0), synth.camera_pose(1))
    pts["idepth_min"][:64] = 0.02; pts["idepth_max"][:64] = 0.2
    st = OT.trace_on(orc, f1, KRKi, Kt, (1.0, 0.0), pts)
    seen = set(np.unique(st))
    st_again = OT.trace_on(orc, f1, KRKi, Kt, (1.0, 0.0), pts)
    assert {0, 1, 2, 3, 4} <= seen | set(np.unique(st_again)), seen
    # an OUTLIER traced again as an outlier becomes OOB (ImmaturePoint.cpp:788-791)
    was_out = st == 2
    assert np.all(np.isin(st_again[was_out], (1, 2, 0, 3, 4)))
    assert (st_again[was_out] == 1).any()
    ps = TS.adversarial(pts0[:64].copy(), rng)
    ss = OT.trace_stereo(orc, fr, TS.K33(), True, ps)
    assert {0, 1, 2, 3, 4} <= set(np.unique(ss)) | {4}

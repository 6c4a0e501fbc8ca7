"""CPU tests of the restated CoarseTracker (oracle/tracker.cpp): internal consistency that pins the
restatement where the reference offers no golden vectors."""
import numpy as np
import pytest
import oracle_py as O
import synth


@pytest.fixture(scope="module")
def setup(scene, frames):
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    f0, f1 = orc.frame_new(), orc.frame_new()
    orc.make_images(f0, frames[0][0])
    orc.make_images(f1, frames[1][0])
    rng = np.random.default_rng(1)
    pts = synth.pick_points(rng, frames[0][1], 2000)
    orc.tracker_set_ref(f0, pts)
    Ttrue = synth.T_rel(synth.camera_pose(0), synth.camera_pose(1))
    return orc, f0, f1, pts, Ttrue


def test_pyramid_shape_is_kitti_5_levels(setup):
    orc = setup[0]
    assert orc.levels == 5
    assert [orc.level_size(l) for l in range(5)] == [(1232, 368), (616, 184), (308, 92), (154, 46), (77, 23)]


def test_make_images_level0_is_input_and_gradients_are_central_differences(setup, frames):
    orc, f0 = setup[0], setup[1]
    img = frames[0][0]
    dI, ag = orc.frame_get(f0, 0)
    assert np.array_equal(dI[..., 0], img)
    assert np.array_equal(dI[1:-1, 1:-1, 1], 0.5 * (img[1:-1, 2:] - img[1:-1, :-2]))
    assert np.array_equal(dI[1:-1, 1:-1, 2], 0.5 * (img[2:, 1:-1] - img[:-2, 1:-1]))
    # row wrap at the image sides (HessianBlocks.cpp:182-184): x=0 reads the previous row's last pixel
    flat = img.reshape(-1)
    w = img.shape[1]
    idx = 5 * w
    assert dI[5, 0, 1] == np.float32(0.5) * (flat[idx + 1] - flat[idx - 1])
    assert np.array_equal(ag[1:-1], dI[1:-1, :, 1] ** 2 + dI[1:-1, :, 2] ** 2)
    l1, _ = orc.frame_get(f0, 1)
    ref = np.float32(0.25) * (((img[0::2, 0::2] + img[0::2, 1::2]) + img[1::2, 0::2]) + img[1::2, 1::2])
    assert np.array_equal(l1[..., 0], ref)


def test_template_counts_and_raster_order(setup):
    orc, pts = setup[0], setup[3]
    u, v, idp, col = orc.tracker_get_pc(0)
    assert 2000 <= u.size <= 5 * 2000  # each splat dilates to at most 5 pixels on level 0
    key = v.astype(np.int64) * synth.W + u.astype(np.int64)
    assert np.all(np.diff(key) > 0), "pc_* must be in raster order (CoarseTracker.cpp:507-531)"
    assert np.all(idp > 0) and np.all(np.isfinite(col))


def test_sse_tracking_recovers_true_motion(setup):
    orc, f0, f1, pts, Ttrue = setup
    r = orc.track(f1, np.eye(4)[:3], (0, 0), orc.levels - 1, [np.nan] * 5, 0)
    assert r["ok"]
    assert np.abs(r["T"][:, 3] - Ttrue[:, 3]).max() < 2e-3
    assert np.abs(r["T"][:, :3] - Ttrue[:, :3]).max() < 2e-4


def test_identity_sse_jacobian_equals_g2o_edge_jacobian(setup):
    """SURVEY.md §8c anchor: E1's 1x6 pose row equals calcGSSSE's first six J entries and the photometric
    columns agree (dso_g2o_edge.cpp:483-499 vs CoarseTracker.cpp:564-575) when both use the same K."""
    orc, f0, f1, pts, Ttrue = setup
    lvl = 1
    res = orc.calc_res_gs(f1, lvl, Ttrue, (0.01, 2.0), 1e9)
    wb = res["warped"]
    n = int(res["rs"][1])
    idp, u, v, dx, dy, r, hw, ref = wb[:, :n].astype(np.float64)
    K, _ = orc.level_K(lvl)
    fx, fy = float(K[0, 0]), float(K[1, 1])
    dxf, dyf = dx * fx, dy * fy
    Jsse = np.stack([idp * dxf, idp * dyf, -idp * (u * dxf + v * dyf), -(u * v * dxf + dyf * (1 + v * v)), u * v * dyf + dxf * (1 + u * u), u * dyf - v * dxf], 1)
    err, J = orc.edge_eval(f1, lvl, Ttrue, Ttrue, (0.01, 2.0))
    assert err.size == n  # every in-border point has an edge and a finite intensity
    # f64 vs f32 projection: ~1e-4 px of coordinate noise times the local gradient
    assert np.median(np.abs(err - r)) < 1e-3 and np.abs(err - r).max() < 0.1
    scale = np.abs(Jsse).max()
    assert np.abs(J[:, :6] - Jsse).max() < 2e-3 * scale
    a = np.exp(0.01)
    assert np.allclose(J[:, 6], a * (0.0 - ref), rtol=1e-6)
    assert np.all(J[:, 7] == -1)


def smooth_image(w, h, seed):
    """Low-frequency image (wavelengths >= 60 px): central differences ~ true derivatives."""
    rng = np.random.default_rng(seed)
    x, y = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    img = np.full((h, w), 128.0)
    for _ in range(6):
        lam = rng.uniform(60, 200)
        ang = rng.uniform(0, np.pi)
        img += 15 * np.sin(2 * np.pi / lam * (x * np.cos(ang) + y * np.sin(ang)) + rng.uniform(0, 6.28))
    return img.astype(np.float32)


def test_edge_jacobian_finite_differences(setup, frames):
    """Central differences of EdgeSE3PosePhotoDSO::computeError against linearizeOplus on a smooth image
    (on textured images the bilinear-interpolated central-difference gradient is not the derivative of the
    bilinear-interpolated intensity, so the comparison is only meaningful where the image is band-limited)."""
    _, _, _, pts, Ttrue = setup
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    f0, f1 = orc.frame_new(), orc.frame_new()
    orc.make_images(f0, smooth_image(synth.W, synth.H, 1))
    orc.make_images(f1, smooth_image(synth.W, synth.H, 2))
    orc.tracker_set_ref(f0, pts)
    lvl = 0
    photo = (0.0, 0.0)
    err0, J = orc.edge_eval(f1, lvl, Ttrue, Ttrue, photo)
    scale = np.abs(J[:, :6]).max(0)
    for k in range(6):
        eps = 5e-3 if k < 3 else 5e-4  # ~0.3 px of image motion: well above the float32 coordinate quantum
        d = np.zeros(6)
        d[k] = eps
        Tp = O.se3_mul(O.se3_exp(d), Ttrue)
        Tm = O.se3_mul(O.se3_exp(-d), Ttrue)
        ep, _ = orc.edge_eval(f1, lvl, Ttrue, Tp, photo)
        em, _ = orc.edge_eval(f1, lvl, Ttrue, Tm, photo)
        fd = (ep - em) / (2 * eps)
        rel = np.abs(fd - J[:, k]) / (np.abs(J[:, k]) + 0.01 * scale[k])
        assert np.median(rel) < 2e-2, (k, np.median(rel))
        assert np.percentile(rel, 90) < 1e-1, (k, np.percentile(rel, 90))
    for k, eps_k in ((6, 1e-2), (7, 1e-1)):  # large steps: ab is cast to float inside the edge
        ph_p, ph_m = list(photo), list(photo)
        ph_p[k - 6] += eps_k
        ph_m[k - 6] -= eps_k
        ep, _ = orc.edge_eval(f1, lvl, Ttrue, Ttrue, ph_p)
        em, _ = orc.edge_eval(f1, lvl, Ttrue, Ttrue, ph_m)
        fd = (ep - em) / (2 * eps_k)
        assert np.allclose(fd, J[:, k], rtol=2e-3, atol=1e-3)


def test_g2o_variant_improves_from_near_truth(setup):
    orc, f0, f1, pts, Ttrue = setup
    rng = np.random.default_rng(5)
    Tn = synth.perturb_T(Ttrue, rng, 0.02, np.deg2rad(0.1))
    r = orc.track(f1, Tn, (0, 0), orc.levels - 1, [np.nan] * 5, 1)
    assert r["ok"]
    assert np.abs(r["T"][:, 3] - Ttrue[:, 3]).max() < np.abs(Tn[:, 3] - Ttrue[:, 3]).max()


def test_g2o_lm_trial_accounting(setup, frames):
    """The restated g2o LM reports its damping trials (test infrastructure: the GPU cases of
    tests/test_gpu_tracker.py::test_track_g2o_rejected_trials rely on these scenarios containing rejected trials).
    Per level the LM runs at most 2 iterations of at most 10 trials; a converged run near the truth accepts all of them."""
    orc, f0, f1, pts, Ttrue = setup
    T0 = synth.perturb_T(Ttrue, np.random.default_rng(5), 0.02, np.deg2rad(0.1))
    r = orc.track(f1, T0, (0.0, 0.0), orc.levels - 1, [np.nan] * 5, 1)
    trials, rejected = orc.g2o_trial_counts()
    assert r["ok"] and rejected == 0 and trials == int(r["iterations"].sum())
    # a larger offset: trials are rejected (lambda grows) and the run still converges
    T1 = synth.perturb_T(Ttrue, np.random.default_rng(0), 0.3, np.deg2rad(1.5))
    r = orc.track(f1, T1, (0.0, 0.0), orc.levels - 1, [np.nan] * 5, 1)
    trials, rejected = orc.g2o_trial_counts()
    assert r["ok"] and rejected >= 3 and trials - rejected >= 1
    assert trials <= 10 * int(r["iterations"].sum())
    assert np.abs(r["T"][:, 3] - Ttrue[:, 3]).max() < 1e-2

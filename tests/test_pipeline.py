"""SURVEY config 1 as a harness over the operators (tests/pipeline.py): a synthetic KITTI-shape stereo sequence run frame by frame —
both pyramids, tracking, epipolar search, and at every 5th frame the whole key-frame chain (candidate filter, activation,
windowed optimisation, marginalisation, pixel selection, static stereo) — once on the oracle, once on the device, each backend
consuming only its own previous results.

north_star: final trajectories within 1e-4 m / 1e-5 rad per frame. The chain contains the windowed solve, whose float sums the
reference itself does not reproduce from run to run (tests/test_oracle_spread.py); the per-frame bound is therefore
max(north_star, 3 x the oracle's own per-frame spread under the reference's 6-worker accumulation), with the spread measured here by
re-running the oracle chain under other worker assignments."""
import numpy as np
import pytest
import synth
import pipeline as PL


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def run(backend, left, right, **kw):
    P = PL.StereoPipeline(backend, **kw)
    oks = []
    for l, r in zip(left, right):
        oks.append(P.step(l, r)["ok"])
    return P, oks


def test_oracle_pipeline_follows_the_true_path(scene):
    """CPU: 16 stereo frames at 640x192 through the whole chain (window of 4, so frames and points get marginalised)."""
    shape = dict(w=640, h=192, K=(360.0, 360.0, 319.5, 95.5))
    poses = [synth.camera_pose(0.25 * k) for k in range(16)]
    left = [synth.render(scene, p, shape["w"], shape["h"], shape["K"])[0] for p in poses]
    right = [synth.render(scene, synth.right_of(p), shape["w"], shape["h"], shape["K"])[0] for p in poses]
    P, oks = run(PL.Backend(shape), left, right, kf_every=3, max_kf=4, immature_density=600, point_density=800)
    assert all(oks)
    T0 = np.vstack([synth.T_cw(poses[0]), [0, 0, 0, 1]])
    for k in range(16):
        true_c2w = np.linalg.inv(np.vstack([synth.T_cw(poses[k]), [0, 0, 0, 1]]) @ np.linalg.inv(T0))
        assert np.abs(P.traj[k][:3, 3] - true_c2w[:3, 3]).max() < 3e-2, k
    kf = [e for e in P.log if "rmse" in e]
    assert len(kf) == 5 and sum(e.get("activated", 0) for e in kf) > 500 and sum(e.get("marginalized", 0) for e in kf) > 300
    assert all(np.isfinite(e["rmse"]) and e["rmse"] < 8 for e in kf)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_device_200_frame_sequence_matches_oracle(pkg, scene, variant):
    """Config 1: 200 stereo frames, 1232x368, key frame every 5th frame, window of 7, 1500 immature / 2000 active points."""
    shape = dict(w=synth.W, h=synth.H, K=synth.K4)
    N = 200
    poses = [synth.camera_pose(0.2 * k) for k in range(N)]
    left = [synth.render_torch(scene, p)[0] for p in poses]
    right = [synth.render_torch(scene, synth.right_of(p))[0] for p in poses]
    kw = dict(kf_every=5, max_kf=7, immature_density=1500.0, point_density=2000.0, variant=variant)
    Po, oko = run(PL.Backend(shape), left, right, **kw)
    Bg = PL.Backend(shape, pkg)
    Pg, okg = run(Bg, left, right, **kw)
    Bg.close()
    assert okg == oko and all(oko)
    # the oracle's own spread: the same chain under three other worker assignments of the float accumulators. The chain is
    # chaotic in the usual VO sense — a residual that flips IN/OUTLIER changes which points get activated at the next key frame —
    # so the oracle's own trajectories drift apart with the frame index (1e-4 m is exceeded after ~10 frames, see the printout).
    seeds = (0, 1, 2)
    runs = [run(PL.Backend(shape, reduce=(6, seed)), left, right, **kw)[0] for seed in seeds]
    spread_t = np.max([[np.abs(Ps.traj[k][:3, 3] - Po.traj[k][:3, 3]).max() for k in range(N)] for Ps in runs], axis=0)
    spread_r = np.max([[rot_angle(Ps.traj[k][:3, :3], Po.traj[k][:3, :3]) for k in range(N)] for Ps in runs], axis=0)
    env_t = np.maximum(1e-4, 3 * np.maximum.accumulate(spread_t))   # north_star, or 3 x the largest spread seen up to that frame
    env_r = np.maximum(1e-5, 3 * np.maximum.accumulate(spread_r))
    dt = np.array([np.abs(Pg.traj[k][:3, 3] - Po.traj[k][:3, 3]).max() for k in range(N)])
    dr = np.array([rot_angle(Pg.traj[k][:3, :3], Po.traj[k][:3, :3]) for k in range(N)])
    first_t = int(np.argmax(spread_t > 1e-4)) if (spread_t > 1e-4).any() else N
    print(f"variant {variant}: device vs oracle max |dt| {dt.max():.2e} m, max dR {dr.max():.2e} rad; oracle vs itself max |dt| {spread_t.max():.2e} m, "
          f"max dR {spread_r.max():.2e} rad (first frame above 1e-4 m: {first_t}); device frames within 1e-4 m / 1e-5 rad: {(dt < 1e-4).mean():.3f} / {(dr < 1e-5).mean():.3f}")
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(variant=variant, frames=N, device_vs_oracle_dt=dt.tolist(), device_vs_oracle_dR=dr.tolist(), oracle_spread_dt=spread_t.tolist(),
                   oracle_spread_dR=spread_r.tolist(), keyframes_oracle=Po.log, keyframes_device=Pg.log), open(f"gpurun_out/r2_sequence200_variant{variant}.json", "w"))
    # the sequence must be a real one: the path is followed and the map turns over
    T0 = np.vstack([synth.T_cw(poses[0]), [0, 0, 0, 1]])
    true_end = np.linalg.inv(np.vstack([synth.T_cw(poses[-1]), [0, 0, 0, 1]]) @ np.linalg.inv(T0))
    assert np.abs(Po.traj[-1][:3, 3] - true_end[:3, 3]).max() < 0.05 * np.linalg.norm(true_end[:3, 3])
    assert np.abs(Pg.traj[-1][:3, 3] - true_end[:3, 3]).max() < 0.05 * np.linalg.norm(true_end[:3, 3])
    assert sum(e.get("marginalized", 0) for e in Po.log) > 5000
    # while the oracle reproduces itself to 1e-4 m the device must do so too; afterwards it stays inside the oracle's own envelope
    assert np.all(dt <= env_t), (int(np.argmax(dt - env_t)), float(dt.max()), float(spread_t.max()))
    assert np.all(dr <= env_r), (int(np.argmax(dr - env_r)), float(dr.max()), float(spread_r.max()))
    # bookkeeping: exact at the key frames where the oracle's own re-runs still agree with each other (nothing upstream has diverged
    # yet); afterwards the per-key-frame counts are as chaotic as the trajectories, so the totals over the sequence are compared
    keys = ("new_immature", "stereo_good", "candidates", "to_optimize", "activated", "points", "ref_points", "marginalized")
    exact = 0
    for i, (eo, eg) in enumerate(zip(Po.log, Pg.log)):
        if all(Ps.log[i].get(key) == eo.get(key) for Ps in runs for key in keys):
            for key in keys:
                assert eo.get(key) == eg.get(key), (i, key, eo, eg)
            exact += 1
    assert exact >= 2
    for key in ("new_immature", "activated", "marginalized", "points"):
        so = sum(e.get(key, 0) for e in Po.log); sg = sum(e.get(key, 0) for e in Pg.log)
        ssp = max(abs(sum(e.get(key, 0) for e in Ps.log) - so) for Ps in runs)
        assert abs(sg - so) <= max(3 * ssp, 0.03 * so), (key, so, sg, ssp)

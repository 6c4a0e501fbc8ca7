"""Trajectory-level parity ('pipeline'): a synthetic stereo-free mini sequence is tracked frame by frame against
key frames that are switched every few frames, each frame initialised from the previous frame's result — once with the device path,
once with the oracle. north_star: final trajectories within 1e-4 m / 1e-5 rad per frame."""
import numpy as np
import pytest
import oracle_py as O
import synth

W_, H_, K_ = 640, 192, (360.0, 360.0, 319.5, 95.5)
NFRAMES, KF_EVERY, STEP = 19, 6, 0.25


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1) / 2
    return float(np.arccos(np.clip(c, -1, 1)))


def run_sequence(api, is_device, imgs, depths, poses, variant):
    new_frame = api.frame_create if is_device else api.frame_new
    traj, oks = [], []
    fid_kf = None
    T_kf_cur = np.eye(4)[:3]
    rng = np.random.default_rng(4)
    for k in range(NFRAMES):
        f = new_frame()
        api.make_images(f, imgs[k])
        if k % KF_EVERY == 0:
            # new key frame: template from the true depth (the mapping back-end is not part of this test)
            api.tracker_set_ref(f, synth.pick_points(rng, depths[k], 1200, W_, H_), (0.0, 0.0))
            fid_kf, kf_index = f, k
            T_kf_cur = np.eye(4)[:3]
            traj.append(np.vstack([synth.T_cw(poses[k]), [0, 0, 0, 1]]))  # key-frame world pose is given
            oks.append(True)
            continue
        r = api.track(f, T_kf_cur, (0.0, 0.0), api.levels - 1, [np.nan] * 5, variant)
        T_kf_cur = r["T"].copy()     # constant-position model: the next frame starts from this result
        oks.append(bool(r["ok"]))
        T_w_kf = np.vstack([synth.T_cw(poses[kf_index]), [0, 0, 0, 1]])
        traj.append(np.vstack([r["T"], [0, 0, 0, 1]]) @ T_w_kf)
    return np.stack(traj), oks


@pytest.fixture(scope="module")
def sequence(scene):
    poses = [synth.camera_pose(STEP * k) for k in range(NFRAMES)]
    rend = [synth.render(scene, p, W_, H_, K_) for p in poses]
    return [r[0] for r in rend], [r[1] for r in rend], poses


def test_oracle_sequence_follows_the_true_path(sequence):
    imgs, depths, poses = sequence
    traj, oks = run_sequence(O.Oracle(W_, H_, K_, synth.BASELINE), False, imgs, depths, poses, 0)
    assert all(oks)
    for k in range(NFRAMES):
        Tt = synth.T_cw(poses[k])
        assert np.abs(traj[k][:3, 3] - Tt[:, 3]).max() < 2e-2, k


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_device_trajectory_matches_oracle(pkg, sequence, variant):
    imgs, depths, poses = sequence
    to, oko = run_sequence(O.Oracle(W_, H_, K_, synth.BASELINE), False, imgs, depths, poses, variant)
    ctx = pkg.Context(W_, H_, K_, synth.BASELINE)
    tg, okg = run_sequence(ctx, True, imgs, depths, poses, variant)
    ctx.close()
    assert okg == oko
    for k in range(NFRAMES):
        assert np.abs(tg[k][:3, 3] - to[k][:3, 3]).max() < 1e-4, (k, float(np.abs(tg[k][:3, 3] - to[k][:3, 3]).max()))
        assert rot_angle(tg[k][:3, :3], to[k][:3, :3]) < 1e-5, k

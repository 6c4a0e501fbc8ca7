"""Synthetic sliding window for the BA tests/bench: n keyframes along the path, P active
points spread evenly over the hosts with colour/weights from the ImmaturePoint constructor, residuals to every other
keyframe, states perturbed from the evaluation points. Pure data: a list of frames/points/residuals that both the
oracle window and the device window are filled from."""
import numpy as np
import synth


def make_window(scene, n=7, P=2000, seed=7, spacing=0.8, w=synth.W, h=synth.H, K=synth.K4, idepth_noise=0.02, state_sigma=1e-3,
                images=None):
    rng = np.random.default_rng(seed)
    poses = [synth.camera_pose(k * spacing) for k in range(n)]
    if images is None:
        images = [synth.render(scene, p, w, h, K) for p in poses]
    frames = []
    for k in range(n):
        st = np.zeros(10)
        if k > 0:  # perturb the state around the evaluation point (unscaled state units)
            st[:6] = rng.normal(0, state_sigma, 6)
            st[6:8] = rng.normal(0, state_sigma, 2) * np.array([0.1, 0.01])
        frames.append(dict(T_w2c=synth.T_cw(poses[k]), image=images[k][0], depth=images[k][1], a=0.0, b=0.0, frameID=k + 1,
                           state=st, energyTH=8 * 8 * 8))
    points = []
    per = P // n
    for hidx in range(n):
        depth = frames[hidx]["depth"]
        us = rng.integers(8, w - 9, per * 2)
        vs = rng.integers(8, h - 9, per * 2)
        seen = set()
        cnt = 0
        for u, v in zip(us, vs):
            if (u, v) in seen or cnt >= per:
                continue
            seen.add((u, v))
            idz = float(1.0 / depth[v, u])
            idz_zero = idz * (1 + rng.normal(0, idepth_noise))
            idd = idz_zero * (1 + rng.normal(0, 1e-3))
            points.append(dict(host=hidx, u=float(u), v=float(v), idepth=np.float32(idd), idepth_zero=np.float32(idz_zero),
                               has_prior=bool(rng.random() < 0.1), targets=[t for t in range(n) if t != hidx]))
            cnt += 1
    return dict(n=n, frames=frames, points=points, poses=poses)


def fill_oracle(win, orc, OBA, immature_init):
    """Build the oracle-side window; returns (OracleBA, frame ids, per-point (color, weights))."""
    fids = []
    for f in win["frames"]:
        fid = orc.frame_new()
        orc.make_images(fid, f["image"])
        fids.append(fid)
    ba = OBA(orc)
    for k, f in enumerate(win["frames"]):
        idx = ba.add_frame(fids[k], f["T_w2c"], f["a"], f["b"], f["frameID"])
        ba.set_state(idx, f["state"])
        ba.set_energy_th(idx, f["energyTH"])
    cw = []
    for p in win["points"]:
        ok, col, wts, _, _ = immature_init(orc, fids[p["host"]], p["u"], p["v"])
        cw.append((col, wts))
        pi = ba.add_point(p["host"], p["u"], p["v"], p["idepth"], p["idepth_zero"], col, wts, p["has_prior"])
        for t in p["targets"]:
            ba.add_residual(pi, t)
    ba.prepare()
    return ba, fids, cw


def fill_device(win, ctx, Window, cw):
    """Build the device-side window from the same data (colour/weights `cw` as produced by the D1 operator)."""
    fids = []
    for f in win["frames"]:
        fid = ctx.frame_create()
        ctx.make_images(fid, f["image"])
        fids.append(fid)
    W = Window(ctx)
    for k, f in enumerate(win["frames"]):
        idx = W.add_frame(fids[k], f["T_w2c"], f["a"], f["b"], f["frameID"])
        W.set_state(idx, f["state"])
        W.set_energy_th(idx, f["energyTH"])
    pts = win["points"]
    W.set_points([p["host"] for p in pts], [p["u"] for p in pts], [p["v"] for p in pts], [p["idepth"] for p in pts],
                 [p["idepth_zero"] for p in pts], np.stack([c for c, _ in cw]), np.stack([w for _, w in cw]), [p["has_prior"] for p in pts])
    rp, rt = [], []
    for pi, p in enumerate(pts):
        for t in p["targets"]:
            rp.append(pi); rt.append(t)
    W.set_residuals(rp, rt)
    W.prepare()
    return W, fids


def shard_window(win, begin, end):
    """The sub-window a rank owns under point sharding: all frames, points [begin, end) of allPoints."""
    out = dict(win)
    out["points"] = win["points"][begin:end]
    return out

"""Seeded synthetic KITTI-shape stereo data — test/bench infrastructure.

A convex textured "street corridor" (ground, two walls, ceiling, end wall) rendered by analytic
ray/plane intersection, with a band-limited procedural texture defined in world coordinates, so
every view of a surface point has the same intensity (photometric consistency is exact up to
sampling/quantisation) and the true depth of every pixel is known.

Working resolution 1232x368 (5 pyramid levels): the reference cannot run at 1241x376 internally
(util/globalCalib.cpp:53 halves only while both sides are even), so the 1241x376 raw image is
centre-cropped; here the crop is folded into the principal point.
"""
import numpy as np

SEED = 20260118
W, H = 1232, 368
# KITTI intrinsics with the 1241x376 -> 1232x368 crop (4 px left/top) folded into cx, cy
FX, FY, CX, CY = 718.856, 718.856, 607.1928 - 4.0, 185.2157 - 4.0
BASELINE = 0.537
K4 = (FX, FY, CX, CY)


def make_scene(seed=SEED, n_waves=20):
    rng = np.random.default_rng(seed)
    # planes n.X = c in world coordinates (x right, y down, z forward)
    planes = [
        (np.array([0.0, 1.0, 0.0]), 1.65, (0, 2)),    # ground  y = 1.65   texture coords (x,z)
        (np.array([1.0, 0.0, 0.0]), -6.0, (2, 1)),    # left wall  x = -6  (z,y)
        (np.array([1.0, 0.0, 0.0]), 7.5, (2, 1)),     # right wall x = 7.5
        (np.array([0.0, 1.0, 0.0]), -7.0, (0, 2)),    # ceiling y = -7
        (np.array([0.0, 0.0, 1.0]), 420.0, (0, 1)),   # end wall z = 420
    ]
    tex = []
    for _ in planes:
        lam = np.exp(rng.uniform(np.log(0.45), np.log(6.0), n_waves))  # wavelengths [m]
        ang = rng.uniform(0, np.pi, n_waves)
        f = 2 * np.pi / lam
        amp = rng.uniform(0.5, 1.0, n_waves) * (lam ** 0.35)
        amp *= 40.0 * np.sqrt(2.0) / np.sqrt((amp ** 2).sum())  # sigma ~ 40 grey levels
        tex.append(dict(fx=f * np.cos(ang), fy=f * np.sin(ang), ph=rng.uniform(0, 2 * np.pi, n_waves), amp=amp))
    return dict(planes=planes, tex=tex)


def camera_pose(k, seed_shift=0.0):
    """Left camera k: returns (R_wc, o) — camera-to-world rotation and camera centre."""
    yaw = 0.035 * np.sin(0.25 * k + seed_shift)
    pitch = 0.004 * np.sin(0.17 * k + 1.0 + seed_shift)
    o = np.array([0.3 * np.sin(0.1 * k + seed_shift), 0.02 * np.sin(0.3 * k), 1.0 * k])
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    return Ry @ Rx, o


def right_of(pose, baseline=BASELINE):
    R, o = pose
    return R, o + R @ np.array([baseline, 0.0, 0.0])


def T_rel(pose_a, pose_b):
    """3x4 [R|t] taking points in camera a to camera b."""
    Ra, oa = pose_a
    Rb, ob = pose_b
    R = Rb.T @ Ra
    t = Rb.T @ (oa - ob)
    return np.hstack([R, t[:, None]])


def T_cw(pose):
    R, o = pose
    return np.hstack([R.T, (-R.T @ o)[:, None]])


def render(scene, pose, w=W, h=H, K=K4, quantise=True, exposure_ab=(0.0, 0.0)):
    """Returns (image float32 [h,w] in 0..255, depth float32 [h,w])."""
    fx, fy, cx, cy = K
    R, o = pose
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    d_c = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
    d_w = d_c @ R.T
    best_s = np.full((h, w), np.inf)
    best_id = np.zeros((h, w), np.int32)
    for i, (n, c, _) in enumerate(scene["planes"]):
        denom = d_w @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (c - n @ o) / denom
        s = np.where((s > 1e-6) & np.isfinite(s), s, np.inf)
        upd = s < best_s
        best_s = np.where(upd, s, best_s)
        best_id = np.where(upd, i, best_id)
    X = o[None, None, :] + d_w * best_s[..., None]
    img = np.full((h, w), 128.0)
    for i, (_, _, axes) in enumerate(scene["planes"]):
        m = best_id == i
        if not m.any():
            continue
        p0, p1 = X[..., axes[0]][m], X[..., axes[1]][m]
        t = scene["tex"][i]
        val = (t["amp"][None, :] * np.sin(p0[:, None] * t["fx"][None, :] + p1[:, None] * t["fy"][None, :] + t["ph"][None, :])).sum(1)
        img[m] += val
    a, b = exposure_ab
    img = np.exp(a) * img + b
    img = np.clip(img, 0, 255)
    if quantise:
        img = np.round(img)
    return img.astype(np.float32), best_s.astype(np.float32)


def pick_points(rng, depth, n, w=W, h=H, hdi_range=(1e-4, 1e-2)):
    """n tracker splats {u, v, idepth, weight} at distinct integer pixels."""
    us = np.arange(4, w - 5)
    vs = np.arange(4, h - 4)
    flat = rng.choice(us.size * vs.size, size=n, replace=False)
    u = us[flat % us.size].astype(np.float32)
    v = vs[flat // us.size].astype(np.float32)
    idepth = (1.0 / depth[v.astype(int), u.astype(int)]).astype(np.float32)
    hdi = rng.uniform(hdi_range[0], hdi_range[1], n)
    weight = np.sqrt(1e-3 / (hdi + 1e-12)).astype(np.float32)  # CoarseTracker.cpp:350
    return np.stack([u, v, idepth, weight], 1).astype(np.float32)


def perturb_T(T, rng, sigma_t=0.05, sigma_r=np.deg2rad(0.5)):
    """Left-multiply T by a small random motion (xi ~ N(0, diag(sigma_t, sigma_r)))."""
    xi_t = rng.normal(0, sigma_t, 3)
    w = rng.normal(0, sigma_r, 3)
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    Rr = np.eye(3) + (np.sin(th) / th) * Kx + ((1 - np.cos(th)) / th ** 2) * Kx @ Kx if th > 0 else np.eye(3)
    R, t = T[:, :3], T[:, 3]
    return np.hstack([Rr @ R, (Rr @ t + xi_t)[:, None]])


def render_torch(scene, pose, w=W, h=H, K=K4, device="cuda", quantise=True):
    """synth.render on a CUDA device (float64), for the long sequences of the -m gpu tests and the bench: the same ray caster, ~1000x
    faster than numpy. Returns (image float32 [h,w], depth float32 [h,w]) as numpy arrays. Not bit-identical to render() (different
    sin / summation order) — callers render every frame of a run with ONE of the two."""
    import torch
    fx, fy, cx, cy = K
    R = torch.as_tensor(pose[0], dtype=torch.float64, device=device)
    o = torch.as_tensor(pose[1], dtype=torch.float64, device=device)
    v, u = torch.meshgrid(torch.arange(h, dtype=torch.float64, device=device), torch.arange(w, dtype=torch.float64, device=device), indexing="ij")
    d_c = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1)
    d_w = d_c @ R.T
    best_s = torch.full((h, w), float("inf"), dtype=torch.float64, device=device)
    best_id = torch.zeros((h, w), dtype=torch.int64, device=device)
    for i, (n, c, _) in enumerate(scene["planes"]):
        nt = torch.as_tensor(n, dtype=torch.float64, device=device)
        s = (c - nt @ o) / (d_w @ nt)
        s = torch.where((s > 1e-6) & torch.isfinite(s), s, torch.full_like(s, float("inf")))
        upd = s < best_s
        best_s = torch.where(upd, s, best_s)
        best_id = torch.where(upd, torch.full_like(best_id, i), best_id)
    X = o[None, None, :] + d_w * best_s[..., None]
    img = torch.full((h, w), 128.0, dtype=torch.float64, device=device)
    for i, (_, _, axes) in enumerate(scene["planes"]):
        m = best_id == i
        if not bool(m.any()):
            continue
        t = scene["tex"][i]
        p0, p1 = X[..., axes[0]][m], X[..., axes[1]][m]
        tfx, tfy, tph, tamp = (torch.as_tensor(t[k], dtype=torch.float64, device=device) for k in ("fx", "fy", "ph", "amp"))
        val = (tamp[None, :] * torch.sin(p0[:, None] * tfx[None, :] + p1[:, None] * tfy[None, :] + tph[None, :])).sum(1)
        img[m] += val
    img = img.clamp(0, 255)
    if quantise:
        img = torch.round(img)
    return img.to(torch.float32).cpu().numpy(), best_s.to(torch.float32).cpu().numpy()

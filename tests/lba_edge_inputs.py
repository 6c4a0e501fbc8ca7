"""Vertex estimates for the E2 operator tests: host poses T_wh (camera-to-world of every frame, slightly perturbed), host
photometric vertices, one inverse-depth vertex per residual, the camera vertex, b0 per host."""
import numpy as np
import synth


def make(win, seed=0):
    rng = np.random.default_rng(seed)
    n = win["n"]
    T_wh = []
    for f in win["frames"]:
        T = f["T_w2c"]
        R, t = T[:, :3], T[:, 3]
        Tc2w = np.hstack([R.T, (-R.T @ t)[:, None]])
        T_wh.append(synth.perturb_T(Tc2w, rng, 2e-3, 2e-4))
    photo = np.stack([rng.normal(0, 0.01, n), rng.normal(0, 1.0, n)], 1)
    idepth = []
    for p in win["points"]:
        for _ in p["targets"]:
            idepth.append(float(p["idepth"]) * (1 + rng.normal(0, 5e-3)))
    return np.stack(T_wh), photo, np.array(idepth), rng.normal(0, 0.5, n)


def cam_vertex(K4, seed=0):
    rng = np.random.default_rng(seed + 100)
    return np.array(K4, np.float64) + rng.normal(0, 0.05, 4)

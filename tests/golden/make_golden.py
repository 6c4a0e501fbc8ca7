"""Generates tests/golden/hotpath_v1.npz — small committed fixtures of the hot path.

PROVENANCE: the reference (gyubeomim/stereo-dso-g2o) holds no golden vectors for this path and cannot be compiled in this
environment (Eigen / g2o / Boost / OpenCV absent), so these vectors are produced by the ORACLE (the CPU restatement under
oracle/), not by the reference itself: they pin the restatement and the device path against drift, they do not pin either
to the reference ("parity unpinned"). Run from the repo root: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O          # noqa: E402
import oracle_trace_py as OT   # noqa: E402

W, H = 320, 192
K4 = (300.0, 300.0, 159.5, 95.5)
BASELINE = 0.5


def images(seed=11):
    """Band-limited random texture and a copy translated by a known sub-pixel-free shift (a fronto-parallel stereo pair)."""
    rng = np.random.default_rng(seed)
    big = rng.normal(0, 1, (H + 64, W + 64))
    k = np.exp(-0.5 * (np.arange(-6, 7) / 2.0) ** 2); k /= k.sum()
    for ax in (0, 1):
        big = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, big)
    big = 128 + 60 * big / big.std()
    left = np.round(np.clip(big[32:32 + H, 32:32 + W], 0, 255)).astype(np.float32)
    right = np.round(np.clip(big[32:32 + H, 32 + 6:32 + 6 + W], 0, 255)).astype(np.float32)  # disparity 6 px
    return left, right


def build():
    left, right = images()
    orc = O.Oracle(W, H, K4, BASELINE)
    fl, fr = orc.frame_new(), orc.frame_new()
    orc.make_images(fl, left); orc.make_images(fr, right)
    out = dict(left=left.astype(np.uint8), right=right.astype(np.uint8), K4=np.array(K4), baseline=np.array(BASELINE), levels=np.array(orc.levels))
    for lvl in range(orc.levels):
        dI, ag = orc.frame_get(fl, lvl)
        out[f"dI_l{lvl}_sub"] = dI[1:-1:5, ::7].copy()
        out[f"ag_l{lvl}_sub"] = ag[1:-1:5, ::7].copy()
        out[f"dI_l{lvl}_sum"] = dI[1:-1].astype(np.float64).sum(axis=(0, 1))
    rng = np.random.default_rng(3)
    xy = np.stack([rng.uniform(3, W - 4, 64), rng.uniform(3, H - 4, 64)], 1).astype(np.float32)
    out["interp_xy"] = xy
    out["interp33"] = orc.interp33(fl, 0, xy)
    out["interp33bilin"] = orc.interp33(fl, 0, xy, bilin=True)
    uv = np.stack([rng.integers(12, W - 24, 48), rng.integers(12, H - 12, 48)], 1).astype(np.float32)
    pts, ok = OT.immature_init(orc, fl, uv)
    out["imm_uv"] = uv; out["imm_init"] = pts.copy().view(np.uint8); out["imm_ok"] = ok
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    st = OT.trace_stereo(orc, fr, K, True, pts)
    out["stereo_status"] = st; out["stereo_pts"] = pts.copy().view(np.uint8)
    # tracker template + one fused calcRes/calcGS evaluation at a small motion
    n = 300
    u = rng.integers(6, W - 7, n).astype(np.float32); v = rng.integers(6, H - 6, n).astype(np.float32)
    splats = np.stack([u, v, np.full(n, BASELINE * K4[0] / 6.0 / K4[0] / BASELINE * (1 / 1.0), np.float32) * 0 + np.float32(6.0 / (K4[0] * BASELINE)),
                       rng.uniform(0.5, 3.0, n).astype(np.float32)], 1).astype(np.float32)
    orc.tracker_set_ref(fl, splats, (0.0, 0.0))
    out["splats"] = splats
    for lvl in range(orc.levels):
        pc = orc.tracker_get_pc(lvl)
        out[f"pc_n_l{lvl}"] = np.array(pc[0].size)
        out[f"pc_l{lvl}"] = np.stack(pc, 1)
    T = np.eye(4)[:3].copy(); T[0, 3] = -BASELINE  # the "new" frame is the right camera: pure translation by the baseline
    r = orc.calc_res_gs(fr, 0, T, (0.0, 0.0), 20.0)
    out["gs_T"] = T; out["gs_rs"] = r["rs"]; out["gs_H"] = r["H"]; out["gs_b"] = r["b"]; out["gs_warped_n"] = np.array(r["warped_n"])
    out["gs_warped"] = r["warped"]
    tr = orc.track(fr, np.eye(4)[:3], (0.0, 0.0), orc.levels - 1, [np.nan] * 5, 0)
    out["track_T"] = tr["T"]; out["track_aff"] = tr["aff"]; out["track_ok"] = np.array(tr["ok"]); out["track_lastRes"] = tr["lastResiduals"]
    return out


if __name__ == "__main__":
    d = build()
    np.savez_compressed(os.path.join(HERE, "hotpath_v1.npz"), **d)
    print("wrote", os.path.join(HERE, "hotpath_v1.npz"), os.path.getsize(os.path.join(HERE, "hotpath_v1.npz")), "bytes")

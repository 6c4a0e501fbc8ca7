"""Generates tests/golden/ba_v1.npz — committed fixtures of the windowed-BA side of the path (linearize, accumulators, solve,
the LM driver, the g2o body of FullSystem::optimize, the g2o vertex / edge operators).

PROVENANCE: as for hotpath_v1.npz (make_golden.py) — the reference holds no golden vectors for this path and cannot be compiled
here, so these vectors come from the ORACLE (oracle/ba.cpp, lba_g2o.cpp, c_api.cpp), not from the reference: they pin the
restatement against drift ("parity unpinned"). The inputs (8-bit images, window description) are stored in the
file, so the check does not depend on the synthetic renderer. Run from the repo root: python tests/golden/make_golden_ba.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ba_synth            # noqa: E402
import oracle_ba_py as OB  # noqa: E402
import oracle_py as O      # noqa: E402
import synth               # noqa: E402

W, H = 320, 192
K4 = (300.0, 300.0, 159.5, 95.5)
N, P = 4, 160


def window_from_arrays(G):
    """the neutral window description ba_synth.fill_oracle / fill_device take, rebuilt from the stored arrays"""
    n = int(G["n"])
    frames = [dict(T_w2c=G["T_w2c"][k], image=G["images"][k].astype(np.float32), a=0.0, b=0.0, frameID=k + 1, state=G["states"][k], energyTH=8 * 8 * 8)
              for k in range(n)]
    points = [dict(host=int(h), u=float(u), v=float(v), idepth=np.float32(i), idepth_zero=np.float32(z), has_prior=bool(pr),
                   targets=[t for t in range(n) if t != int(h)])
              for h, u, v, i, z, pr in zip(G["p_host"], G["p_u"], G["p_v"], G["p_idepth"], G["p_idepth_zero"], G["p_prior"])]
    return dict(n=n, frames=frames, points=points)


def vertex_inputs():
    rng = np.random.default_rng(21)
    return dict(pose_est=np.stack([synth.perturb_T(np.eye(4)[:3], rng, 0.5, 0.3).reshape(12) for _ in range(6)]), pose_upd=rng.normal(0, 0.05, (6, 6)),
                photo_est=rng.normal(0, 0.1, (6, 2)), photo_upd=rng.normal(0, 0.01, (6, 2)), idepth_est=rng.uniform(0.05, 2, (6, 1)), idepth_upd=rng.normal(0, 0.3, (6, 1)),
                uv_est=rng.uniform(5, 300, (6, 2)), uv_upd=rng.normal(0, 2, (6, 1)), uv_aux=rng.normal(0, 1, (6, 2)), cam_est=np.tile(np.array(K4), (6, 1)), cam_upd=rng.normal(0, 0.5, (6, 4)))


def evaluate(G):
    """every oracle output the fixture pins, from the stored inputs"""
    win = window_from_arrays(G)
    orc = O.Oracle(W, H, K4, synth.BASELINE)
    ba, fids, cw = ba_synth.fill_oracle(win, orc, OB.OracleBA, OB.immature_init)
    out = {}
    out["E_lin"] = np.array(ba.linearize_all(False))
    r = ba.get_res(0)
    out["newState"] = r["newState"]; out["newEnergy"] = r["newEnergy"]
    ba.linearize_all(True)
    Ht, bt, _ = ba.accumulate_top(0, True)
    Hs, bs = ba.accumulate_sc(True)
    out["H_top"], out["b_top"], out["H_sc"], out["b_sc"] = Ht, bt, Hs, bs
    s = ba.solve(0)
    out["x"] = s[0] if isinstance(s, tuple) else s["x"]
    # the g2o body of FullSystem::optimize on the same window, from perturbed vertices
    st = ba.get_state()
    T_wh = np.stack([np.hstack([T[:, :3].T, (-T[:, :3].T @ T[:, 3])[:, None]]) for T in st["T_w2c"]])
    Tp = T_wh + G["lba_dT"]
    idepth = np.array([float(p["idepth"]) for p in win["points"] for _ in p["targets"]])
    o = ba.lba_g2o(np.array(K4, float), Tp, np.zeros((N, 2)), idepth, 3)
    out["lba_iterations"], out["lba_trials"], out["lba_chi2"] = np.array(o["iterations"]), np.array(o["trials"]), np.array(o["chi2"])
    out["lba_T_wh"], out["lba_cam"], out["lba_idepth"], out["lba_newState"] = o["T_wh"], o["cam"], o["idepth"], o["newState"]
    # FullSystem::optimize, SSE body, on a fresh window
    ba2, _, _ = ba_synth.fill_oracle(win, O.Oracle(W, H, K4, synth.BASELINE), OB.OracleBA, OB.immature_init)
    rmse, its = ba2.optimize(4)
    st2 = ba2.get_state()
    out["opt_rmse"], out["opt_iterations"], out["opt_states"], out["opt_idepth"] = np.array(rmse), np.array(its), st2["states"], st2["idepth"]
    # g2o vertex updates V1-V5 (dso_g2o_vertex.cpp:15-106)
    for kind, name in ((1, "pose"), (2, "photo"), (3, "idepth"), (4, "uv"), (5, "cam")):
        aux = G["uv_aux"] if name == "uv" else None
        out[f"oplus_{name}"] = O.vertex_oplus(kind, G[f"{name}_est"], G[f"{name}_upd"], aux)
    return out


def build():
    scene = synth.make_scene()
    poses = [synth.camera_pose(k * 0.3) for k in range(N)]
    rend = [synth.render(scene, p, W, H, K4) for p in poses]
    images8 = [np.round(np.clip(im, 0, 255)).astype(np.uint8) for im, _ in rend]
    win = ba_synth.make_window(scene, n=N, P=P, seed=5, spacing=0.3, w=W, h=H, K=K4, idepth_noise=0.004, state_sigma=3e-4, images=[(i8.astype(np.float32), d) for i8, (_, d) in zip(images8, rend)])
    pts = win["points"]
    G = dict(n=np.array(N), images=np.stack(images8), T_w2c=np.stack([f["T_w2c"] for f in win["frames"]]), states=np.stack([f["state"] for f in win["frames"]]),
             p_host=np.array([p["host"] for p in pts], np.int32), p_u=np.array([p["u"] for p in pts], np.float32), p_v=np.array([p["v"] for p in pts], np.float32),
             p_idepth=np.array([p["idepth"] for p in pts], np.float32), p_idepth_zero=np.array([p["idepth_zero"] for p in pts], np.float32),
             p_prior=np.array([p["has_prior"] for p in pts], np.uint8), lba_dT=np.random.default_rng(0).normal(0, 1e-3, (N, 3, 4)) * np.array([0, 0, 0, 1.0]))
    G.update(vertex_inputs())
    G.update({"out_" + k: v for k, v in evaluate(G).items()})
    return G


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "ba_v1.npz"), **build())

"""Windowed-BA LM iteration (linearizeAll + accumulate + stitch + solve + resubstitute): device vs the oracle port on one host
thread, at SURVEY config 3 (7 KF, ~2k points) and config 4 (10 KF, 20k points). Prints one JSON line per config."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import conftest, synth, ba_synth
import oracle_py as O, oracle_ba_py as OB

ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=50); ap.add_argument("--big", type=int, default=1)
a = ap.parse_args()
pkg = conftest.load_pkg()
scene = synth.make_scene()
cfgs = [("config3", 7, 2002, synth.W, synth.H, synth.K4)]
if a.big: cfgs.append(("config4", 10, 20000, 1920, 1088, (360.0 * 3, 360.0 * 3, 959.5, 543.5)))
for name, n, P, w, h, K in cfgs:
    win = ba_synth.make_window(scene, n=n, P=P, seed=7, spacing=0.35, w=w, h=h, K=K)
    orc = O.Oracle(w, h, K, synth.BASELINE)
    ba, _, cw = ba_synth.fill_oracle(win, orc, OB.OracleBA, OB.immature_init)
    ctx = pkg.Context(w, h, K, synth.BASELINE)
    stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
    W, _ = ba_synth.fill_device(win, ctx, pkg.Window, cw)
    R = W.counts()["res"]
    def it_dev():
        W.linearize_all_async(True); W.assemble(); W.solve_assembled(2, want=False)
    for _ in range(5): it_dev()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.iters): it_dev()
    e1.record(stream); torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / a.iters
    reps = 3 if P > 5000 else 10
    ba.linearize_all(True); ba.solve(2)
    t0 = time.perf_counter()
    for _ in range(reps):
        ba.linearize_all(True); x, _, _ = ba.solve(2); ba.resubstitute(x)
    ms_cpu = 1e3 * (time.perf_counter() - t0) / reps
    # the fork's live LBA: FullSystem::optimize g2o body (E2 graph + restated g2o LM), wall time of 3 LM iterations, before the SSE
    # driver below moves the window's state
    st0 = ba.get_state()
    T_wh = np.stack([np.hstack([T[:, :3].T, (-T[:, :3].T @ T[:, 3])[:, None]]) for T in st0["T_w2c"]])
    rng = np.random.default_rng(0)
    Tp = np.stack([synth.perturb_T(T, rng, 3e-3, 3e-4) for T in T_wh])
    idp = np.array([float(p["idepth"]) for p in win["points"] for _ in p["targets"]])
    W.lba_g2o(np.array(K, float), Tp, np.zeros((n, 2)), idp, 3)
    t0 = time.perf_counter(); gg = W.lba_g2o(np.array(K, float), Tp, np.zeros((n, 2)), idp, 3); t_g2o_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); go = ba.lba_g2o(np.array(K, float), Tp, np.zeros((n, 2)), idp, 3); t_g2o_cpu = time.perf_counter() - t0
    t0 = time.perf_counter(); ro, io = ba.optimize(6); t_opt_cpu = time.perf_counter() - t0
    t0 = time.perf_counter(); rg, ig = W.optimize(6); t_opt_dev = time.perf_counter() - t0
    print(json.dumps(dict(config=name, frames=n, points=len(win["points"]), residuals=R, size=[w, h], ms_per_lm_iteration_device=ms_dev,
                          ms_per_lm_iteration_cpu_port_1thread=ms_cpu, speedup=ms_cpu / ms_dev, ba_evals_per_s_device=8 * R / (ms_dev * 1e-3),
                          optimize6_wall_ms_device=1e3 * t_opt_dev, optimize6_wall_ms_cpu_port=1e3 * t_opt_cpu, rmse_device=rg, rmse_cpu=ro, iterations=[ig, io],
                          lba_g2o_3its_wall_ms_device=1e3 * t_g2o_dev, lba_g2o_3its_wall_ms_cpu_port=1e3 * t_g2o_cpu,
                          lba_g2o_iterations=[gg["iterations"], go["iterations"]], lba_g2o_chi2=[gg["chi2"], go["chi2"]])), flush=True)
    ctx.close()

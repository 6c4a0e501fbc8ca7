"""GPU diagnostic: device-vs-oracle error of every BA stage next to the oracle's own spread under the reference's 6-worker
accumulation (oracle reduce_threads=6, several chunk->worker assignments). Prints one line per stage and configuration."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest, synth, ba_synth
import oracle_py as O, oracle_ba_py as OB
pkg = conftest.load_pkg()
scene = synth.make_scene()
CFG = {
    "small640": dict(n=4, P=240, seed=3, spacing=0.6, w=640, h=192, K=(360.0, 360.0, 319.5, 95.5)),
    "win7_640": dict(n=7, P=2002, seed=11, spacing=0.35, w=640, h=192, K=(360.0, 360.0, 319.5, 95.5)),
    "win7_1232": dict(n=7, P=2002, seed=11, spacing=0.35, w=synth.W, h=synth.H, K=synth.K4),
    "dense10_1920": dict(n=10, P=20000, seed=12, spacing=0.3, w=1920, h=1088, K=(1100.0, 1100.0, 959.5, 543.5)),
}
which = sys.argv[1:] or list(CFG)
for name in which:
    c = CFG[name]
    t0 = time.time()
    win = ba_synth.make_window(scene, n=c["n"], P=c["P"], seed=c["seed"], spacing=c["spacing"], w=c["w"], h=c["h"], K=c["K"])
    orc = O.Oracle(c["w"], c["h"], c["K"], synth.BASELINE)
    ba, _, cw = ba_synth.fill_oracle(win, orc, OB.OracleBA, OB.immature_init)
    ctx = pkg.Context(c["w"], c["h"], c["K"], synth.BASELINE)
    W, _ = ba_synth.fill_device(win, ctx, pkg.Window, cw)
    n = c["n"]; d = 4 + 8 * n
    print(f"== {name}: build {time.time()-t0:.1f}s  counts {W.counts()}", flush=True)
    Eo = ba.linearize_all(True); Eg = W.linearize_all(True)
    ro, rg = ba.get_res(1), W.get_res(1)
    print(f"   energy rel {abs(Eg-Eo)/Eo:.2e}  state equal {np.array_equal(ro['state'], rg['state'])}  active {int(ro['active'].sum())}")
    N = ba.nullspaces()
    A2 = np.zeros((d, 2)); A2[10::8, 0] = 1; A2[11::8, 1] = 1
    Qa, _ = np.linalg.qr(np.hstack([N / np.linalg.norm(N, axis=0), A2]))
    proj = lambda v: v - Qa @ (Qa.T @ v)
    def blockerr(a, b):
        glob = np.abs(b).max(); worst = 0
        for lo, hi in [(0, 4)] + [(4 + 8 * i, 12 + 8 * i) for i in range(n)]:
            s = max(np.abs(b[lo:hi]).max(), 1e-2 * glob)
            worst = max(worst, np.abs(a[lo:hi] - b[lo:hi]).max() / s)
        return worst
    for it in (0, 2):
        ba.set_reduce(1, 0)
        xo, Hfo, bfo = ba.solve(it); xg, Hfg, bfg = W.solve(it)
        ba.resubstitute(xo); so = ba.get_points()["step"].copy()
        W.resubstitute(xo); sg = W.get_points()["step"].copy()
        W.solve(it); W.resubstitute(None); sg_own = W.get_points()["step"].copy()
        sp_x, sp_xp, sp_s = 0, 0, 0
        for seed in range(4):
            ba.set_reduce(6, seed)
            x2, _, _ = ba.solve(it)
            ba.resubstitute(x2); s2 = ba.get_points()["step"].copy()
            sp_x = max(sp_x, blockerr(x2, xo)); sp_xp = max(sp_xp, blockerr(proj(x2), proj(xo)))
            sp_s = max(sp_s, np.abs(s2 - so).max() / np.abs(so).max())
        ba.set_reduce(1, 0)
        ba.solve(it); ba.resubstitute(xo)
        print(f"   it={it}: x raw dev {blockerr(xg, xo):.2e} (oracle spread {sp_x:.2e}) | x proj dev {blockerr(proj(xg), proj(xo)):.2e} (spread {sp_xp:.2e})"
              f" | step(given xo) dev max {np.abs(sg-so).max()/np.abs(so).max():.2e} | step(own x) dev {np.abs(sg_own-so).max()/np.abs(so).max():.2e} (spread {sp_s:.2e})"
              f" | H {np.abs(Hfg-Hfo).max()/np.abs(Hfo).max():.1e} b {np.abs(bfg-bfo).max()/np.abs(bfo).max():.1e}", flush=True)
    # optimize (both from the same fresh window)
    def fresh():
        o2 = O.Oracle(c["w"], c["h"], c["K"], synth.BASELINE)
        b2, _, cw2 = ba_synth.fill_oracle(win, o2, OB.OracleBA, OB.immature_init)
        return o2, b2
    win = ba_synth.make_window(scene, n=c["n"], P=c["P"], seed=c["seed"] + 10, spacing=c["spacing"] * 1.4, w=c["w"], h=c["h"], K=c["K"],
                               idepth_noise=0.03, state_sigma=3e-3)
    res = []
    for thr, seed in ((1, 0), (6, 0), (6, 1), (6, 2)):
        o2, b2 = fresh(); b2.set_reduce(thr, seed)
        t0 = time.time(); r, it = b2.optimize(6); dt = time.time() - t0
        res.append((r, it, b2.get_state(), b2.get_res(1)["active"].copy(), dt))
    ctx2 = pkg.Context(c["w"], c["h"], c["K"], synth.BASELINE)
    _, _, cw2 = ba_synth.fill_oracle(win, O.Oracle(c["w"], c["h"], c["K"], synth.BASELINE), OB.OracleBA, OB.immature_init)
    W2, _ = ba_synth.fill_device(win, ctx2, pkg.Window, cw2)
    rg, ig = W2.optimize(6); sg = W2.get_state(); ag = W2.get_res(1)["active"]
    r0, i0, s0, a0, dt0 = res[0]
    def cmp(s, a, r, it):
        rel = np.abs(s["idepth"] - s0["idepth"]) / np.abs(s0["idepth"])
        return (f"iters {it} rmse rel {abs(r-r0)/r0:.2e} pose max {np.abs(s['T_w2c']-s0['T_w2c']).max():.2e} idepth rel median {np.median(rel):.2e} "
                f"p99.5 {np.quantile(rel, 0.995):.2e} max {rel.max():.2e} frac<1e-4 {(rel<1e-4).mean():.4f} active agree {(a==a0).mean():.5f}")
    print(f"   optimize oracle(1 thread) {dt0*1e3:.0f} ms, iters {i0}")
    print("   optimize dev   :", cmp(sg, ag, rg, ig))
    for k in (1, 2, 3):
        print(f"   optimize orc6/{k-1}:", cmp(res[k][2], res[k][3], res[k][0], res[k][1]), flush=True)
    ctx.close(); ctx2.close()

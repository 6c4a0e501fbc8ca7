"""GPU diagnostic: sensitivity of the BA solve to float accumulation noise (oracle vs device)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest, synth, ba_synth
import oracle_py as O, oracle_ba_py as OB
import test_gpu_ba as T
pkg = conftest.load_pkg()
scene = synth.make_scene()
for (n, P, seed, sp) in ((4, 240, 3, 0.6), (7, 2002, 11, 0.35)):
    win, orc, ba, ctx, W = T.build(pkg, scene, n, P, seed, spacing=sp)
    ba.linearize_all(True); W.linearize_all(True)
    for it in (0, 2):
        xo, Hfo, bfo = ba.solve(it); xg, Hfg, bfg = W.solve(it)
        d = 4 + 8 * n
        def ref_solve(H, b):
            s = 1 / np.sqrt(np.diag(H) + 10)
            return s * np.linalg.solve(s[:, None] * H * s[None, :], s * b)
        N = ba.nullspaces(); Q, _ = np.linalg.qr(N / np.linalg.norm(N, axis=0))
        proj = lambda v: v - Q @ (Q.T @ v)
        xr_g, xr_o = ref_solve(Hfg, bfg), ref_solve(Hfo, bfo)
        if it >= 2: xr_g, xr_o = proj(xr_g), proj(xr_o)
        print(f"n={n} it={it}: |xg-xo|/|xo|={np.linalg.norm(xg-xo)/np.linalg.norm(xo):.2e}  proj: {np.linalg.norm(proj(xg)-proj(xo))/np.linalg.norm(proj(xo)):.2e}"
              f"  solver self: gpu {np.linalg.norm(xg-xr_g)/np.linalg.norm(xr_g):.2e} oracle {np.linalg.norm(xo-xr_o)/np.linalg.norm(xr_o):.2e}"
              f"  cond={np.linalg.cond((1/np.sqrt(np.diag(Hfo)+10))[:,None]*Hfo*(1/np.sqrt(np.diag(Hfo)+10))[None,:]):.2e}"
              f"  |dH|/|H|={np.abs(Hfg-Hfo).max()/np.abs(Hfo).max():.2e} |db|/|b|={np.abs(bfg-bfo).max()/np.abs(bfo).max():.2e}")
        for lo, hi in [(0, 4)] + [(4 + 8 * i, 12 + 8 * i) for i in range(n)]:
            print("   block", lo, " max|dx|/max|x| =", f"{np.abs(xg[lo:hi]-xo[lo:hi]).max()/np.abs(xo[lo:hi]).max():.2e}", " proj:", f"{np.abs(proj(xg)[lo:hi]-proj(xo)[lo:hi]).max()/np.abs(proj(xo)[lo:hi]).max():.2e}")
    ctx.close()

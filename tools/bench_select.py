"""Times the device pixel selector (sdso_make_maps) against the CPU restatement on one 640x480 frame.
Usage: python tools/bench_select.py [--reps 50]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import oracle_py as O
import oracle_select_py as S
import synth

ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=50); ap.add_argument("--w", type=int, default=640); ap.add_argument("--h", type=int, default=480)
a = ap.parse_args()
pkg = G._load_pkg()
w, h = a.w, a.h
K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
rng = np.random.default_rng(0)
img, _ = synth.render(synth.make_scene(), synth.camera_pose(0), w, h, K)
img_u8 = np.kron(rng.integers(0, 255, (h // 4, w // 4)), np.ones((4, 4))).astype(np.float32) + rng.integers(0, 3, (h, w)).astype(np.float32)
out = {}
for name, im in (("scene", img), ("u8_blocks", img_u8)):
    ctx = pkg.Context(w, h, K, 0.1); orc = O.Oracle(w, h, K, 0.1)
    g, o = ctx.frame_create(), orc.frame_new()
    ctx.make_images(g, im); orc.make_images(o, im)
    sel = S.Selector(orc)
    for pot in (1, 3):
        ctx.selector_make_hists(g); sel.make_hists(o)
        ctx.selector_select(g, pot, 1.0, want_map=False)
        t0 = time.perf_counter()
        for _ in range(a.reps):
            _, n = ctx.selector_select(g, pot, 1.0, want_map=False)
        tg = (time.perf_counter() - t0) / a.reps
        t0 = time.perf_counter()
        for _ in range(max(3, a.reps // 10)):
            _, no = sel.select(o, pot, 1.0)
        tc = (time.perf_counter() - t0) / max(3, a.reps // 10)
        out[f"{name}_select_pot{pot}"] = {"gpu_ms_incl_sync": tg * 1e3, "cpu_ms": tc * 1e3, "n": n.tolist(), "match": bool((n == no).all())}
    ctx.selector_reset(); 
    ctx.make_maps(g, 3000.0, want_map=False)
    t0 = time.perf_counter()
    for _ in range(a.reps):
        ctx.selector_reset()
        ctx.selector_make_hists(g)
        _, n = ctx.make_maps(g, 3000.0, want_map=False)
    tg = (time.perf_counter() - t0) / a.reps
    t0 = time.perf_counter()
    for _ in range(max(3, a.reps // 10)):
        sel.potential(3); sel.forget_hist()
        _, no = sel.make_maps(o, 3000.0)
    tc = (time.perf_counter() - t0) / max(3, a.reps // 10)
    out[f"{name}_make_maps"] = {"gpu_ms_incl_sync": tg * 1e3, "cpu_ms": tc * 1e3, "n": n, "match": n == no}
    ctx.close()
import json
print(json.dumps(out, indent=1))

"""Times sdso_distmap_make + sdso_activation_filter against the CPU restatement (640x480, 7 hosts).
Usage: python tools/bench_distmap.py [--reps 20]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G
import oracle_py as O
import oracle_distmap_py as D

ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
pkg = G._load_pkg()
w, h = 640, 480
K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
ctx, orc = pkg.Context(w, h, K, 0.1), O.Oracle(w, h, K, 0.1)
dm = D.DistMap(orc)
out = {}
for name, n_pts, n_cand, mad in (("kf_2000pts_9000cand_mad2", 2000, 9000, 2.0), ("sparse_200pts_9000cand_mad1", 200, 9000, 1.0), ("dense_4000pts_20000cand_mad3", 4000, 20000, 3.0)):
    inp = D.make_inputs(orc, 1, n_hosts=6, n_pts=n_pts, n_cand=n_cand)
    args_m = (inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    args_f = (inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
    ctx.distmap_make(*args_m); ctx.activation_filter(*args_f)
    tm = tf = 0.0
    for _ in range(a.reps):
        t0 = time.perf_counter(); ctx.distmap_make(*args_m, want_map=False); t1 = time.perf_counter(); vg, _, rounds = ctx.activation_filter(*args_f, want_map=False); t2 = time.perf_counter()
        tm += t1 - t0; tf += t2 - t1
    cm = cf = 0.0
    for _ in range(3):
        t0 = time.perf_counter(); dm.make(*args_m); t1 = time.perf_counter(); vo, _ = dm.filter(*args_f); t2 = time.perf_counter()
        cm += t1 - t0; cf += t2 - t1
    out[name] = {"gpu_make_ms": tm / a.reps * 1e3, "gpu_filter_ms": tf / a.reps * 1e3, "cpu_make_ms": cm / 3 * 1e3, "cpu_filter_ms": cf / 3 * 1e3,
                 "rounds": rounds, "accepted": int((vg == 1).sum()), "match": bool((vg == vo).all())}
print(json.dumps(out, indent=1))

"""GPU experiment: batched track-kernel time and per-sequence work as a function of the distance new frame <-> keyframe."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, synth
pkg = bench.load_pkg()
S, SMAX = 296, 37
wl = bench.build_workload(seqs=SMAX)
s = pkg.default_settings(); s.cluster_size = 1; s.gather_batch = 1
ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, settings=s)
for k in range(S):
    seq = wl[k % SMAX]
    ctx.tracker_select_ref(k)
    fr = ctx.frame_create(); ctx.make_images(fr, seq["ref_img"]); ctx.tracker_set_ref(fr, seq["pts"])
fn = [ctx.frame_create() for _ in range(S)]
for j in range(bench.POSES):
    for k in range(S): ctx.make_images(fn[k], wl[k % SMAX]["new_imgs"][j])
    T = np.stack([wl[k % SMAX]["T_init"][j].reshape(12) for k in range(S)])
    for rep in range(2):
        ctx.track_enqueue_multi(list(range(S)), fn, T, np.zeros((S, 2)), ctx.levels - 1, np.full((S, 5), np.nan), 0); r = ctx.track_collect(S)
    ctx.profile_enable(True)
    for rep in range(5):
        ctx.track_enqueue_multi(list(range(S)), fn, T, np.zeros((S, 2)), ctx.levels - 1, np.full((S, 5), np.nan), 0); r = ctx.track_collect(S)
    p = ctx.profile_read(); ctx.profile_enable(False)
    its = r["iterations"].sum(1)
    print(f"pose {j} ({0.25*(j+1):.2f} m): {1e3*p['track_ms']/p['track_launches']:8.1f} us/launch  evals/seq={r['evals']/S:8.0f}  LM iterations per sequence: mean {its.mean():.1f} max {its.max()} min {its.min()}  ok={int(r['ok'].sum())}", flush=True)
ctx.close()

"""GPU experiment: track-kernel time for S independent sequences per launch over (S, cluster size, threads)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, synth
pkg = bench.load_pkg()
SMAX = 36
wl = bench.build_workload(seqs=SMAX)
for C_, BT, U in ((1, 256, 1), (1, 256, 2), (2, 256, 1), (1, 128, 1)):
    s = pkg.default_settings(); s.cluster_size = C_; s.block_threads = BT; s.gather_batch = U
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, settings=s)
    fn = []
    for k, seq in enumerate(wl):
        ctx.tracker_select_ref(k)
        fr = ctx.frame_create(); ctx.make_images(fr, seq["ref_img"]); ctx.tracker_set_ref(fr, seq["pts"])
        f = ctx.frame_create(); ctx.make_images(f, seq["new_imgs"][0]); fn.append(f)
    for S in (74, 148, 296, 320):
        if S * C_ > 320 * (2 if BT == 128 else 1): continue
        T = np.stack([wl[k % SMAX]["T_init"][0].reshape(12) for k in range(S)])
        for rep in range(3):
            ctx.track_enqueue_multi([k % SMAX for k in range(S)], [fn[k % SMAX] for k in range(S)], T, np.zeros((S, 2)), ctx.levels - 1, np.full((S, 5), np.nan), 0); r = ctx.track_collect(S)
        ctx.profile_enable(True)
        ev = 0
        for rep in range(10):
            ctx.track_enqueue_multi([k % SMAX for k in range(S)], [fn[k % SMAX] for k in range(S)], T, np.zeros((S, 2)), ctx.levels - 1, np.full((S, 5), np.nan), 0); ev += ctx.track_collect(S)["evals"]
        p = ctx.profile_read(); ctx.profile_enable(False)
        us = 1e3 * p["track_ms"] / p["track_launches"]
        print(f"C={C_:2d} BT={BT} U={U} S={S:2d}: {us:8.1f} us/launch {us/S:7.1f} us/seq  {ev/10*64/us/1e3:7.1f} GB/s algorithmic  ok={int(r['ok'].sum())}", flush=True)
    ctx.close()

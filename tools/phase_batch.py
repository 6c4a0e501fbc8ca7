"""GPU experiment: phase breakdown (clock64 of rank 0 / thread 0, averaged over the problems) of the batched track kernel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, synth
pkg = bench.load_pkg()
SMAX = 37
wl = bench.build_workload(seqs=SMAX)
for C_, BT, U, S in ((2, 256, 1, 592), (1, 256, 1, 592), (8, 256, 2, 1)):
    s = pkg.default_settings(); s.cluster_size = C_; s.block_threads = BT; s.gather_batch = U
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, settings=s)
    fn = []
    for k in range(S):  # distinct pyramids / templates per problem, as in bench.py
        seq = wl[k % SMAX]
        ctx.tracker_select_ref(k)
        fr = ctx.frame_create(); ctx.make_images(fr, seq["ref_img"]); ctx.tracker_set_ref(fr, seq["pts"])
        f = ctx.frame_create(); ctx.make_images(f, seq["new_imgs"][0]); fn.append(f)
    T = np.stack([wl[k % SMAX]["T_init"][0].reshape(12) for k in range(S)])
    for rep in range(3):
        ctx.track_enqueue_multi(list(range(S)), fn, T, np.zeros((S, 2)), ctx.levels - 1, np.full((S, 5), np.nan), 0); r = ctx.track_collect(S)
    ctx.profile_enable(True)
    cyc = np.zeros(16); ev = 0
    for rep in range(5):
        ctx.track_enqueue_multi(list(range(S)), fn, T, np.zeros((S, 2)), ctx.levels - 1, np.full((S, 5), np.nan), 0); r = ctx.track_collect(S); ev += r["evals"]
        cyc += np.array(ctx.track_phase_cycles())
    p = ctx.profile_read(); ctx.profile_enable(False)
    us = 1e3 * p["track_ms"] / p["track_launches"]
    print(f"C={C_} BT={BT} U={U} S={S}: {us:8.1f} us/launch; its={r['iterations'][0]} evals/seq={ev/5/S:.0f}; cycles per problem: "
          f"prologue[11]={cyc[11]/5/S:.0f} sync[0]={cyc[0]/5/S:.0f} eval[1]={cyc[1]/5/S:.0f} reduce[2]={cyc[2]/5/S:.0f} exch[3]={cyc[3]/5/S:.0f} "
          f"book[4]={cyc[4]/5/S:.0f} [5]={cyc[5]/5/S:.0f} sub[6..10]={np.round(cyc[6:11]/5/S)}", flush=True)
    ctx.close()

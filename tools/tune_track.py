"""GPU experiment: track-kernel time and phase breakdown over (cluster size, threads per CTA).
Run on the GPU box: python tools/tune_track.py [sse|g2o]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, synth

variant = 1 if (len(sys.argv) > 1 and sys.argv[1] == "g2o") else 0
pkg = bench.load_pkg()
wl = bench.build_workload()
configs = [(8, 256, 1), (8, 256, 2), (16, 256, 1), (16, 256, 2), (16, 128, 1)]
if variant == 1: configs = [(8, 256, 1), (16, 256, 1), (4, 256, 1)]
for C_, BT, U in configs:
    s = pkg.default_settings(); s.cluster_size = C_; s.block_threads = BT; s.gather_batch = U
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, settings=s)
    fref = ctx.frame_create(); ctx.make_images(fref, wl["ref_img"]); ctx.tracker_set_ref(fref, wl["pts"])
    fn = [ctx.frame_create() for _ in range(4)]
    for j, f in enumerate(fn): ctx.make_images(f, wl["new_imgs"][j])
    mr = np.full((1, 5), np.nan)
    for j in range(4):
        ctx.track_enqueue([fn[j]], wl["T_init"][j].reshape(1, 12), np.zeros((1, 2)), ctx.levels - 1, mr, variant); ctx.track_collect(1)
    ctx.profile_enable(True)
    ev = 0; its = None; cyc = np.zeros(16)
    for rep in range(5):
        for j in range(4):
            ctx.track_enqueue([fn[j]], wl["T_init"][j].reshape(1, 12), np.zeros((1, 2)), ctx.levels - 1, mr, variant)
            r = ctx.track_collect(1); ev += r["evals"]; its = r["iterations"][0]
            cyc += np.array(ctx.track_phase_cycles())
    p = ctx.profile_read()
    n = p["track_launches"]
    print(f"C={C_:2d} BT={BT:3d} U={U}: {1e3*p['track_ms']/n:8.1f} us/launch  evals/launch={ev/n:9.0f}  its={its}  phase cycles/launch={np.round(cyc[:12]/n).astype(int)}", flush=True)
    # batched hypotheses
    for nb in (4, 9):
        if C_ * nb > 148: continue
        ctx.profile_enable(True)
        for rep in range(5):
            ctx.track_enqueue([fn[j % 4] for j in range(nb)], np.stack([wl["T_init"][j % 4].reshape(12) for j in range(nb)]), np.zeros((nb, 2)), ctx.levels - 1, np.full((nb, 5), np.nan), variant)
            r = ctx.track_collect(nb)
        p = ctx.profile_read()
        print(f"      batch nb={nb}: {1e3*p['track_ms']/p['track_launches']:8.1f} us/launch ({1e3*p['track_ms']/p['track_launches']/nb:6.1f} us/problem)", flush=True)
    ctx.close()

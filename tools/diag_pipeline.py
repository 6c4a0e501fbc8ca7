"""GPU diagnostic: oracle and device pipelines side by side, per-frame pose difference and key-frame bookkeeping."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest, synth, pipeline as PL
pkg = conftest.load_pkg()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 45
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
shape = dict(w=synth.W, h=synth.H, K=synth.K4)
scene = synth.make_scene()
poses = [synth.camera_pose(0.2 * k) for k in range(N)]
L = [synth.render_torch(scene, p)[0] for p in poses]; R = [synth.render_torch(scene, synth.right_of(p))[0] for p in poses]
Bo, Bg = PL.Backend(shape), PL.Backend(shape, pkg)
Po, Pg = PL.StereoPipeline(Bo, variant=variant), PL.StereoPipeline(Bg, variant=variant)
import copy
store = {}
def hook(win, hm):
    store.update(win=copy.deepcopy(win), hm=hm, kf_idx=[kf["frame_index"] for kf in Po.kfs])
Po.on_window = hook
for k in range(N):
    ro, rg = Po.step(L[k], R[k]), Pg.step(L[k], R[k])
    if k % 5 == 0 and k > 0 and "win" in store:
        # teacher-forced: the ORACLE pipeline's window (before its optimisation) optimised by the device, on the device's own key-frame pyramids
        # (same images, same key-frame order), in a second context so the device pipeline's state is untouched
        win = store["win"]
        if "B2" not in store:
            store["B2"] = PL.Backend(shape, pkg)
        B2 = store["B2"]
        fids2 = [B2.new_frame(L[kf["pose_idx"]]) if "pose_idx" in kf else None for kf in Po.kfs] if False else None
        imgs_idx = store.setdefault("kf_frames", [])
        B2f = [B2.new_frame(L[i]) for i in store["kf_idx"]]
        W2 = B2.window(win, B2f)
        HM, bM = store["hm"]
        if HM is not None:
            d_ = 4 + 8 * win["n"]; H2 = np.zeros((d_, d_)); b2 = np.zeros(d_); m = HM.shape[0]; H2[:m, :m] = HM; b2[:m] = bM
            W2.set_marg_prior(H2, b2)
        r2, i2 = W2.optimize(6)
        for f in B2f: B2.release(f)
        print(f"   teacher-forced device optimize of the oracle's window: rmse {r2:.6f} its {i2}  (oracle: {Po.log[-1]['rmse']:.6f} its {Po.log[-1]['iterations']})")
    dt = np.abs(Po.traj[k][:3, 3] - Pg.traj[k][:3, 3]).max()
    print(k, ro["ok"], rg["ok"], f"dt {dt:.2e}", "res", np.round(ro.get("lastResiduals", [0])[:2], 3), np.round(rg.get("lastResiduals", [0])[:2], 3),
          "aff", np.round(Po.aff, 4), np.round(Pg.aff, 4), flush=True)
    if k % 5 == 0:
        print("   O", Po.log[-1]); print("   G", Pg.log[-1], flush=True)

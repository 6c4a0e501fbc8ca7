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
B3 = PL.Backend(shape, pkg)
tf = {}
def on_ref(idx, splat, aff):
    if "fid" in tf: B3.release(tf["fid"])
    tf["fid"] = B3.new_frame(L[idx]); tf["splat"] = splat.copy(); tf["aff"] = aff; tf["idx"] = idx
    B3.tracker_set_ref(tf["fid"], splat, aff)
def on_track(k, T_guess, aff_guess, r):
    f = B3.new_frame(L[k])
    g = B3.track(f, T_guess, aff_guess, variant)
    B3.release(f)
    dT = np.abs(g["T"] - r["T"]).max()
    flag = "  <<<<<<" if dT > 1e-5 else ""
    print(f"   teacher-forced device track of frame {k} vs ref {tf['idx']}: |dT| {dT:.2e} ok {g['ok']}/{r['ok']} its {g['iterations']} / {r.get('iterations')} res0 {g['lastResiduals'][0]:.4f}/{r['lastResiduals'][0]:.4f}{flag}")
    if dT > 1e-4 and "dumped" not in tf:
        tf["dumped"] = True
        np.savez("gpurun_out/r2_track_case.npz", ref_img=L[tf["idx"]], new_img=L[k], splat=tf["splat"], ref_aff=np.array(tf["aff"]), T_guess=T_guess, aff_guess=np.array(aff_guess),
                 T_oracle=r["T"], T_device=g["T"])
Po.on_ref = on_ref; Po.on_track = on_track
gs = {}
def on_ref_g(idx, splat, aff):
    o = tf["splat"]
    ko = {(int(a[0]), int(a[1])): a for a in o}; kg = {(int(a[0]), int(a[1])): a for a in splat}
    common = [k for k in ko if k in kg]
    if common:
        A = np.array([ko[k] for k in common]); G = np.array([kg[k] for k in common])
        rid = np.abs(G[:, 2] - A[:, 2]) / np.abs(A[:, 2]); rw = np.abs(G[:, 3] - A[:, 3]) / np.abs(A[:, 3])
        print(f"   ref @ frame {idx}: oracle {len(o)} splats, device {len(splat)}, common pixels {len(common)}; idepth rel diff median {np.median(rid):.2e} p99 {np.quantile(rid, 0.99):.2e} max {rid.max():.2e}; "
              f"weight rel diff median {np.median(rw):.2e} max {rw.max():.2e}; mean idepth ratio {np.mean(G[:, 2] / A[:, 2]):.6f}; aff o {np.round(tf['aff'], 4)} g {np.round(aff, 4)}")
Pg.on_ref = on_ref_g
B4 = PL.Backend(shape)   # oracle tracker with the oracle pipeline's reference
def on_ref4(idx, splat, aff):
    if "fid4" in tf: B4.release(tf["fid4"])
    tf["fid4"] = B4.new_frame(L[idx]); B4.tracker_set_ref(tf["fid4"], splat, aff)
_old_on_ref = Po.on_ref
def on_ref_both(idx, splat, aff):
    _old_on_ref(idx, splat, aff); on_ref4(idx, splat, aff)
Po.on_ref = on_ref_both
def on_track_g(k, T_guess, aff_guess, r):
    # the DEVICE pipeline's initial guess against the ORACLE pipeline's reference: device tracker vs oracle tracker
    f3, f4 = B3.new_frame(L[k]), B4.new_frame(L[k])
    g = B3.track(f3, T_guess, aff_guess, variant); o = B4.track(f4, T_guess, aff_guess, variant)
    B3.release(f3); B4.release(f4)
    dT = np.abs(g["T"] - o["T"]).max()
    print(f"   device-guess cross-check frame {k}: device vs oracle tracker |dT| {dT:.2e} its {g['iterations']} / {o.get('iterations')} res0 {g['lastResiduals'][0]:.4f}/{o['lastResiduals'][0]:.4f}; "
          f"free-running device result vs this: |dT| {np.abs(r['T'] - g['T']).max():.2e}{'  <<<<<<' if dT > 1e-5 else ''}")
Pg.on_track = on_track_g
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

"""A host-thin stereo VO pipeline over the operators of the hot path — the harness for SURVEY config 1 (200-frame KITTI-shape
sequence, tracking + mapping chained). It is NOT a FullSystem rewrite (key-frame decisions are "every k-th frame", the pointer
graph is a list of dicts): it strings the operators together in the order the reference does —

  per frame   (FullSystem::addActiveFrame, FullSystem.cpp:1058-1165): makeImages of the left AND the right image (:1083-1085),
              trackNewestCoarse against the newest key frame from a constant-velocity guess, traceOn of every key frame's
              immature points into the frame (traceNewCoarse, :745-781)
  key frame   (FullSystem::makeKeyFrame, :1331-1483): flag the oldest key frame when the window is full, residuals of all active
              points towards the new key frame, distance map + candidate loop + optimizeImmaturePoint (activatePointsMT,
              :796-958), the windowed optimisation (SSE body, 6 iterations), removal of points without residuals, the tracker
              reference from the window's points (setCoarseTrackingRef), marginalizePointsF for the points of the flagged frame,
              makeMaps + ImmaturePoint constructor + traceStereo into the right image for the new candidates (makeNewTraces,
              :1599-1630), marginalizeFrame

— once per backend. Every stage consumes what the SAME backend produced before; the oracle and the device are never
re-synchronised, so per-frame trajectory agreement is a statement about the whole chain."""
import time
import numpy as np
import oracle_py as O
import oracle_ba_py as OB
import oracle_trace_py as OT
import oracle_distmap_py as OD
import oracle_select_py as OS
import synth
import trace_synth as TS

SCALE_A, SCALE_B = 10.0, 1000.0


def on_so3(R):
    """nearest rotation (the reference's SE3 is a unit quaternion: products renormalise, so3.hpp:215-232). Without this the R^T-as-
    inverse below feeds the round-off non-orthonormality of one frame's pose into the constant-velocity guess of the next and it
    grows by ~2.8x per frame."""
    U, _, Vt = np.linalg.svd(R)
    return U @ Vt


def inv34(T):
    R, t = on_so3(T[:, :3]), T[:, 3]
    return np.hstack([R.T, (-R.T @ t)[:, None]])


def mul34(A, B):
    return np.hstack([on_so3(A[:, :3] @ B[:, :3]), (A[:, :3] @ B[:, 3] + A[:, 3])[:, None]])


class Backend:
    """the same operator names for the oracle (test infrastructure) and the device library"""

    def __init__(self, shape, pkg=None, reduce=(1, 0)):
        self.dev = pkg is not None
        self.pkg = pkg
        self.shape = shape
        self.reduce = reduce   # oracle only: worker partition of the float accumulators (OracleBA.set_reduce)
        W_, H_, K_ = shape["w"], shape["h"], shape["K"]
        self.api = pkg.Context(W_, H_, K_, synth.BASELINE) if self.dev else O.Oracle(W_, H_, K_, synth.BASELINE)
        self.dm = None if self.dev else OD.DistMap(self.api)
        self.sel = None if self.dev else OS.Selector(self.api)
        self.free = []
        self.t_ops = 0.0   # seconds spent inside operator calls (ABI / oracle entry points incl. argument marshalling): what a C++ caller pays
        self.t_by_op = {}  # the same, per operator name

    def close(self):
        if self.dev:
            self.api.close()

    def _op(self, fn, *a, **kw):
        t0 = time.perf_counter()
        r = fn(*a, **kw)
        self._account(getattr(fn, "__name__", "op"), time.perf_counter() - t0)
        return r

    def _account(self, name, dt):
        self.t_ops += dt
        self.t_by_op[name] = self.t_by_op.get(name, 0.0) + dt

    # -- frames
    def new_frame(self, img):
        if self.dev:
            fid = self.api.frame_create()
        else:
            fid = self.free.pop() if self.free else self.api.frame_new()
        self._op(self.api.make_images, fid, img)
        return fid

    def release(self, fid):
        if self.dev:
            self.api.frame_release(fid)
        else:
            self.free.append(fid)   # the oracle keeps a growing list; slots are recycled here

    def frames(self, win):
        self.fids = [self.new_frame(f["image"]) for f in win["frames"]]

    # -- tracking
    def tracker_set_ref(self, fid, uvidw, aff):
        self._op(self.api.tracker_set_ref, fid, uvidw, aff)

    def track(self, fid, T, aff, variant):
        return self._op(self.api.track, fid, T, aff, self.api.levels - 1, [np.nan] * 5, variant)

    # -- immature points
    def immature_init(self, h, uv):
        return self.immature_init_fid(self.fids[h], uv)

    def immature_init_fid(self, fid, uv):
        return self._op(self.api.immature_init, fid, uv) if self.dev else self._op(OT.immature_init, self.api, fid, uv)

    def trace_on(self, t, KRKi, Kt, pts):
        return self.trace_on_fid(self.fids[t], KRKi, Kt, (1.0, 0.0), pts)

    def trace_on_fid(self, fid, KRKi, Kt, aff, pts):
        return self._op(self.api.trace_on, fid, KRKi, Kt, aff, pts) if self.dev else self._op(OT.trace_on, self.api, fid, KRKi, Kt, aff, pts)

    def trace_stereo(self, fid, K33, mode_right, pts):
        return self._op(self.api.trace_stereo, fid, K33, mode_right, pts) if self.dev else self._op(OT.trace_stereo, self.api, fid, K33, mode_right, pts)

    def make_maps(self, fid, density):
        """PixelSelector::makeMaps + the selectionMap walk of makeNewTraces: (uv [n,2], type [n]) in raster order"""
        if self.dev:
            self._op(self.api.make_maps, fid, density, want_map=False)
            uv, ty = self._op(self.api.selector_points)
            return uv, ty
        self.sel.forget_hist()   # (frame slots are recycled: never trust the selector's per-frame cache)
        m, _ = self._op(self.sel.make_maps, fid, density)
        ys, xs = np.nonzero(m)
        return np.stack([xs, ys], 1).astype(np.float32), m[ys, xs].astype(np.float32)

    # -- window
    def window(self, win, fids=None):
        """(re)build the backend window from the neutral description; colour / weights come from this backend's D1 operator"""
        fids = self.fids if fids is None else fids
        pts = win["points"]
        P = len(pts)
        host = np.array([p["host"] for p in pts], np.int32)
        uv = np.array([[p["u"], p["v"]] for p in pts], np.float32).reshape(-1, 2)
        col, wts = np.zeros((P, 8), np.float32), np.zeros((P, 8), np.float32)
        if P and all("color" in p for p in pts):
            # PointHessian keeps the colour / weights of the ImmaturePoint it was made from (HessianBlocks.cpp:36-53)
            col[:] = np.stack([p["color"] for p in pts]); wts[:] = np.stack([p["weights"] for p in pts])
        else:
            for h in range(win["n"]):
                idx = np.nonzero(host == h)[0]
                if idx.size:
                    rec, _ = self.immature_init_fid(fids[h], uv[idx])
                    col[idx] = rec["color"]; wts[idx] = rec["weights"]
        idp = np.array([p["idepth"] for p in pts], np.float32); idz = np.array([p["idepth_zero"] for p in pts], np.float32)
        prior = np.array([p["has_prior"] for p in pts], np.uint8)
        counts = [len(p["targets"]) for p in pts]
        rp = np.repeat(np.arange(P, dtype=np.int32), counts)
        rt = np.array([t for p in pts for t in p["targets"]], np.int32)
        t0 = time.perf_counter()
        Wn = self.pkg.Window(self.api) if self.dev else OB.OracleBA(self.api)
        for k, f in enumerate(win["frames"]):
            i = Wn.add_frame(fids[k], f["T_w2c"], f["a"], f["b"], f["frameID"])
            Wn.set_state(i, f["state"]); Wn.set_energy_th(i, f["energyTH"])
        t1 = time.perf_counter(); self._account("window_frames", t1 - t0)
        Wn.set_points(host, uv[:, 0].copy(), uv[:, 1].copy(), idp, idz, col, wts, prior)
        t2 = time.perf_counter(); self._account("window_set_points", t2 - t1)
        Wn.set_residuals(rp, rt)
        t3 = time.perf_counter(); self._account("window_set_residuals", t3 - t2)
        Wn.prepare()
        self._account("window_prepare", time.perf_counter() - t3)
        if not self.dev:
            Wn.set_reduce(*self.reduce)
        self.W = Wn
        return Wn

    def distmap_make(self, KRKi, Kt, pt_host, pt_uvid):
        return self._op(self.api.distmap_make, KRKi, Kt, pt_host, pt_uvid) if self.dev else self._op(self.dm.make, KRKi, Kt, pt_host, pt_uvid)

    def activation_filter(self, KRKi, Kt, flagged, cand_host, pts, my_type, mad):
        if self.dev:
            v, m, _ = self._op(self.api.activation_filter, KRKi, Kt, flagged, cand_host, pts, my_type, mad)
            return v, m
        return self._op(self.dm.filter, KRKi, Kt, flagged, cand_host, pts, my_type, mad)

    def activate(self, n, host, pts):
        return self._op(self.W.activate_points, host, pts, variant=0) if self.dev else self._op(OT.activate_points, self.api, n, host, pts, variant=0)

    def set_point_flags(self, flags):
        t0 = time.perf_counter()
        if self.dev:
            self.W.set_point_flags(flags)
        else:
            for i, f in enumerate(flags):
                if f:
                    self.W.set_point_flag(i, int(f))
        self._account("set_point_flags", time.perf_counter() - t0)


def level1_krki_kt(T_host_to_new, K4):
    """K[1] * R * Ki[0], K[1] * t (CoarseTracker.cpp:1233-1235 / FullSystem.cpp:845-847), float"""
    K0 = TS.K33(K4).astype(np.float32)
    K1 = K0.copy(); K1[0, 0] *= 0.5; K1[1, 1] *= 0.5; K1[0, 2] = (K0[0, 2] + 0.5) / 2 - 0.5; K1[1, 2] = (K0[1, 2] + 0.5) / 2 - 0.5
    R, t = T_host_to_new[:, :3].astype(np.float32), T_host_to_new[:, 3].astype(np.float32)
    return ((K1 @ R) @ np.linalg.inv(K0).astype(np.float32)).astype(np.float32).reshape(9), (K1 @ t).astype(np.float32)


def level0_krki_kt(T_host_to_new, K4):
    """hostToFrame_KRKi / hostToFrame_Kt of traceNewCoarse (FullSystem.cpp:760-764), float"""
    K = TS.K33(K4).astype(np.float64)
    KRKi = (K @ T_host_to_new[:, :3] @ np.linalg.inv(K)).astype(np.float32)
    return KRKi, (K @ T_host_to_new[:, 3]).astype(np.float32)


class StereoPipeline:
    def __init__(self, backend, kf_every=5, max_kf=7, immature_density=1500.0, point_density=2000.0, variant=0, opt_its=6):
        self.B = backend
        self.K4 = backend.shape["K"]
        self.kf_every, self.max_kf, self.variant, self.opt_its = kf_every, max_kf, variant, opt_its
        self.immature_density, self.point_density = immature_density, point_density
        self.kfs = []          # key frames of the window, oldest first
        self.points = []       # active points: dict(host=frameID, u, v, idepth, idepth_zero, has_prior, targets=[frameID...])
        self.HM = self.bM = None
        self.next_frame_id = 0
        self.min_act_dist = 2.0
        self.traj = []         # camToWorld 4x4 per frame
        self.T_w2c_hist = []   # worldToCam 3x4 per frame
        self.aff = (0.0, 0.0)
        self.log = []          # per key frame: counts for the report / for comparing two backends
        self.on_window = None  # diagnostics: called with (window description, marg prior) right before the windowed optimisation
        self.on_track = None   # diagnostics: called with (frame index, T_guess, aff_guess, result) after every tracked frame
        self.on_ref = None     # diagnostics: called with (frame index of the key frame, splats, aff) whenever the tracker reference is set

    # ---------------------------------------------------------------------------------------------------------------
    def step(self, img_left, img_right):
        B = self.B
        fl, fr = B.new_frame(img_left), B.new_frame(img_right)   # both pyramids of the stereo frame
        k = len(self.traj)
        if k == 0:
            T_w2c = np.eye(4)[:3]
            self._record(T_w2c)
            self._make_keyframe(fl, fr, T_w2c, first=True)
            return dict(ok=True)
        ref = self.kfs[-1]
        T_ref = self._pose_of(ref)
        if len(self.T_w2c_hist) >= 2:   # constant velocity in the camera frame
            dT = mul34(self.T_w2c_hist[-1], inv34(self.T_w2c_hist[-2]))
            guess_w2c = mul34(dT, self.T_w2c_hist[-1])
        else:
            guess_w2c = self.T_w2c_hist[-1]
        T_guess, aff_guess = mul34(guess_w2c, inv34(T_ref)), self.aff
        r = B.track(fl, T_guess, aff_guess, self.variant)
        if self.on_track:
            self.on_track(k, T_guess, aff_guess, r)
        T_w2c = mul34(r["T"], T_ref)
        self.aff = tuple(float(x) for x in r["aff"])
        self._record(T_w2c)
        self._trace_immature(fl, T_w2c)
        if k % self.kf_every == 0:
            self._make_keyframe(fl, fr, T_w2c, first=False)
        else:
            B.release(fl); B.release(fr)
        return dict(ok=bool(r["ok"]), lastResiduals=r["lastResiduals"])

    def _record(self, T_w2c):
        self.T_w2c_hist.append(T_w2c.copy())
        self.traj.append(np.vstack([inv34(T_w2c), [0, 0, 0, 1]]))

    def _pose_of(self, kf):
        return kf["T_cur"]

    def _trace_immature(self, fid, T_w2c):
        """traceNewCoarse: every key frame's immature points into the new frame (the device backend: ONE launch for all hosts)"""
        jobs = []
        for kf in self.kfs:
            pts = kf["immature"]
            if pts is None or pts.size == 0:
                continue
            T_h2f = mul34(T_w2c, inv34(self._pose_of(kf)))
            KRKi, Kt = level0_krki_kt(T_h2f, self.K4)
            a = float(np.exp(self.aff[0] - kf["aff_cur"][0])); b = float(self.aff[1] - a * kf["aff_cur"][1])   # AffLight::fromToVecExposure, exposures 1
            jobs.append((kf, KRKi, Kt, (np.float32(a), np.float32(b))))
        if not jobs:
            return
        if self.B.dev:
            allp = np.ascontiguousarray(np.concatenate([j[0]["immature"] for j in jobs]))
            host_of = np.concatenate([np.full(j[0]["immature"].size, i, np.int32) for i, j in enumerate(jobs)])
            self.B._op(self.B.api.trace_on_hosts, fid, np.stack([j[1] for j in jobs]), np.stack([j[2] for j in jobs]), np.array([j[3] for j in jobs], np.float32), host_of, allp,
                       want_status=False)
            off = 0
            for kf, _, _, _ in jobs:
                m = kf["immature"].size
                kf["immature"] = np.ascontiguousarray(allp[off:off + m]); off += m
        else:
            for kf, KRKi, Kt, aff in jobs:
                self.B.trace_on_fid(fid, KRKi, Kt, aff, kf["immature"])

    # ---------------------------------------------------------------------------------------------------------------
    def _window_description(self):
        ids = [kf["frameID"] for kf in self.kfs]
        pos = {fid: i for i, fid in enumerate(ids)}
        frames = [dict(T_w2c=kf["T_eval"], a=kf["a_eval"], b=kf["b_eval"], frameID=kf["frameID"], state=kf["state"], energyTH=kf["energyTH"]) for kf in self.kfs]
        pts = [dict(host=pos[p["host"]], u=p["u"], v=p["v"], idepth=p["idepth"], idepth_zero=p["idepth_zero"], has_prior=p["has_prior"],
                    targets=[pos[t] for t in p["targets"] if t in pos], color=p["color"], weights=p["weights"]) for p in self.points]
        return dict(n=len(self.kfs), frames=frames, points=pts), [kf["fid"] for kf in self.kfs]

    def _make_keyframe(self, fl, fr, T_w2c, first):
        B, K4 = self.B, self.K4
        K33 = TS.K33(K4)
        kf = dict(fid=fl, fid_right=fr, frameID=self.next_frame_id, T_eval=T_w2c.copy(), T_cur=T_w2c.copy(), a_eval=self.aff[0], b_eval=self.aff[1],
                  aff_cur=self.aff, state=np.zeros(10), energyTH=8 * 8 * 8, immature=None, my_type=None, frame_index=len(self.traj) - 1)
        kf["state"][6], kf["state"][7] = self.aff[0] / SCALE_A, self.aff[1] / SCALE_B
        self.next_frame_id += 1
        flagged = [self.kfs[0]["frameID"]] if len(self.kfs) >= self.max_kf else []
        self.kfs.append(kf)
        n = len(self.kfs)
        newest = n - 1
        entry = dict(frameID=kf["frameID"], n_kf=n)
        if not first:
            for p in self.points:   # a residual of every active point towards the new key frame (:1380-1395)
                p["targets"].append(kf["frameID"])
            # ---- activatePointsMT: sparsity control (:798-818), distance map, candidate loop, optimizeImmaturePoint
            npts = len(self.points)
            d = self.point_density
            if npts < d * 0.66: self.min_act_dist -= 0.8
            if npts < d * 0.8: self.min_act_dist -= 0.5
            elif npts < d * 0.9: self.min_act_dist -= 0.2
            elif npts < d: self.min_act_dist -= 0.1
            if npts > d * 1.5: self.min_act_dist += 0.8
            if npts > d * 1.3: self.min_act_dist += 0.5
            if npts > d * 1.15: self.min_act_dist += 0.2
            if npts > d: self.min_act_dist += 0.1
            self.min_act_dist = float(min(4.0, max(0.0, self.min_act_dist)))
            hosts = self.kfs[:-1]
            KK = [level1_krki_kt(mul34(T_w2c, inv34(self._pose_of(h))), K4) for h in hosts]
            KRKi1, Kt1 = np.stack([k_[0] for k_ in KK]), np.stack([k_[1] for k_ in KK])
            hpos = {h["frameID"]: i for i, h in enumerate(hosts)}
            act = [p for p in self.points if p["host"] in hpos]
            pt_host = np.array([hpos[p["host"]] for p in act], np.int32)
            order = np.argsort(pt_host, kind="stable")
            pt_uvid = np.array([[p["u"], p["v"], p["idepth"]] for p in act], np.float32).reshape(-1, 3)[order]
            B.distmap_make(KRKi1, Kt1, pt_host[order], pt_uvid)
            cand = [h["immature"] for h in hosts if h["immature"] is not None and h["immature"].size]
            if cand:
                cand_host = np.concatenate([np.full(h["immature"].size, i, np.int32) for i, h in enumerate(hosts) if h["immature"] is not None and h["immature"].size])
                my_type = np.concatenate([h["my_type"] for h in hosts if h["immature"] is not None and h["immature"].size]).astype(np.float32)
                cand = np.concatenate(cand)
                host_flag = np.array([1 if h["frameID"] in flagged else 0 for h in hosts], np.uint8)
                verdict, _ = B.activation_filter(KRKi1, Kt1, host_flag, cand_host, cand, my_type, self.min_act_dist)
                win, fids = self._window_description()
                B.window(win, fids)
                sel = np.nonzero(verdict == 1)[0]
                a = B.activate(n, cand_host[sel], np.ascontiguousarray(cand[sel])) if sel.size else dict(result=np.zeros(0, int))
                n_act = 0
                for k_ in np.nonzero(a["result"] == 1)[0]:
                    c = cand[sel[k_]]
                    targets = [self.kfs[t]["frameID"] for t in range(n) if a["states"][k_, t] == 0]
                    self.points.append(dict(host=hosts[int(cand_host[sel[k_]])]["frameID"], u=float(c["u"]), v=float(c["v"]), idepth=np.float32(a["idepth"][k_]),
                                            idepth_zero=np.float32(a["idepth"][k_]), has_prior=False, targets=targets,
                                            color=np.array(c["color"], np.float32), weights=np.array(c["weights"], np.float32)))
                    n_act += 1
                keep = verdict == 0   # 1: consumed by optimizeImmaturePoint (activated or discarded), 2: deleted
                off = 0
                for i, h in enumerate(hosts):
                    if h["immature"] is None or h["immature"].size == 0:
                        continue
                    m = h["immature"].size
                    h["immature"] = np.ascontiguousarray(h["immature"][keep[off:off + m]]); h["my_type"] = h["my_type"][keep[off:off + m]]
                    off += m
                entry.update(candidates=int(cand.size), to_optimize=int(sel.size), activated=n_act, min_act_dist=self.min_act_dist)
            # ---- the windowed optimisation
            win, fids = self._window_description()
            Wn = B.window(win, fids)
            d_ = 4 + 8 * n
            if self.HM is not None:
                HM = np.zeros((d_, d_)); bM = np.zeros(d_)
                m = self.HM.shape[0]
                HM[:m, :m] = self.HM; bM[:m] = self.bM   # the new frame enters with zero prior (EnergyFunctional::insertFrame, :490-500)
                Wn.set_marg_prior(HM, bM)
            if self.on_window:
                self.on_window(win, (self.HM, self.bM))
            rmse, its = B._op(Wn.optimize, self.opt_its)
            st = B._op(Wn.get_state)
            res = B._op(Wn.get_res, 1, brief=True)
            for i, f in enumerate(self.kfs):
                f["state"] = st["states"][i].copy()
                f["T_cur"] = st["T_w2c"][i].copy()
                f["aff_cur"] = (float(st["states"][i][6] * SCALE_A), float(st["states"][i][7] * SCALE_B))
            kf["T_eval"] = st["T_w2c"][newest].copy()     # setEvalPT of the newest frame (FullSystemOptimize.cpp:996-1005)
            kf["a_eval"], kf["b_eval"] = kf["aff_cur"]
            kf["energyTH"] = float(Wn.get_energy_th()[newest])   # what setNewFrameEnergyTH left behind in the final linearizeAll (FullSystemOptimize.cpp:163)
            self.aff = kf["aff_cur"]
            self.T_w2c_hist[-1] = kf["T_cur"].copy()
            self.traj[-1] = np.vstack([inv34(kf["T_cur"]), [0, 0, 0, 1]])
            # residuals that ended OOB / OUTLIER leave the graph (:1012-1040); points without residuals are removed (removeOutliers)
            ridx = 0
            alive = []
            centers = []   # per surviving point: centerProjectedTo of its residual into the newest key frame (if IN)
            pts_dev = B._op(Wn.get_points)
            for pi, p in enumerate(self.points):
                tg = [t for t in p["targets"] if any(t == f["frameID"] for f in self.kfs)]
                keep_t, center = [], None
                for t in tg:
                    if res["active"][ridx]:
                        keep_t.append(t)
                        if t == kf["frameID"] and res["state"][ridx] == 0:
                            center = res["center"][ridx].copy()
                    ridx += 1
                p["targets"] = keep_t
                p["idepth"] = np.float32(st["idepth"][pi])
                p["HdiF"] = float(pts_dev["HdiF"][pi])
                if keep_t:
                    alive.append(p); centers.append(center)
            n_removed = len(self.points) - len(alive)
            self.points = alive
            entry.update(rmse=float(rmse), iterations=int(its), points=len(self.points), removed=n_removed)
            # ---- setCoarseTrackingRef: splats from the window's points seen IN in the newest key frame (CoarseTracker.cpp:290-356)
            splat = np.array([[int(c[0] + 0.5), int(c[1] + 0.5), c[2], np.sqrt(np.float32(1e-3 / (np.float64(np.float32(p["HdiF"])) + 1e-12)))]
                              for p, c in zip(self.points, centers) if c is not None], np.float32).reshape(-1, 4)
            B.tracker_set_ref(kf["fid"], splat, self.aff)
            if self.on_ref:
                self.on_ref(kf["frame_index"], splat, self.aff)
            entry.update(ref_points=int(splat.shape[0]))
        # ---- makeNewTraces (:1599-1630): selector -> ImmaturePoint constructor; static stereo into the right image for the range
        uv, ty = B.make_maps(kf["fid"], self.immature_density)
        w, h = B.shape["w"], B.shape["h"]
        border = (uv[:, 0] > 4) & (uv[:, 1] > 4) & (uv[:, 0] < w - 5) & (uv[:, 1] < h - 5)   # patternPadding + 1 (:1613)
        uv, ty = np.ascontiguousarray(uv[border]), ty[border]
        pts, ok = B.immature_init_fid(kf["fid"], uv)
        pts, ty = np.ascontiguousarray(pts[ok]), ty[ok]
        pts["idepth_min_stereo"] = 0.0; pts["idepth_max_stereo"] = np.nan
        stt = B.trace_stereo(kf["fid_right"], K33, True, pts)
        good = stt == 0
        pts["idepth_min"] = np.where(good, pts["idepth_min_stereo"], pts["idepth_min"]).astype(np.float32)
        pts["idepth_max"] = np.where(good, pts["idepth_max_stereo"], pts["idepth_max"]).astype(np.float32)
        kf["immature"], kf["my_type"] = pts, ty.astype(np.float32)
        entry.update(new_immature=int(pts.size), stereo_good=int(good.sum()))
        if first:
            # the first key frame: points with a good static-stereo depth become active at once (the fork initialises from stereo,
            # FullSystem::initializeFromInitializer), the tracker reference is built from them
            gi = np.nonzero(good & (pts["idepth_stereo"] > 0))[0]
            gi = gi[:: max(1, gi.size // int(self.point_density))]
            for i in gi:
                self.points.append(dict(host=kf["frameID"], u=float(pts["u"][i]), v=float(pts["v"][i]), idepth=np.float32(pts["idepth_stereo"][i]),
                                        idepth_zero=np.float32(pts["idepth_stereo"][i]), has_prior=True, targets=[], HdiF=1e-3,
                                        color=np.array(pts["color"][i], np.float32), weights=np.array(pts["weights"][i], np.float32)))
            keep = np.ones(pts.size, bool); keep[gi] = False
            kf["immature"], kf["my_type"] = np.ascontiguousarray(pts[keep]), kf["my_type"][keep]
            splat = np.array([[p["u"], p["v"], p["idepth"], 1.0] for p in self.points], np.float32).reshape(-1, 4)
            B.tracker_set_ref(kf["fid"], splat, self.aff)
            if self.on_ref:
                self.on_ref(kf["frame_index"], splat, self.aff)
            entry.update(points=len(self.points), ref_points=int(splat.shape[0]))
        # ---- marginalisation of the flagged key frame: its points first (marginalizePointsF), then the frame (marginalizeFrame)
        if flagged:
            fid0 = flagged[0]
            flags = [1 if p["host"] == fid0 else 0 for p in self.points]
            # (window and point order are those of the last build: points removed above had no active residual and are inert)
            win, fids = self._window_description()
            Wn = B.window(win, fids)
            d_ = 4 + 8 * n
            if self.HM is not None:
                HM = np.zeros((d_, d_)); bM = np.zeros(d_)
                m = self.HM.shape[0]
                HM[:m, :m] = self.HM; bM[:m] = self.bM
                Wn.set_marg_prior(HM, bM)
            B._op(Wn.linearize_all, True)
            B.set_point_flags(flags)
            B._op(Wn.marginalize_points)
            B._op(Wn.marginalize_frame, 0)
            self.HM, self.bM = B._op(Wn.get_marg_prior)
            old = self.kfs.pop(0)
            B.release(old["fid"]); B.release(old["fid_right"])
            self.points = [p for p in self.points if p["host"] != fid0]
            for p in self.points:
                p["targets"] = [t for t in p["targets"] if t != fid0]
            self.points = [p for p in self.points if p["targets"] or p["host"] == kf["frameID"]]
            entry.update(marginalized=int(sum(flags)))
        elif self.HM is None and not first:
            pass
        self.log.append(entry)

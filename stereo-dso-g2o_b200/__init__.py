"""sdso_b200 — thin ctypes view of libsdso_b200.so (the C ABI in include/sdso_b200.h).

This module is test/bench plumbing: it owns no algorithm. Every method forwards to one C entry
point, which in turn launches the hand-written sm_100a kernels under csrc/. There is no CPU
fallback: importing succeeds without a GPU (so the symbol table can be checked), but creating a
context without a usable CUDA device raises, and a missing shared library raises at import.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdso_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sdso_b200.h")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C {_HERE}` (or __graft_entry__.build()). "
        "The B200 hot path has no CPU fallback.")
lib = C.CDLL(LIB_PATH)

VARIANT_SSE, VARIANT_G2O = 0, 1
OK = 0


class Settings(C.Structure):
    _fields_ = [
        ("huberTH", C.c_float), ("coarseCutoffTH", C.c_float), ("outlierTH", C.c_float),
        ("outlierTHSumComponent", C.c_float), ("overallEnergyTHWeight", C.c_float), ("maxPixSearch", C.c_float),
        ("minTraceTestRadius", C.c_int32), ("trace_stepsize", C.c_float), ("trace_GNIterations", C.c_int32),
        ("trace_GNThreshold", C.c_float), ("trace_extraSlackOnTH", C.c_float), ("trace_slackInterval", C.c_float),
        ("trace_minImprovementFactor", C.c_float), ("affineOptModeA", C.c_float), ("affineOptModeB", C.c_float),
        ("gammaWeightsPixelSelect", C.c_int32), ("g2o_stop_flag_persists", C.c_int32), ("cluster_size", C.c_int32),
        ("block_threads", C.c_int32), ("gather_batch", C.c_int32),
        ("idepthFixPrior", C.c_float), ("idepthFixPriorMargFac", C.c_float), ("initialRotPrior", C.c_float),
        ("initialTransPrior", C.c_float), ("initialAffBPrior", C.c_float), ("initialAffAPrior", C.c_float),
        ("initialCalibHessian", C.c_float), ("margWeightFac", C.c_float), ("solverModeDelta", C.c_double),
        ("minOptIterations", C.c_int32), ("thOptIterations", C.c_float), ("frameEnergyTHConstWeight", C.c_float),
        ("frameEnergyTHN", C.c_float), ("frameEnergyTHFacMedian", C.c_float),
        ("minGradHistCut", C.c_float), ("minGradHistAdd", C.c_float), ("gradDownweightPerLevel", C.c_float),
        ("desiredImmatureDensity", C.c_float), ("minTraceQuality", C.c_float), ("track_cache", C.c_int32),
    ]


def default_settings():
    s = Settings()
    lib.sdso_default_settings(C.byref(s))
    return s


_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(t)


lib.sdso_last_error.restype = C.c_char_p
lib.sdso_launch_count.restype = C.c_uint64
lib.sdso_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, _fp, C.c_float, C.POINTER(Settings)]
lib.sdso_ctx_destroy.argtypes = [C.c_void_p]
lib.sdso_last_error.argtypes = [C.c_void_p]
lib.sdso_launch_count.argtypes = [C.c_void_p]
lib.sdso_set_stream.argtypes = [C.c_void_p, C.c_void_p]
lib.sdso_synchronize.argtypes = [C.c_void_p]
lib.sdso_pyr_levels.argtypes = [C.c_void_p]
lib.sdso_level_size.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
lib.sdso_level_K.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
lib.sdso_frame_create.argtypes = [C.c_void_p, _ip]
lib.sdso_frame_release.argtypes = [C.c_void_p, C.c_int]
lib.sdso_make_images.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int]
lib.sdso_make_images_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int]
lib.sdso_frame_download.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp]
lib.sdso_interp33.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_int, _fp, C.c_int]
lib.sdso_tracker_make_k.argtypes = [C.c_void_p, _fp]
lib.sdso_tracker_set_ref.argtypes = [C.c_void_p, C.c_int, _fp, C.c_int, _dp]
lib.sdso_tracker_set_pc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _dp]
lib.sdso_tracker_get_pc.argtypes = [C.c_void_p, C.c_int, _ip, _fp, _fp, _fp, _fp]
lib.sdso_calc_res_gs.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_float, _dp, _dp, _dp, _ip, _fp]
lib.sdso_track.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _ip, _ip]
lib.sdso_track_enqueue.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, C.c_int, _dp, C.c_int]
lib.sdso_edge_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _ip, _dp, _dp]
lib.sdso_profile_enable.argtypes = [C.c_void_p, C.c_int]
lib.sdso_profile_read.argtypes = [C.c_void_p, _dp, _ip, _dp, _ip]
lib.sdso_set_gamma.argtypes = [C.c_void_p, _fp]
lib.sdso_track_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
lib.sdso_track_collect.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip, _ip, C.POINTER(C.c_uint64)]


class SdsoError(RuntimeError):
    pass


class Context:
    """One sdso_ctx (one GPU, one working resolution)."""

    def __init__(self, w, h, K, baseline=0.0, device=0, settings=None):
        self._h = C.c_void_p()
        Kc = (C.c_float * 4)(*[float(x) for x in K])
        sp = C.byref(settings) if settings is not None else None
        rc = lib.sdso_ctx_create(C.byref(self._h), device, w, h, Kc, float(baseline), sp)
        if rc != OK:
            self._h = C.c_void_p()
            raise SdsoError(f"sdso_ctx_create failed with code {rc} (no CUDA device? there is no CPU fallback)")
        self.w, self.h = w, h
        self.levels = lib.sdso_pyr_levels(self._h)

    def close(self):
        if self._h:
            lib.sdso_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise SdsoError(f"sdso error {rc}: {lib.sdso_last_error(self._h).decode()}")

    # -- plumbing
    def set_stream(self, stream_ptr):
        self._ck(lib.sdso_set_stream(self._h, C.c_void_p(stream_ptr)))

    def synchronize(self):
        self._ck(lib.sdso_synchronize(self._h))

    def launch_count(self):
        return int(lib.sdso_launch_count(self._h))

    def level_size(self, lvl):
        w, h = C.c_int(), C.c_int()
        self._ck(lib.sdso_level_size(self._h, lvl, C.byref(w), C.byref(h)))
        return w.value, h.value

    def level_K(self, lvl):
        K = np.zeros(9, np.float32)
        Ki = np.zeros(9, np.float32)
        self._ck(lib.sdso_level_K(self._h, lvl, _ptr(K, _fp), _ptr(Ki, _fp)))
        return K.reshape(3, 3), Ki.reshape(3, 3)

    # -- A1
    def frame_create(self):
        fid = C.c_int()
        self._ck(lib.sdso_frame_create(self._h, C.byref(fid)))
        return fid.value

    def frame_release(self, fid):
        self._ck(lib.sdso_frame_release(self._h, fid))

    def make_images(self, fid, image, exposure=1.0, use_hcalib=True):
        img = _f32(image)
        assert img.size == self.w * self.h
        self._ck(lib.sdso_make_images(self._h, fid, img.ctypes.data, float(exposure), int(use_hcalib)))

    def make_images_ptr(self, fid, host_ptr, exposure=1.0, use_hcalib=True):
        self._ck(lib.sdso_make_images(self._h, fid, C.c_void_p(host_ptr), float(exposure), int(use_hcalib)))

    def make_images_device(self, fid, dev_ptr, exposure=1.0, use_hcalib=True):
        self._ck(lib.sdso_make_images_device(self._h, fid, C.c_void_p(dev_ptr), float(exposure), int(use_hcalib)))

    def frame_download(self, fid, lvl):
        w, h = self.level_size(lvl)
        dI = np.zeros((h, w, 3), np.float32)
        ag = np.zeros((h, w), np.float32)
        self._ck(lib.sdso_frame_download(self._h, fid, lvl, _ptr(dI, _fp), _ptr(ag, _fp)))
        return dI, ag

    def interp33(self, fid, lvl, xy, bilin=False):
        xy = _f32(xy).reshape(-1, 2)
        out = np.zeros((xy.shape[0], 3), np.float32)
        self._ck(lib.sdso_interp33(self._h, fid, lvl, _ptr(xy, _fp), xy.shape[0], _ptr(out, _fp), int(bilin)))
        return out

    # -- A3/A4
    def tracker_make_k(self, K):
        Kc = _f32(K)
        self._ck(lib.sdso_tracker_make_k(self._h, _ptr(Kc, _fp)))

    def tracker_set_ref(self, fid, uvidw, aff=(0.0, 0.0)):
        p = _f32(uvidw).reshape(-1, 4)
        a = _f64(aff)
        self._ck(lib.sdso_tracker_set_ref(self._h, fid, _ptr(p, _fp), p.shape[0], _ptr(a, _dp)))

    def tracker_set_pc(self, fid, lvl, u, v, idepth, color, aff=(0.0, 0.0)):
        u, v, idepth, color = _f32(u), _f32(v), _f32(idepth), _f32(color)
        a = _f64(aff)
        self._ck(lib.sdso_tracker_set_pc(self._h, fid, lvl, u.size, _ptr(u, _fp), _ptr(v, _fp), _ptr(idepth, _fp),
                                         _ptr(color, _fp), _ptr(a, _dp)))

    def tracker_get_pc(self, lvl):
        w, h = self.level_size(lvl)
        cap = w * h
        u, v, idp, col = (np.zeros(cap, np.float32) for _ in range(4))
        n = C.c_int()
        self._ck(lib.sdso_tracker_get_pc(self._h, lvl, C.byref(n), _ptr(u, _fp), _ptr(v, _fp), _ptr(idp, _fp), _ptr(col, _fp)))
        n = n.value
        return u[:n].copy(), v[:n].copy(), idp[:n].copy(), col[:n].copy()

    # -- A5/A6
    def calc_res_gs(self, new_fid, lvl, T, aff, cutoff, want_warped=True):
        T = _f64(T).reshape(12)
        aff = _f64(aff)
        rs, H, b = np.zeros(6), np.zeros(64), np.zeros(8)
        wn = C.c_int()
        w, h = self.level_size(lvl)
        warped = np.zeros(8 * (w * h + 4), np.float32) if want_warped else None
        self._ck(lib.sdso_calc_res_gs(self._h, new_fid, lvl, _ptr(T, _dp), _ptr(aff, _dp), float(cutoff), _ptr(rs, _dp),
                                      _ptr(H, _dp), _ptr(b, _dp), C.byref(wn), _ptr(warped, _fp) if want_warped else None))
        out = dict(rs=rs, H=H.reshape(8, 8), b=b, warped_n=wn.value)
        if want_warped:
            out["warped"] = warped[:8 * wn.value].reshape(8, wn.value).copy()
        return out

    # -- E1
    def edge_eval(self, new_fid, lvl, T_select, T_pose, photo):
        Ts, Tp, ph = _f64(T_select).reshape(12), _f64(T_pose).reshape(12), _f64(photo)
        w, h = self.level_size(lvl)
        err, J = np.zeros(w * h), np.zeros((w * h, 8))
        n = C.c_int()
        self._ck(lib.sdso_edge_eval(self._h, new_fid, lvl, _ptr(Ts, _dp), _ptr(Tp, _dp), _ptr(ph, _dp), C.byref(n), _ptr(err, _dp), _ptr(J, _dp)))
        return err[:n.value].copy(), J[:n.value].copy()

    def profile_enable(self, on=True):
        self._ck(lib.sdso_profile_enable(self._h, int(on)))

    def profile_read(self):
        t, m = C.c_double(), C.c_double()
        nt, nm = C.c_int(), C.c_int()
        self._ck(lib.sdso_profile_read(self._h, C.byref(t), C.byref(nt), C.byref(m), C.byref(nm)))
        return dict(track_ms=t.value, track_launches=nt.value, images_ms=m.value, images_launches=nm.value)

    def track_phase_cycles(self):
        c = (C.c_longlong * 16)()
        self._ck(lib.sdso_track_phase_cycles(self._h, c))
        return list(c)

    def set_gamma(self, B):
        B = _f32(B)
        assert B.size == 256
        self._ck(lib.sdso_set_gamma(self._h, _ptr(B, _fp)))

    # -- A7
    def track(self, new_fid, T, aff, coarsest, min_res_for_abort, variant=VARIANT_SSE):
        T = _f64(T).reshape(12).copy()
        aff = _f64(aff).copy()
        mr = _f64(min_res_for_abort)
        lr, fl = np.zeros(5), np.zeros(3)
        it = np.zeros(5, np.int32)
        ok = C.c_int()
        self._ck(lib.sdso_track(self._h, new_fid, _ptr(T, _dp), _ptr(aff, _dp), coarsest, _ptr(mr, _dp), variant,
                                _ptr(lr, _dp), _ptr(fl, _dp), _ptr(it, _ip), C.byref(ok)))
        return dict(T=T.reshape(3, 4), aff=aff, lastResiduals=lr, flow=fl, iterations=it, ok=bool(ok.value))

    def track_enqueue(self, new_fids, T, aff, coarsest, min_res_for_abort, variant=VARIANT_SSE):
        nb = len(new_fids)
        f = np.ascontiguousarray(new_fids, dtype=np.int32)
        T = _f64(T).reshape(nb, 12)
        aff = _f64(aff).reshape(nb, 2)
        mr = _f64(min_res_for_abort).reshape(nb, 5)
        self._ck(lib.sdso_track_enqueue(self._h, nb, _ptr(f, _ip), _ptr(T, _dp), _ptr(aff, _dp), coarsest, _ptr(mr, _dp), variant))

    def track_collect(self, nb):
        T, aff, lr, fl = np.zeros((nb, 12)), np.zeros((nb, 2)), np.zeros((nb, 5)), np.zeros((nb, 3))
        it, ok = np.zeros((nb, 5), np.int32), np.zeros(nb, np.int32)
        ev = C.c_uint64()
        self._ck(lib.sdso_track_collect(self._h, nb, _ptr(T, _dp), _ptr(aff, _dp), _ptr(lr, _dp), _ptr(fl, _dp), _ptr(it, _ip),
                                        _ptr(ok, _ip), C.byref(ev)))
        return dict(T=T.reshape(nb, 3, 4), aff=aff, lastResiduals=lr, flow=fl, iterations=it, ok=ok.astype(bool), evals=int(ev.value))


# ---------------------------------------------------------------------------------------------------
# B1-B12: windowed bundle adjustment (EnergyFunctional / PointFrameResidual operator surface)
_u8p = C.POINTER(C.c_ubyte)
lib.sdso_ba_reset.argtypes = [C.c_void_p]
lib.sdso_ba_set_calib.argtypes = [C.c_void_p, _fp, _dp]
lib.sdso_ba_add_frame.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_double, C.c_int, _ip]
lib.sdso_ba_set_state.argtypes = [C.c_void_p, C.c_int, _dp]
lib.sdso_ba_set_energy_th.argtypes = [C.c_void_p, C.c_int, C.c_float]
lib.sdso_ba_set_points.argtypes = [C.c_void_p, C.c_int, _ip, _fp, _fp, _fp, _fp, _fp, _fp, _u8p]
lib.sdso_ba_set_residuals.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
lib.sdso_ba_set_point_flags.argtypes = [C.c_void_p, _u8p]
lib.sdso_ba_prepare.argtypes = [C.c_void_p]
lib.sdso_ba_counts.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
lib.sdso_ba_get_precalc.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp]
lib.sdso_ba_get_adjoints.argtypes = [C.c_void_p, _dp, _dp, _fp]
lib.sdso_ba_nullspaces.argtypes = [C.c_void_p, _dp]
lib.sdso_ba_linearize_all.argtypes = [C.c_void_p, C.c_int, _dp]
lib.sdso_ba_apply_res.argtypes = [C.c_void_p, C.c_int]
lib.sdso_ba_fix_linearization.argtypes = [C.c_void_p, C.c_int, _ip]
lib.sdso_ba_get_res.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _ip, _ip, _fp, _fp, _fp, _fp]
lib.sdso_ba_get_points.argtypes = [C.c_void_p, _fp]
lib.sdso_ba_accumulate_top.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _fp]
lib.sdso_ba_accumulate_sc.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
lib.sdso_ba_solve.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, _dp, _dp]
lib.sdso_ba_resubstitute.argtypes = [C.c_void_p, _dp, _dp, _dp]
lib.sdso_ba_set_marg_prior.argtypes = [C.c_void_p, _dp, _dp]
lib.sdso_ba_get_marg_prior.argtypes = [C.c_void_p, _dp, _dp]


class Window:
    """The sliding window of one Context: index-based mirror of EnergyFunctional + PointFrameResidual
    (frames / points / residuals as integer ids)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.h = ctx._h
        self._ck = ctx._ck
        self._ck(lib.sdso_ba_reset(self.h))

    def set_calib(self, K, delta=None):
        Kc = _f32(K)
        d = _f64(delta) if delta is not None else None
        self._ck(lib.sdso_ba_set_calib(self.h, _ptr(Kc, _fp), _ptr(d, _dp) if d is not None else None))

    def add_frame(self, fid, T_w2c, a=0.0, b=0.0, frameID=1):
        T = _f64(T_w2c).reshape(12)
        idx = C.c_int()
        self._ck(lib.sdso_ba_add_frame(self.h, fid, _ptr(T, _dp), a, b, frameID, C.byref(idx)))
        return idx.value

    def set_state(self, idx, state10):
        s = _f64(state10)
        self._ck(lib.sdso_ba_set_state(self.h, idx, _ptr(s, _dp)))

    def set_energy_th(self, idx, th):
        self._ck(lib.sdso_ba_set_energy_th(self.h, idx, float(th)))

    def set_points(self, host, u, v, idepth, idepth_zero, color8, weights8, has_prior):
        host = np.ascontiguousarray(host, np.int32)
        u, v, idepth, idepth_zero = _f32(u), _f32(v), _f32(idepth), _f32(idepth_zero)
        c, w = _f32(color8).reshape(-1, 8), _f32(weights8).reshape(-1, 8)
        hp = np.ascontiguousarray(has_prior, np.uint8)
        self._ck(lib.sdso_ba_set_points(self.h, host.size, _ptr(host, _ip), _ptr(u, _fp), _ptr(v, _fp), _ptr(idepth, _fp),
                                        _ptr(idepth_zero, _fp), _ptr(c, _fp), _ptr(w, _fp), _ptr(hp, _u8p)))

    def set_residuals(self, point, target):
        p, t = np.ascontiguousarray(point, np.int32), np.ascontiguousarray(target, np.int32)
        self._ck(lib.sdso_ba_set_residuals(self.h, p.size, _ptr(p, _ip), _ptr(t, _ip)))

    def set_point_flags(self, flags):
        f = np.ascontiguousarray(flags, np.uint8)
        self._ck(lib.sdso_ba_set_point_flags(self.h, _ptr(f, _u8p)))

    def prepare(self):
        self._ck(lib.sdso_ba_prepare(self.h))

    def counts(self):
        n, P, R, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(lib.sdso_ba_counts(self.h, C.byref(n), C.byref(P), C.byref(R), C.byref(d)))
        return dict(frames=n.value, points=P.value, res=R.value, dim=d.value)

    def precalc(self, h, t):
        out = np.zeros(49, np.float32)
        self._ck(lib.sdso_ba_get_precalc(self.h, h, t, _ptr(out, _fp)))
        return out

    def adjoints(self):
        n = self.counts()["frames"]
        ah, at, d = np.zeros((n * n, 8, 8)), np.zeros((n * n, 8, 8)), np.zeros((n * n, 8), np.float32)
        self._ck(lib.sdso_ba_get_adjoints(self.h, _ptr(ah, _dp), _ptr(at, _dp), _ptr(d, _fp)))
        return ah, at, d

    def nullspaces(self):
        N = np.zeros((self.counts()["dim"], 7))
        self._ck(lib.sdso_ba_nullspaces(self.h, _ptr(N, _dp)))
        return N

    def linearize_all(self, fix=False):
        e = C.c_double()
        self._ck(lib.sdso_ba_linearize_all(self.h, int(fix), C.byref(e)))
        return e.value

    def linearize_all_async(self, fix=False):
        self._ck(lib.sdso_ba_linearize_all(self.h, int(fix), None))

    def apply_res(self, copy=True):
        self._ck(lib.sdso_ba_apply_res(self.h, int(copy)))

    def fix_linearization(self, rids=None):
        if rids is None:
            self._ck(lib.sdso_ba_fix_linearization(self.h, 0, None))
        else:
            r = np.ascontiguousarray(rids, np.int32)
            self._ck(lib.sdso_ba_fix_linearization(self.h, r.size, _ptr(r, _ip)))

    def get_res(self, which=0, brief=False):
        """per-residual read-back in the caller's order. brief=True: states, energies, flags and centerProjectedTo only (what
        FullSystem's bookkeeping after optimize() reads) — the Jacobian blocks (600 B per residual) stay on the device."""
        R = self.counts()["res"]
        ns, st, ac, li = (np.zeros(R, np.int32) for _ in range(4))
        ne, nw = np.zeros(R), np.zeros(R)
        ce = np.zeros((R, 3), np.float32)
        if brief:
            self._ck(lib.sdso_ba_get_res(self.h, which, _ptr(ns, _ip), _ptr(st, _ip), _ptr(ne, _dp), _ptr(nw, _dp), _ptr(ac, _ip), _ptr(li, _ip),
                                         None, None, _ptr(ce, _fp), None))
            return dict(newState=ns, state=st, newEnergy=ne, newEnergyWithOutlier=nw, active=ac, linearized=li, center=ce)
        J, jp, rz = np.zeros((R, 74), np.float32), np.zeros((R, 8), np.float32), np.zeros((R, 8), np.float32)
        self._ck(lib.sdso_ba_get_res(self.h, which, _ptr(ns, _ip), _ptr(st, _ip), _ptr(ne, _dp), _ptr(nw, _dp), _ptr(ac, _ip), _ptr(li, _ip),
                                     _ptr(J, _fp), _ptr(jp, _fp), _ptr(ce, _fp), _ptr(rz, _fp)))
        return dict(newState=ns, state=st, newEnergy=ne, newEnergyWithOutlier=nw, active=ac, linearized=li, J=J, JpJdF=jp, center=ce, res_toZero=rz)

    def get_points(self):
        P = self.counts()["points"]
        o = np.zeros((P, 16), np.float32)
        self._ck(lib.sdso_ba_get_points(self.h, _ptr(o, _fp)))
        return dict(Hdd_A=o[:, 0], bd_A=o[:, 1], Hcd_A=o[:, 2:6], Hdd_L=o[:, 6], bd_L=o[:, 7], Hcd_L=o[:, 8:12], HdiF=o[:, 12], bdSumF=o[:, 13],
                    step=o[:, 14], priorF=o[:, 15])

    def accumulate_top(self, mode, use_prior):
        c = self.counts()
        d, n = c["dim"], c["frames"]
        H, b, blk = np.zeros((d, d)), np.zeros(d), np.zeros((n * n, 13, 13), np.float32)
        self._ck(lib.sdso_ba_accumulate_top(self.h, mode, int(use_prior), _ptr(H, _dp), _ptr(b, _dp), _ptr(blk, _fp)))
        return H, b, blk

    def accumulate_sc(self, shift=True):
        d = self.counts()["dim"]
        H, b = np.zeros((d, d)), np.zeros(d)
        self._ck(lib.sdso_ba_accumulate_sc(self.h, int(shift), _ptr(H, _dp), _ptr(b, _dp)))
        return H, b

    def solve(self, iteration, lam=1e-5):
        d = self.counts()["dim"]
        x, H, b = np.zeros(d), np.zeros((d, d)), np.zeros(d)
        self._ck(lib.sdso_ba_solve(self.h, iteration, lam, _ptr(x, _dp), _ptr(H, _dp), _ptr(b, _dp)))
        return x, H, b

    def solve_async(self, iteration, lam=1e-5):
        self._ck(lib.sdso_ba_solve(self.h, iteration, lam, None, None, None))

    def resubstitute(self, x=None):
        n = self.counts()["frames"]
        xs = _f64(x) if x is not None else None
        fs, cs = np.zeros((n, 10)), np.zeros(4)
        self._ck(lib.sdso_ba_resubstitute(self.h, _ptr(xs, _dp) if xs is not None else None, _ptr(fs, _dp), _ptr(cs, _dp)))
        return fs, cs

    def set_marg_prior(self, HM, bM):
        HM, bM = _f64(HM), _f64(bM)
        self._ck(lib.sdso_ba_set_marg_prior(self.h, _ptr(HM, _dp), _ptr(bM, _dp)))

    def get_marg_prior(self):
        d = self.counts()["dim"]
        HM, bM = np.zeros((d, d)), np.zeros(d)
        self._ck(lib.sdso_ba_get_marg_prior(self.h, _ptr(HM, _dp), _ptr(bM, _dp)))
        return HM, bM


# ---------------------------------------------------------------------------------------------------
# D1-D3: immature points (ImmaturePoint constructor, traceOn, traceStereo)
IMMATURE_DTYPE = np.dtype([
    ("u", "f4"), ("v", "f4"), ("idepth_min", "f4"), ("idepth_max", "f4"), ("quality", "f4"), ("energyTH", "f4"),
    ("color", "f4", 8), ("weights", "f4", 8), ("gradH", "f4", 4),
    ("u_stereo", "f4"), ("v_stereo", "f4"), ("idepth_min_stereo", "f4"), ("idepth_max_stereo", "f4"), ("idepth_stereo", "f4"),
    ("lastTraceUV", "f4", 2), ("lastTracePixelInterval", "f4"),
    ("lastTraceStatus", "i4"), ("bestIdx", "i4"), ("numSteps", "i4")])
IPS_GOOD, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED = range(6)
lib.sdso_immature_init.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_void_p, _ip]
lib.sdso_trace_on.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp, C.c_int, C.c_void_p, _ip]
lib.sdso_trace_stereo.argtypes = [C.c_void_p, C.c_int, _fp, C.c_int, C.c_int, C.c_void_p, _ip]


def _immature_init(self, host_fid, uv):
    uv = _f32(uv).reshape(-1, 2)
    n = uv.shape[0]
    pts = np.zeros(n, IMMATURE_DTYPE)
    ok = np.zeros(n, np.int32)
    self._ck(lib.sdso_immature_init(self._h, host_fid, n, _ptr(uv, _fp), pts.ctypes.data, _ptr(ok, _ip)))
    return pts, ok.astype(bool)


def _trace_on(self, fid, KRKi, Kt, aff, pts):
    """In-place update of the records; returns the status array."""
    assert pts.dtype == IMMATURE_DTYPE and pts.flags["C_CONTIGUOUS"]
    K_, t_, a_ = _f32(KRKi).reshape(9), _f32(Kt).reshape(3), _f32(aff).reshape(2)
    st = np.zeros(pts.size, np.int32)
    self._ck(lib.sdso_trace_on(self._h, fid, _ptr(K_, _fp), _ptr(t_, _fp), _ptr(a_, _fp), pts.size, pts.ctypes.data, _ptr(st, _ip)))
    return st


def _trace_stereo(self, fid, K, mode_right, pts):
    assert pts.dtype == IMMATURE_DTYPE and pts.flags["C_CONTIGUOUS"]
    K_ = _f32(K).reshape(9)
    st = np.zeros(pts.size, np.int32)
    self._ck(lib.sdso_trace_stereo(self._h, fid, _ptr(K_, _fp), int(mode_right), pts.size, pts.ctypes.data, _ptr(st, _ip)))
    return st


lib.sdso_trace_on_hosts.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp, _fp, C.c_int, _ip, C.c_void_p, _ip]
lib.sdso_trace_stereo_resident.argtypes = [C.c_void_p, C.c_int, _fp, C.c_int, C.c_int, _ip]
lib.sdso_immature_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
lib.sdso_immature_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]


def _trace_on_hosts(self, fid, KRKi, Kt, aff, host_of_point, pts=None, want_status=True):
    """traceOn of the points of SEVERAL host key frames in one launch. pts=None: the device-resident pool (immature_upload)."""
    K_, t_, a_ = _f32(KRKi).reshape(-1, 9), _f32(Kt).reshape(-1, 3), _f32(aff).reshape(-1, 2)
    ho = np.ascontiguousarray(host_of_point, dtype=np.int32)
    if pts is not None:
        assert pts.dtype == IMMATURE_DTYPE and pts.flags["C_CONTIGUOUS"] and pts.size == ho.size
    st = np.zeros(ho.size, np.int32) if want_status else None
    self._ck(lib.sdso_trace_on_hosts(self._h, fid, K_.shape[0], _ptr(K_, _fp), _ptr(t_, _fp), _ptr(a_, _fp), ho.size, _ptr(ho, _ip),
                                     pts.ctypes.data if pts is not None else None, _ptr(st, _ip) if want_status else None))
    return st


def _trace_stereo_resident(self, fid, K, mode_right, n, want_status=True):
    K_ = _f32(K).reshape(9)
    st = np.zeros(n, np.int32) if want_status else None
    self._ck(lib.sdso_trace_stereo_resident(self._h, fid, _ptr(K_, _fp), int(mode_right), n, _ptr(st, _ip) if want_status else None))
    return st


def _immature_upload(self, pts):
    assert pts.dtype == IMMATURE_DTYPE and pts.flags["C_CONTIGUOUS"]
    self._ck(lib.sdso_immature_upload(self._h, pts.size, pts.ctypes.data))


def _immature_download(self, n, first=0):
    pts = np.zeros(n, IMMATURE_DTYPE)
    self._ck(lib.sdso_immature_download(self._h, first, n, pts.ctypes.data))
    return pts


Context.immature_init = _immature_init
Context.trace_on = _trace_on
Context.trace_stereo = _trace_stereo
Context.trace_on_hosts = _trace_on_hosts
Context.trace_stereo_resident = _trace_stereo_resident
Context.immature_upload = _immature_upload
Context.immature_download = _immature_download


# ---------------------------------------------------------------------------------------------------
# V1-V5, E3: g2o vertices / trace edge as operators over SoA batches
VERTEX_SE3_POSE, VERTEX_PHOTOMETRIC, VERTEX_INVERSE_DEPTH, VERTEX_UV, VERTEX_CAM = 1, 2, 3, 4, 5
lib.sdso_vertex_oplus.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
lib.sdso_edge_trace_uv_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _fp, _dp, _fp, _dp, _dp, _dp, _ip]


def _vertex_oplus(self, kind, estimate, update, aux=None):
    """oplusImpl of a batch of vertices of one kind; returns the updated estimates (same shape as `estimate`)."""
    est = np.array(estimate, np.float64, order="C")
    shape = est.shape
    width = {VERTEX_SE3_POSE: 12, VERTEX_PHOTOMETRIC: 2, VERTEX_INVERSE_DEPTH: 1, VERTEX_UV: 2, VERTEX_CAM: 4}[kind]
    n = est.size // width
    upd = _f64(update)
    a = _f64(aux) if aux is not None else None
    self._ck(lib.sdso_vertex_oplus(self._h, kind, n, _ptr(est, _dp), _ptr(upd, _dp), _ptr(a, _dp) if a is not None else None))
    return est.reshape(shape)


def _edge_trace_uv_eval(self, fid, uv, rot, meas, aff, dxdy, error=None, J=None):
    """EdgeTracePointUVDSO computeError + linearizeOplus for n edges; error / J (optional) carry the members' previous contents."""
    uv_, rot_, me_, dx_ = _f64(uv).reshape(-1, 2), _f32(rot).reshape(-1, 2), _f64(meas).reshape(-1), _f64(dxdy).reshape(-1, 2)
    n = uv_.shape[0]
    a_ = _f32(aff).reshape(2)
    err = np.zeros(n) if error is None else np.array(error, np.float64)
    Jo = np.zeros(n) if J is None else np.array(J, np.float64)
    flag = np.zeros(n, np.int32)
    self._ck(lib.sdso_edge_trace_uv_eval(self._h, fid, n, _ptr(uv_, _dp), _ptr(rot_, _fp), _ptr(me_, _dp), _ptr(a_, _fp), _ptr(dx_, _dp),
                                         _ptr(err, _dp), _ptr(Jo, _dp), _ptr(flag, _ip)))
    return err, Jo, flag


Context.vertex_oplus = _vertex_oplus
Context.edge_trace_uv_eval = _edge_trace_uv_eval


# ---------------------------------------------------------------------------------------------------
# point-sharded windowed BA
lib.sdso_shard_range.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip]
lib.sdso_ba_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.sdso_ba_assemble.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), _ip]
lib.sdso_ba_allreduce.argtypes = [C.c_void_p, _dp]
lib.sdso_ba_solve_assembled.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
lib.sdso_nccl_unique_id.argtypes = [C.c_void_p]
lib.sdso_nccl_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.sdso_nccl_destroy.argtypes = [C.c_void_p]
lib.sdso_allreduce_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int]


def shard_range(npoints, rank, nranks):
    b, e = C.c_int(), C.c_int()
    rc = lib.sdso_shard_range(npoints, rank, nranks, C.byref(b), C.byref(e))
    if rc != OK:
        raise SdsoError(f"sdso_shard_range({npoints}, {rank}, {nranks}) -> {rc}")
    return b.value, e.value


def nccl_unique_id():
    buf = (C.c_ubyte * 128)()
    rc = lib.sdso_nccl_unique_id(buf)
    if rc != OK:
        raise SdsoError(f"sdso_nccl_unique_id -> {rc} (NCCL not loadable)")
    return bytes(buf)


def _nccl_init(self, rank, nranks, uid):
    buf = (C.c_ubyte * 128).from_buffer_copy(uid)
    self._ck(lib.sdso_nccl_init(self._h, rank, nranks, buf))


Context.nccl_init = _nccl_init

lib.sdso_peer_alloc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.sdso_peer_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.sdso_peer_select.argtypes = [C.c_void_p, C.c_int]
lib.sdso_peer_status.argtypes = [C.c_void_p, C.POINTER(C.c_int)]


def _peer_alloc(self, nranks, max_doubles):
    """allocate this rank's exchange block; returns its 64-byte CUDA IPC handle (to be all-gathered by the caller's rendezvous)"""
    buf = (C.c_ubyte * 64)()
    self._ck(lib.sdso_peer_alloc(self._h, int(nranks), int(max_doubles), buf))
    return bytes(buf)


def _peer_connect(self, rank, nranks, handles):
    blob = b"".join(handles)
    assert len(blob) == 64 * nranks
    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    self._ck(lib.sdso_peer_connect(self._h, rank, nranks, buf))


def _peer_select(self, which):
    self._ck(lib.sdso_peer_select(self._h, int(which)))


def _peer_status(self):
    v = C.c_int()
    self._ck(lib.sdso_peer_status(self._h, C.byref(v)))
    return v.value


Context.peer_alloc = _peer_alloc
Context.peer_connect = _peer_connect
Context.peer_select = _peer_select
Context.peer_status = _peer_status


def _w_set_shard(self, rank, nranks):
    self._ck(lib.sdso_ba_set_shard(self.h, rank, nranks))


def _w_assemble(self):
    ptr, cnt = C.c_void_p(), C.c_int()
    self._ck(lib.sdso_ba_assemble(self.h, C.byref(ptr), C.byref(cnt)))
    return ptr.value, cnt.value


def _w_allreduce(self, want_energy=False):
    e = C.c_double()
    self._ck(lib.sdso_ba_allreduce(self.h, C.byref(e) if want_energy else None))
    return e.value if want_energy else None


def _w_solve_assembled(self, iteration, want=True):
    d = self.counts()["dim"]
    if not want:
        self._ck(lib.sdso_ba_solve_assembled(self.h, iteration, None, None, None))
        return None
    x, H, b = np.zeros(d), np.zeros((d, d)), np.zeros(d)
    self._ck(lib.sdso_ba_solve_assembled(self.h, iteration, _ptr(x, _dp), _ptr(H, _dp), _ptr(b, _dp)))
    return x, H, b


Window.set_shard = _w_set_shard
Window.assemble = _w_assemble
Window.allreduce = _w_allreduce
Window.solve_assembled = _w_solve_assembled


# ---------------------------------------------------------------------------------------------------
# batched makeImages (uint8 / float sources, asynchronous upload) and multi-reference tracking
_vpp = C.POINTER(C.c_void_p)
lib.sdso_upload_images_async.argtypes = [C.c_void_p, C.c_int, _ip, _vpp, C.c_int]
lib.sdso_make_images_uploaded.argtypes = [C.c_void_p, C.c_int, _ip, _fp, C.c_int]
lib.sdso_make_images_batch_device.argtypes = [C.c_void_p, C.c_int, _ip, _vpp, C.c_int, _fp, C.c_int]
lib.sdso_tracker_select_ref.argtypes = [C.c_void_p, C.c_int]
lib.sdso_track_enqueue_multi.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, C.c_int, _dp, C.c_int]


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def _upload_images_async(self, fids, host_ptrs, u8=False):
    f = np.ascontiguousarray(fids, dtype=np.int32)
    self._ck(lib.sdso_upload_images_async(self._h, f.size, _ptr(f, _ip), _ptr_array(host_ptrs), int(u8)))


def _make_images_uploaded(self, fids, use_hcalib=True):
    f = np.ascontiguousarray(fids, dtype=np.int32)
    self._ck(lib.sdso_make_images_uploaded(self._h, f.size, _ptr(f, _ip), None, int(use_hcalib)))


def _make_images_batch_device(self, fids, dev_ptrs, u8=False, use_hcalib=True):
    f = np.ascontiguousarray(fids, dtype=np.int32)
    self._ck(lib.sdso_make_images_batch_device(self._h, f.size, _ptr(f, _ip), _ptr_array(dev_ptrs), int(u8), None, int(use_hcalib)))


def _tracker_select_ref(self, slot):
    self._ck(lib.sdso_tracker_select_ref(self._h, int(slot)))


def _track_enqueue_multi(self, ref_slots, new_fids, T, aff, coarsest, min_res_for_abort, variant=VARIANT_SSE):
    nb = len(new_fids)
    f = np.ascontiguousarray(new_fids, dtype=np.int32)
    r = np.ascontiguousarray(ref_slots, dtype=np.int32) if ref_slots is not None else None
    T = _f64(T).reshape(nb, 12)
    aff = _f64(aff).reshape(nb, 2)
    mr = _f64(min_res_for_abort).reshape(nb, 5)
    self._ck(lib.sdso_track_enqueue_multi(self._h, nb, _ptr(r, _ip) if r is not None else None, _ptr(f, _ip), _ptr(T, _dp), _ptr(aff, _dp),
                                          coarsest, _ptr(mr, _dp), variant))


Context.upload_images_async = _upload_images_async
Context.make_images_uploaded = _make_images_uploaded
Context.make_images_batch_device = _make_images_batch_device
Context.tracker_select_ref = _tracker_select_ref
Context.track_enqueue_multi = _track_enqueue_multi


lib.sdso_ba_new_frame_energy_th.argtypes = [C.c_void_p, _fp]
lib.sdso_ba_optimize.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
lib.sdso_ba_get_state.argtypes = [C.c_void_p, _dp, _dp, _fp, _dp]


def _w_new_frame_energy_th(self):
    th = C.c_float()
    self._ck(lib.sdso_ba_new_frame_energy_th(self.h, C.byref(th)))
    return th.value


def _w_optimize(self, iters=6):
    r, d = C.c_double(), C.c_int()
    self._ck(lib.sdso_ba_optimize(self.h, iters, C.byref(r), C.byref(d)))
    return r.value, d.value


def _w_get_state(self):
    c = self.counts()
    st, T, idp, cal = np.zeros((c["frames"], 10)), np.zeros((c["frames"], 3, 4)), np.zeros(c["points"], np.float32), np.zeros(4)
    self._ck(lib.sdso_ba_get_state(self.h, _ptr(st, _dp), _ptr(T, _dp), _ptr(idp, _fp), _ptr(cal, _dp)))
    return dict(states=st, T_w2c=T, idepth=idp, calib=cal)


lib.sdso_ba_get_energy_th.argtypes = [C.c_void_p, _fp]


def _w_get_energy_th(self):
    th = np.zeros(self.counts()["frames"], np.float32)
    self._ck(lib.sdso_ba_get_energy_th(self.h, _ptr(th, _fp)))
    return th


Window.get_energy_th = _w_get_energy_th
Window.new_frame_energy_th = _w_new_frame_energy_th
Window.optimize = _w_optimize
Window.get_state = _w_get_state

lib.sdso_ba_marginalize_points.argtypes = [C.c_void_p]
lib.sdso_ba_marginalize_frame.argtypes = [C.c_void_p, C.c_int]
lib.sdso_ba_energies.argtypes = [C.c_void_p, _dp, _dp]


def _w_marginalize_points(self):
    self._ck(lib.sdso_ba_marginalize_points(self.h))


def _w_marginalize_frame(self, idx):
    self._ck(lib.sdso_ba_marginalize_frame(self.h, idx))


def _w_energies(self):
    m, l = C.c_double(), C.c_double()
    self._ck(lib.sdso_ba_energies(self.h, C.byref(m), C.byref(l)))
    return m.value, l.value


Window.marginalize_points = _w_marginalize_points
Window.marginalize_frame = _w_marginalize_frame
Window.energies = _w_energies

lib.sdso_lba_edge_eval.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _fp, _fp, _ip]


def _w_lba_edge_eval(self, T_wh, photo, idepth, cam, b0):
    c = self.counts()
    n, R = c["frames"], c["res"]
    T_wh, photo, idepth, cam, b0 = _f64(T_wh).reshape(n, 12), _f64(photo).reshape(n, 2), _f64(idepth).reshape(R), _f64(cam).reshape(4), _f64(b0).reshape(n)
    o = dict(error=np.zeros((R, 8)), J_xi=np.zeros((R, 8, 6)), J_photo=np.zeros((R, 8, 2)), J_idepth=np.zeros((R, 8)), J_C=np.zeros((R, 8, 4)),
             newState=np.zeros(R, np.int32), newEnergy=np.zeros(R), newEnergyWithOutlier=np.zeros(R), center=np.zeros((R, 3), np.float32),
             idepth_hessian=np.zeros(R, np.float32), level=np.zeros(R, np.int32))
    self._ck(lib.sdso_lba_edge_eval(self.h, _ptr(T_wh, _dp), _ptr(photo, _dp), _ptr(idepth, _dp), _ptr(cam, _dp), _ptr(b0, _dp),
                                    _ptr(o["error"], _dp), _ptr(o["J_xi"], _dp), _ptr(o["J_photo"], _dp), _ptr(o["J_idepth"], _dp), _ptr(o["J_C"], _dp),
                                    _ptr(o["newState"], _ip), _ptr(o["newEnergy"], _dp), _ptr(o["newEnergyWithOutlier"], _dp), _ptr(o["center"], _fp),
                                    _ptr(o["idepth_hessian"], _fp), _ptr(o["level"], _ip)))
    return o


Window.lba_edge_eval = _w_lba_edge_eval

lib.sdso_activate_points.argtypes = [C.c_void_p, C.c_int, _ip, C.c_void_p, C.c_int, C.c_int, _ip, _fp, _ip, _fp]


def _w_activate_points(self, host, pts, variant=VARIANT_SSE, min_obs=1):
    assert pts.dtype == IMMATURE_DTYPE and pts.flags["C_CONTIGUOUS"]
    n, nf = pts.size, self.counts()["frames"]
    host = np.ascontiguousarray(host, np.int32)
    res, st = np.zeros(n, np.int32), np.zeros((n, nf), np.int32)
    idp, en = np.zeros(n, np.float32), np.zeros(n, np.float32)
    self._ck(lib.sdso_activate_points(self.h, n, _ptr(host, _ip), pts.ctypes.data, variant, min_obs, _ptr(res, _ip), _ptr(idp, _fp), _ptr(st, _ip), _ptr(en, _fp)))
    return dict(result=res, idepth=idp, states=st, energy=en)


Window.activate_points = _w_activate_points

lib.sdso_lba_g2o.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip, _dp, _ip, _fp, _fp, _ip, _ip]


def _w_lba_g2o(self, cam, T_wh, photo, idepth, iters=3):
    c = self.counts()
    n, R = c["frames"], c["res"]
    cam, T_wh, photo, idepth = _f64(cam).copy(), _f64(T_wh).reshape(n, 12).copy(), _f64(photo).reshape(n, 2).copy(), _f64(idepth).reshape(R).copy()
    used, ns = np.zeros(n, np.int32), np.zeros(R, np.int32)
    chi2 = C.c_double()
    its, trials = C.c_int(), C.c_int()
    ce, ih = np.zeros((R, 3), np.float32), np.zeros(R, np.float32)
    self._ck(lib.sdso_lba_g2o(self.h, iters, _ptr(cam, _dp), _ptr(T_wh, _dp), _ptr(photo, _dp), _ptr(idepth, _dp), _ptr(used, _ip), C.byref(chi2),
                              _ptr(ns, _ip), _ptr(ce, _fp), _ptr(ih, _fp), C.byref(its), C.byref(trials)))
    return dict(iterations=its.value, trials=trials.value, cam=cam, T_wh=T_wh.reshape(n, 3, 4), photo=photo, idepth=idepth, used_host=used,
                chi2=chi2.value, newState=ns, center=ce, idepth_hessian=ih)


Window.lba_g2o = _w_lba_g2o


# ---------------------------------------------------------------------------------------------------
# candidate pixel selection (PixelSelector2.cpp)
_ubp = C.POINTER(C.c_ubyte)
lib.sdso_selector_pattern_host.argtypes = [C.c_uint, _ubp, C.c_size_t]
lib.sdso_selector_pattern_host.restype = None
lib.sdso_selector_reset.argtypes = [C.c_void_p]
lib.sdso_selector_random_pattern.argtypes = [C.c_void_p, _ubp]
lib.sdso_selector_potential.argtypes = [C.c_void_p, C.c_int, _ip]
lib.sdso_selector_make_hists.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
lib.sdso_selector_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, _fp, _ip]
lib.sdso_make_maps.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, _fp, _ip]
lib.sdso_selector_points.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _ip]


def selector_pattern_host(n, seed=3141592):
    out = np.zeros(n, np.uint8)
    lib.sdso_selector_pattern_host(seed, _ptr(out, _ubp), n)
    return out


def _selector_reset(self):
    self._ck(lib.sdso_selector_reset(self._h))


def _selector_random_pattern(self):
    w, h = self.level_size(0)
    out = np.zeros(w * h, np.uint8)
    self._ck(lib.sdso_selector_random_pattern(self._h, _ptr(out, _ubp)))
    return out


def _selector_potential(self, set=0):
    p = C.c_int(0)
    self._ck(lib.sdso_selector_potential(self._h, int(set), C.byref(p)))
    return p.value


def _selector_make_hists(self, fid):
    w, h = self.level_size(0)
    ths = np.zeros((h // 32, w // 32), np.float32)
    sm = np.zeros_like(ths)
    self._ck(lib.sdso_selector_make_hists(self._h, fid, _ptr(ths, _fp), _ptr(sm, _fp)))
    return ths, sm


def _selector_select(self, fid, pot, th_factor=1.0, want_map=True):
    w, h = self.level_size(0)
    m = np.zeros((h, w), np.float32) if want_map else None
    n = np.zeros(3, np.int32)
    self._ck(lib.sdso_selector_select(self._h, fid, int(pot), float(th_factor), _ptr(m, _fp) if want_map else None, _ptr(n, _ip)))
    return m, n


def _make_maps(self, fid, density=None, recursions_left=1, th_factor=1.0, want_map=True):
    """PixelSelector::makeMaps; density defaults to setting_desiredImmatureDensity (FullSystem.cpp:1605)."""
    w, h = self.level_size(0)
    if density is None:
        density = default_settings().desiredImmatureDensity
    m = np.zeros((h, w), np.float32) if want_map else None
    n = C.c_int(0)
    self._ck(lib.sdso_make_maps(self._h, fid, float(density), int(recursions_left), float(th_factor), _ptr(m, _fp) if want_map else None, C.byref(n)))
    return m, n.value


def _selector_points(self):
    w, h = self.level_size(0)
    uv = np.zeros((w * h, 2), np.float32)
    ty = np.zeros(w * h, np.float32)
    n = C.c_int(0)
    self._ck(lib.sdso_selector_points(self._h, w * h, _ptr(uv, _fp), _ptr(ty, _fp), C.byref(n)))
    return uv[:n.value].copy(), ty[:n.value].copy()


Context.selector_reset = _selector_reset
Context.selector_random_pattern = _selector_random_pattern
Context.selector_potential = _selector_potential
Context.selector_make_hists = _selector_make_hists
Context.selector_select = _selector_select
Context.make_maps = _make_maps
Context.selector_points = _selector_points


# ---------------------------------------------------------------------------------------------------
# coarse distance map + activation candidate filter (CoarseTracker.cpp:1216-1366, FullSystem.cpp:838-901)
lib.sdso_distmap_make.argtypes = [C.c_void_p, C.c_int, _fp, _fp, C.c_int, _ip, _fp, _fp]
lib.sdso_distmap_add.argtypes = [C.c_void_p, C.c_int, _ip, _fp]
lib.sdso_activation_filter.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _ubp, C.c_int, _ip, C.c_void_p, _fp, C.c_float, _ip, _ip, _fp]


def _distmap_make(self, KRKi, Kt, pt_host, pt_uvid, want_map=True):
    w1, h1 = self.level_size(1)
    K_, t_ = _f32(KRKi).reshape(-1, 9), _f32(Kt).reshape(-1, 3)
    ph = np.ascontiguousarray(pt_host, dtype=np.int32)
    pv = _f32(pt_uvid).reshape(-1, 3)
    m = np.zeros((h1, w1), np.float32) if want_map else None
    self._ck(lib.sdso_distmap_make(self._h, K_.shape[0], _ptr(K_, _fp), _ptr(t_, _fp), ph.size, _ptr(ph, _ip), _ptr(pv, _fp),
                                   _ptr(m, _fp) if want_map else None))
    return m


def _distmap_add(self, uv):
    w1, h1 = self.level_size(1)
    uv = np.ascontiguousarray(uv, dtype=np.int32).reshape(-1, 2)
    m = np.zeros((h1, w1), np.float32)
    self._ck(lib.sdso_distmap_add(self._h, uv.shape[0], _ptr(uv, _ip), _ptr(m, _fp)))
    return m


def _activation_filter(self, KRKi, Kt, host_flagged, cand_host, pts, my_type, current_min_act_dist, want_map=True):
    w1, h1 = self.level_size(1)
    K_, t_ = _f32(KRKi).reshape(-1, 9), _f32(Kt).reshape(-1, 3)
    fl = np.ascontiguousarray(host_flagged, dtype=np.uint8)
    ch = np.ascontiguousarray(cand_host, dtype=np.int32)
    ty = _f32(my_type)
    pts = np.ascontiguousarray(pts)
    verdict = np.zeros(ch.size, np.int32)
    rounds = C.c_int(0)
    m = np.zeros((h1, w1), np.float32) if want_map else None
    self._ck(lib.sdso_activation_filter(self._h, K_.shape[0], _ptr(K_, _fp), _ptr(t_, _fp), _ptr(fl, _ubp), ch.size, _ptr(ch, _ip), pts.ctypes.data,
                                        _ptr(ty, _fp), float(current_min_act_dist), _ptr(verdict, _ip), C.byref(rounds),
                                        _ptr(m, _fp) if want_map else None))
    return verdict, m, rounds.value


Context.distmap_make = _distmap_make
Context.distmap_add = _distmap_add
Context.activation_filter = _activation_filter


# ---------------------------------------------------------------------------------------------------
# input preparation + trajectory rows (Undistort.cpp:222-260, 398-489; FullSystem.cpp:236-285)
lib.sdso_undistort_setup.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int]
lib.sdso_undistort.argtypes = [C.c_void_p, _ubp, C.c_float, C.c_float, _fp, C.c_int, C.c_int, _fp]
lib.sdso_trajectory_row.argtypes = [_dp, C.c_char_p, C.c_int]


def _undistort_setup(self, w_org, h_org, remap_x, remap_y, G=None, vignette_inv=None, photometric_calibration=2, use_exposure=True):
    rx, ry = _f32(remap_x), _f32(remap_y)
    g = _f32(G) if G is not None else None
    v = _f32(vignette_inv) if vignette_inv is not None else None
    self._ck(lib.sdso_undistort_setup(self._h, w_org, h_org, _ptr(rx, _fp), _ptr(ry, _fp), _ptr(g, _fp) if g is not None else None,
                                      _ptr(v, _fp) if v is not None else None, int(photometric_calibration), int(use_exposure)))


def _undistort(self, raw, exposure=1.0, factor=1.0, frame=-1, use_hcalib=True, want_image=True):
    w, h = self.level_size(0)
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    out = np.zeros((h, w), np.float32) if want_image else None
    e = C.c_float(0)
    self._ck(lib.sdso_undistort(self._h, _ptr(raw, _ubp), float(exposure), float(factor), _ptr(out, _fp) if want_image else None, int(frame),
                                int(use_hcalib), C.byref(e)))
    return out, e.value


def trajectory_row(T):
    T = _f64(T).reshape(12)
    buf = C.create_string_buffer(512)
    n = lib.sdso_trajectory_row(_ptr(T, _dp), buf, 512)
    if n < 0:
        raise RuntimeError("sdso_trajectory_row failed")
    return buf.value.decode()


Context.undistort_setup = _undistort_setup
Context.undistort = _undistort

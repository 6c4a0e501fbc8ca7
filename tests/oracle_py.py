"""ctypes view of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(_ROOT, "oracle")
# SDSO_ORACLE_LIB: bench.py points this at the -march=native build it makes on the box it times the CPU arm on
LIB_PATH = os.environ.get("SDSO_ORACLE_LIB") or os.path.join(ORACLE_DIR, "_build", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp"))]
    if (not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _load():
    if not os.path.exists(LIB_PATH):
        if os.environ.get("SDSO_ORACLE_LIB"):
            raise ImportError(f"SDSO_ORACLE_LIB={LIB_PATH} does not exist")
        build()
    return C.CDLL(LIB_PATH)


lib = _load()
_dp, _fp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


lib.orc_create.restype = C.c_void_p
lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
lib.orc_destroy.argtypes = [C.c_void_p]
lib.orc_levels.argtypes = [C.c_void_p]
lib.orc_level_size.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
lib.orc_level_K.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
lib.orc_set_affine_opt_mode.argtypes = [C.c_void_p, C.c_float, C.c_float]
lib.orc_frame_new.argtypes = [C.c_void_p]
lib.orc_make_images.argtypes = [C.c_void_p, C.c_int, _fp, C.c_float, C.c_int]
lib.orc_frame_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, _fp]
lib.orc_interp33.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_int, _fp]
lib.orc_interp33bilin.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_int, _fp]
lib.orc_tracker_makeK.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
lib.orc_tracker_set_ref.argtypes = [C.c_void_p, C.c_int, _fp, C.c_int, _dp]
lib.orc_tracker_pc.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp, _fp]
lib.orc_tracker_set_pc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _dp]
lib.orc_calc_res_sse.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_float, _dp, _ip]
lib.orc_get_warped.argtypes = [C.c_void_p, _fp]
lib.orc_calc_gs_sse.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
lib.orc_track_sse.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _ip]
lib.orc_track_g2o.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp, _dp, _ip]
lib.orc_edge_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _ip]
lib.orc_evals.restype = C.c_ulonglong
lib.orc_evals.argtypes = [C.c_void_p]
lib.orc_reset_evals.argtypes = [C.c_void_p]
lib.orc_g2o_trial_counts.argtypes = [C.c_void_p, _ip]
for name in ("orc_se3_exp", "orc_se3_log", "orc_se3_adj", "orc_se3_inv"):
    getattr(lib, name).argtypes = [_dp, _dp]
lib.orc_se3_mul.argtypes = [_dp, _dp, _dp]
lib.orc_ldlt_solve.argtypes = [C.c_int, _dp, _dp, _dp]


lib.orc_vertex_oplus.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
lib.orc_edge_trace_uv.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _fp, _dp, _fp, _dp, _dp, _dp, _ip]


def vertex_oplus(kind, estimate, update, aux=None):
    """V1-V5 oplusImpl (dso_g2o_vertex.cpp); kind 1 pose, 2 photometric, 3 inverse depth, 4 uv, 5 camera."""
    est = np.array(estimate, np.float64, order="C")
    width = {1: 12, 2: 2, 3: 1, 4: 2, 5: 4}[kind]
    upd = _f64(update)
    a = _f64(aux) if aux is not None else None
    lib.orc_vertex_oplus(kind, est.size // width, _p(est, _dp), _p(upd, _dp), _p(a, _dp) if a is not None else None)
    return est


def edge_trace_uv(orc, fid, uv, rot, meas, aff, dxdy, error=None, J=None):
    """E3 computeError + linearizeOplus (dso_g2o_edge.cpp:571-619)."""
    uv_, rot_, me_, dx_ = _f64(uv).reshape(-1, 2), _f32(rot).reshape(-1, 2), _f64(meas).reshape(-1), _f64(dxdy).reshape(-1, 2)
    n = uv_.shape[0]
    a_ = _f32(aff).reshape(2)
    err = np.zeros(n) if error is None else np.array(error, np.float64)
    Jo = np.zeros(n) if J is None else np.array(J, np.float64)
    flag = np.zeros(n, np.int32)
    lib.orc_edge_trace_uv(orc._h, fid, n, _p(uv_, _dp), _p(rot_, _fp), _p(me_, _dp), _p(a_, _fp), _p(dx_, _dp), _p(err, _dp), _p(Jo, _dp), _p(flag, _ip))
    return err, Jo, flag


def se3_exp(a):
    a = _f64(a)
    T = np.zeros(12)
    lib.orc_se3_exp(_p(a, _dp), _p(T, _dp))
    return T.reshape(3, 4)


def se3_log(T):
    T = _f64(T).reshape(12)
    a = np.zeros(6)
    lib.orc_se3_log(_p(T, _dp), _p(a, _dp))
    return a


def se3_adj(T):
    T = _f64(T).reshape(12)
    A = np.zeros(36)
    lib.orc_se3_adj(_p(T, _dp), _p(A, _dp))
    return A.reshape(6, 6)


def se3_mul(A, B):
    A, B = _f64(A).reshape(12), _f64(B).reshape(12)
    Cm = np.zeros(12)
    lib.orc_se3_mul(_p(A, _dp), _p(B, _dp), _p(Cm, _dp))
    return Cm.reshape(3, 4)


def se3_inv(A):
    A = _f64(A).reshape(12)
    B = np.zeros(12)
    lib.orc_se3_inv(_p(A, _dp), _p(B, _dp))
    return B.reshape(3, 4)


def ldlt_solve(A, b):
    A, b = _f64(A), _f64(b)
    x = np.zeros_like(b)
    lib.orc_ldlt_solve(b.size, _p(A, _dp), _p(b, _dp), _p(x, _dp))
    return x


class Oracle:
    def __init__(self, w, h, K, baseline=0.0):
        self._h = C.c_void_p(lib.orc_create(w, h, K[0], K[1], K[2], K[3], baseline))
        self.w, self.h = w, h
        self.levels = lib.orc_levels(self._h)

    def close(self):
        if self._h:
            lib.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level_size(self, lvl):
        w, h = C.c_int(), C.c_int()
        lib.orc_level_size(self._h, lvl, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_K(self, lvl):
        K, Ki = np.zeros(9, np.float32), np.zeros(9, np.float32)
        lib.orc_level_K(self._h, lvl, _p(K, _fp), _p(Ki, _fp))
        return K.reshape(3, 3), Ki.reshape(3, 3)

    def set_affine_opt_mode(self, a, b):
        lib.orc_set_affine_opt_mode(self._h, a, b)

    def frame_new(self):
        return lib.orc_frame_new(self._h)

    def make_images(self, fid, image, exposure=1.0, use_hcalib=True):
        img = _f32(image)
        lib.orc_make_images(self._h, fid, _p(img, _fp), exposure, int(use_hcalib))

    def frame_get(self, fid, lvl):
        w, h = self.level_size(lvl)
        dI, ag = np.zeros((h, w, 3), np.float32), np.zeros((h, w), np.float32)
        lib.orc_frame_get(self._h, fid, lvl, _p(dI, _fp), _p(ag, _fp))
        return dI, ag

    def interp33(self, fid, lvl, xy, bilin=False):
        xy = _f32(xy).reshape(-1, 2)
        out = np.zeros((xy.shape[0], 3), np.float32)
        (lib.orc_interp33bilin if bilin else lib.orc_interp33)(self._h, fid, lvl, _p(xy, _fp), xy.shape[0], _p(out, _fp))
        return out

    def tracker_make_k(self, K):
        lib.orc_tracker_makeK(self._h, K[0], K[1], K[2], K[3])

    def tracker_set_ref(self, fid, uvidw, aff=(0.0, 0.0)):
        p = _f32(uvidw).reshape(-1, 4)
        a = _f64(aff)
        lib.orc_tracker_set_ref(self._h, fid, _p(p, _fp), p.shape[0], _p(a, _dp))

    def tracker_get_pc(self, lvl):
        w, h = self.level_size(lvl)
        u, v, idp, col = (np.zeros(w * h, np.float32) for _ in range(4))
        n = lib.orc_tracker_pc(self._h, lvl, _p(u, _fp), _p(v, _fp), _p(idp, _fp), _p(col, _fp))
        return u[:n].copy(), v[:n].copy(), idp[:n].copy(), col[:n].copy()

    def tracker_set_pc(self, fid, lvl, u, v, idepth, color, aff=(0.0, 0.0)):
        u, v, idepth, color = _f32(u), _f32(v), _f32(idepth), _f32(color)
        a = _f64(aff)
        lib.orc_tracker_set_pc(self._h, fid, lvl, u.size, _p(u, _fp), _p(v, _fp), _p(idepth, _fp), _p(color, _fp), _p(a, _dp))

    def calc_res_gs(self, new_fid, lvl, T, aff, cutoff):
        T, aff = _f64(T).reshape(12), _f64(aff)
        rs, H, b = np.zeros(6), np.zeros(64), np.zeros(8)
        wn = C.c_int()
        lib.orc_calc_res_sse(self._h, new_fid, lvl, _p(T, _dp), _p(aff, _dp), cutoff, _p(rs, _dp), C.byref(wn))
        warped = np.zeros((8, wn.value), np.float32)
        lib.orc_get_warped(self._h, _p(warped, _fp))
        lib.orc_calc_gs_sse(self._h, lvl, _p(T, _dp), _p(aff, _dp), _p(H, _dp), _p(b, _dp))
        return dict(rs=rs, H=H.reshape(8, 8), b=b, warped_n=wn.value, warped=warped)

    def track(self, new_fid, T, aff, coarsest, min_res, variant=0):
        T, aff, mr = _f64(T).reshape(12).copy(), _f64(aff).copy(), _f64(min_res)
        lr, fl, it = np.zeros(5), np.zeros(3), np.zeros(5, np.int32)
        fn = lib.orc_track_sse if variant == 0 else lib.orc_track_g2o
        ok = fn(self._h, new_fid, _p(T, _dp), _p(aff, _dp), coarsest, _p(mr, _dp), _p(lr, _dp), _p(fl, _dp), _p(it, _ip))
        return dict(T=T.reshape(3, 4), aff=aff, lastResiduals=lr, flow=fl, iterations=it, ok=bool(ok))

    def edge_eval(self, new_fid, lvl, T_select, T_pose, photo):
        Ts, Tp, ph = _f64(T_select).reshape(12), _f64(T_pose).reshape(12), _f64(photo)
        w, h = self.level_size(lvl)
        err, J = np.zeros(w * h), np.zeros((w * h, 8))
        n = lib.orc_edge_eval(self._h, new_fid, lvl, _p(Ts, _dp), _p(Tp, _dp), _p(ph, _dp), _p(err, _dp), _p(J, _dp), None)
        return err[:n].copy(), J[:n].copy()

    def g2o_trial_counts(self):
        """(damping trials, rejected trials) of the last g2o track call"""
        o = np.zeros(2, np.int32)
        lib.orc_g2o_trial_counts(self._h, _p(o, _ip))
        return int(o[0]), int(o[1])

    def evals(self):
        return int(lib.orc_evals(self._h))

    def reset_evals(self):
        lib.orc_reset_evals(self._h)

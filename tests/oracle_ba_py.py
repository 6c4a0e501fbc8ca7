"""ctypes view of the oracle's windowed-BA entry points (oracle/c_api.cpp, orc_ba_*) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import numpy as np
import oracle_py as O

lib = O.lib
_dp, _fp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
V = C.c_void_p
lib.orc_immature_init.argtypes = [V, C.c_int, C.c_float, C.c_float, _fp, _fp, _fp, _fp]
lib.orc_ba_reset.argtypes = [V]
lib.orc_ba_set_calib_delta.argtypes = [V, _dp]
lib.orc_ba_add_frame.argtypes = [V, C.c_int, _dp, C.c_double, C.c_double, C.c_int]
lib.orc_ba_set_state.argtypes = [V, C.c_int, _dp]
lib.orc_ba_set_energy_th.argtypes = [V, C.c_int, C.c_float]
lib.orc_ba_add_point.argtypes = [V, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _fp, _fp, C.c_int]
lib.orc_ba_add_residual.argtypes = [V, C.c_int, C.c_int]
lib.orc_ba_add_points.argtypes = [V, C.c_int, _ip, _fp, _fp, _fp, _fp, _fp, _fp, C.POINTER(C.c_ubyte)]
lib.orc_ba_add_residuals.argtypes = [V, C.c_int, _ip, _ip]
lib.orc_ba_set_point_flag.argtypes = [V, C.c_int, C.c_int]
lib.orc_ba_prepare.argtypes = [V]
lib.orc_ba_counts.argtypes = [V, _ip, _ip, _ip]
lib.orc_ba_precalc.argtypes = [V, C.c_int, C.c_int, _fp]
lib.orc_ba_adjoints.argtypes = [V, _dp, _dp, _fp]
lib.orc_ba_linearize_all.restype = C.c_double
lib.orc_ba_linearize_all.argtypes = [V, C.c_int]
lib.orc_ba_apply_res.argtypes = [V, C.c_int]
lib.orc_ba_fix_linearization.argtypes = [V, C.c_int]
lib.orc_ba_get_res.argtypes = [V, C.c_int, _ip, _ip, _dp, _dp, _ip, _fp, _fp, _fp, _fp]
lib.orc_ba_get_points.argtypes = [V, _fp]
lib.orc_ba_accumulate_top.argtypes = [V, C.c_int, C.c_int, _dp, _dp, _fp]
lib.orc_ba_accumulate_sc.argtypes = [V, C.c_int, _dp, _dp]
lib.orc_ba_solve.argtypes = [V, C.c_int, C.c_double, _dp, _dp, _dp]
lib.orc_ba_resubstitute.argtypes = [V, _dp, _dp, _dp]
lib.orc_ba_set_marg_prior.argtypes = [V, _dp, _dp]
lib.orc_ba_get_marg_prior.argtypes = [V, _dp, _dp]
lib.orc_ba_marginalize_points.argtypes = [V]
lib.orc_ba_marginalize_frame.argtypes = [V, C.c_int]
lib.orc_ba_orthogonalize.argtypes = [V, _dp, _dp]
lib.orc_ba_energies.restype = C.c_double
lib.orc_ba_energies.argtypes = [V, _dp]
lib.orc_ba_nullspaces.argtypes = [V, _dp]
lib.orc_ba_set_reduce.argtypes = [V, C.c_int, C.c_uint]
lib.orc_ba_get_energy_th.argtypes = [V, _fp]

_p = O._p
_f32, _f64 = O._f32, O._f64


def immature_init(orc, fid, u, v):
    """D1: colour[8], weights[8], gradH(2x2), energyTH of the pattern around (u, v) (ImmaturePoint.cpp:33-88)."""
    col, wts, gH, eth = np.zeros(8, np.float32), np.zeros(8, np.float32), np.zeros(4, np.float32), C.c_float()
    ok = lib.orc_immature_init(orc._h, fid, u, v, _p(col, _fp), _p(wts, _fp), _p(gH, _fp), C.byref(eth))
    return bool(ok), col, wts, gH.reshape(2, 2), eth.value


class OracleBA:
    """Index-based mirror of the reference's window: frames, points, residuals (oracle/oracle_ba.hpp)."""

    def __init__(self, orc):
        self.orc = orc
        self.h = orc._h
        lib.orc_ba_reset(self.h)

    def add_frame(self, fid, T_w2c, a=0.0, b=0.0, frameID=1):
        T = _f64(T_w2c).reshape(12)
        return lib.orc_ba_add_frame(self.h, fid, _p(T, _dp), a, b, frameID)

    def set_state(self, idx, state10):
        s = _f64(state10)
        lib.orc_ba_set_state(self.h, idx, _p(s, _dp))

    def set_energy_th(self, idx, th):
        lib.orc_ba_set_energy_th(self.h, idx, th)

    def set_calib_delta(self, d4):
        d = _f64(d4)
        lib.orc_ba_set_calib_delta(self.h, _p(d, _dp))

    def add_point(self, host, u, v, idepth, idepth_zero, color, weights, has_prior=False):
        c, w = _f32(color), _f32(weights)
        return lib.orc_ba_add_point(self.h, host, u, v, idepth, idepth_zero, _p(c, _fp), _p(w, _fp), int(has_prior))

    def set_points(self, host, u, v, idepth, idepth_zero, color8, weights8, has_prior):
        """batched add_point (same argument layout as the device Window.set_points)"""
        host = np.ascontiguousarray(host, np.int32)
        u, v, idepth, idepth_zero = _f32(u), _f32(v), _f32(idepth), _f32(idepth_zero)
        c, w = _f32(color8).reshape(-1, 8), _f32(weights8).reshape(-1, 8)
        hp = np.ascontiguousarray(has_prior, np.uint8)
        lib.orc_ba_add_points(self.h, host.size, _p(host, _ip), _p(u, _fp), _p(v, _fp), _p(idepth, _fp), _p(idepth_zero, _fp), _p(c, _fp), _p(w, _fp),
                              hp.ctypes.data_as(C.POINTER(C.c_ubyte)))

    def set_residuals(self, point, target):
        p, t = np.ascontiguousarray(point, np.int32), np.ascontiguousarray(target, np.int32)
        lib.orc_ba_add_residuals(self.h, p.size, _p(p, _ip), _p(t, _ip))

    def add_residual(self, pidx, target):
        return lib.orc_ba_add_residual(self.h, pidx, target)

    def set_point_flag(self, pidx, flag):
        lib.orc_ba_set_point_flag(self.h, pidx, flag)

    def prepare(self):
        lib.orc_ba_prepare(self.h)

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        d = lib.orc_ba_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(frames=a.value, points=b.value, res=c.value, dim=d)

    def precalc(self, h, t):
        out = np.zeros(49, np.float32)
        lib.orc_ba_precalc(self.h, h, t, _p(out, _fp))
        return out

    def adjoints(self):
        n = self.counts()["frames"]
        ah, at, d = np.zeros((n * n, 8, 8)), np.zeros((n * n, 8, 8)), np.zeros((n * n, 8), np.float32)
        lib.orc_ba_adjoints(self.h, _p(ah, _dp), _p(at, _dp), _p(d, _fp))
        return ah, at, d

    def linearize_all(self, fix=False):
        return lib.orc_ba_linearize_all(self.h, int(fix))

    def apply_res(self, copy=True):
        lib.orc_ba_apply_res(self.h, int(copy))

    def fix_linearization(self, ridx):
        lib.orc_ba_fix_linearization(self.h, ridx)

    def get_res(self, which=0, brief=False):
        R = self.counts()["res"]
        ns, st, ne, nw, ac = np.zeros(R, np.int32), np.zeros(R, np.int32), np.zeros(R), np.zeros(R), np.zeros(R, np.int32)
        J, jp, ce, rz = np.zeros((R, 74), np.float32), np.zeros((R, 8), np.float32), np.zeros((R, 3), np.float32), np.zeros((R, 8), np.float32)
        lib.orc_ba_get_res(self.h, which, _p(ns, _ip), _p(st, _ip), _p(ne, _dp), _p(nw, _dp), _p(ac, _ip), _p(J, _fp), _p(jp, _fp), _p(ce, _fp), _p(rz, _fp))
        return dict(newState=ns, state=st, newEnergy=ne, newEnergyWithOutlier=nw, active=ac, J=J, JpJdF=jp, center=ce, res_toZero=rz)

    def get_points(self):
        P = self.counts()["points"]
        o = np.zeros((P, 16), np.float32)
        lib.orc_ba_get_points(self.h, _p(o, _fp))
        return dict(Hdd_A=o[:, 0], bd_A=o[:, 1], Hcd_A=o[:, 2:6], Hdd_L=o[:, 6], bd_L=o[:, 7], Hcd_L=o[:, 8:12], HdiF=o[:, 12], bdSumF=o[:, 13], step=o[:, 14], priorF=o[:, 15])

    def accumulate_top(self, mode, use_prior):
        c = self.counts()
        d, n = c["dim"], c["frames"]
        H, b, blk = np.zeros((d, d)), np.zeros(d), np.zeros((n * n, 13, 13), np.float32)
        lib.orc_ba_accumulate_top(self.h, mode, int(use_prior), _p(H, _dp), _p(b, _dp), _p(blk, _fp))
        return H, b, blk

    def accumulate_sc(self, shift=True):
        d = self.counts()["dim"]
        H, b = np.zeros((d, d)), np.zeros(d)
        lib.orc_ba_accumulate_sc(self.h, int(shift), _p(H, _dp), _p(b, _dp))
        return H, b

    def get_energy_th(self):
        th = np.zeros(self.counts()["frames"], np.float32)
        lib.orc_ba_get_energy_th(self.h, _p(th, _fp))
        return th

    def set_reduce(self, threads=1, seed=0):
        """Worker partition of the float accumulators (the reference's NUM_THREADS=6 IndexThreadReduce; any seed is an assignment the
        reference's dynamic chunk queue can produce). (1, 0) = the single-threaded path."""
        lib.orc_ba_set_reduce(self.h, threads, seed)

    def solve(self, iteration, lam=1e-5):
        d = self.counts()["dim"]
        x, H, b = np.zeros(d), np.zeros((d, d)), np.zeros(d)
        lib.orc_ba_solve(self.h, iteration, lam, _p(x, _dp), _p(H, _dp), _p(b, _dp))
        return x, H, b

    def resubstitute(self, x):
        n = self.counts()["frames"]
        x = _f64(x)
        fs, cs = np.zeros((n, 10)), np.zeros(4)
        lib.orc_ba_resubstitute(self.h, _p(x, _dp), _p(fs, _dp), _p(cs, _dp))
        return fs, cs

    def set_marg_prior(self, HM, bM):
        HM, bM = _f64(HM), _f64(bM)
        lib.orc_ba_set_marg_prior(self.h, _p(HM, _dp), _p(bM, _dp))

    def get_marg_prior(self):
        d = self.counts()["dim"]
        HM, bM = np.zeros((d, d)), np.zeros(d)
        lib.orc_ba_get_marg_prior(self.h, _p(HM, _dp), _p(bM, _dp))
        return HM, bM

    def marginalize_points(self):
        lib.orc_ba_marginalize_points(self.h)

    def marginalize_frame(self, idx):
        lib.orc_ba_marginalize_frame(self.h, idx)

    def orthogonalize(self, b=None, H=None):
        bb = _f64(b).copy() if b is not None else None
        HH = _f64(H).copy() if H is not None else None
        lib.orc_ba_orthogonalize(self.h, _p(bb, _dp) if bb is not None else None, _p(HH, _dp) if HH is not None else None)
        return bb, HH

    def nullspaces(self):
        d = self.counts()["dim"]
        N = np.zeros((d, 7))
        lib.orc_ba_nullspaces(self.h, _p(N, _dp))
        return N

lib.orc_ba_optimize.restype = C.c_double
lib.orc_ba_optimize.argtypes = [V, C.c_int, _ip]
lib.orc_ba_new_frame_energy_th.restype = C.c_float
lib.orc_ba_new_frame_energy_th.argtypes = [V]
lib.orc_ba_get_state.argtypes = [V, _dp, _dp, _fp, _dp]


def _optimize(self, iters=6):
    done = C.c_int()
    rmse = lib.orc_ba_optimize(self.h, iters, C.byref(done))
    return rmse, done.value


def _new_frame_energy_th(self):
    return lib.orc_ba_new_frame_energy_th(self.h)


def _get_state(self):
    c = self.counts()
    st, T, idp, cal = np.zeros((c["frames"], 10)), np.zeros((c["frames"], 3, 4)), np.zeros(c["points"], np.float32), np.zeros(4)
    lib.orc_ba_get_state(self.h, _p(st, _dp), _p(T, _dp), _p(idp, _fp), _p(cal, _dp))
    return dict(states=st, T_w2c=T, idepth=idp, calib=cal)


OracleBA.optimize = _optimize
OracleBA.new_frame_energy_th = _new_frame_energy_th
OracleBA.get_state = _get_state


def _energies(self):
    """(calcMEnergyF, calcLEnergyF)"""
    l = C.c_double()
    m = lib.orc_ba_energies(self.h, C.byref(l))
    return m, l.value


OracleBA.energies = _energies

lib.orc_lba_edge_eval.argtypes = [V, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _fp, _fp, _ip]


def _lba_edge_eval(self, T_wh, photo, idepth, cam, b0):
    c = self.counts()
    n, R = c["frames"], c["res"]
    T_wh, photo, idepth, cam, b0 = _f64(T_wh).reshape(n, 12), _f64(photo).reshape(n, 2), _f64(idepth).reshape(R), _f64(cam).reshape(4), _f64(b0).reshape(n)
    o = dict(error=np.zeros((R, 8)), J_xi=np.zeros((R, 8, 6)), J_photo=np.zeros((R, 8, 2)), J_idepth=np.zeros((R, 8)), J_C=np.zeros((R, 8, 4)),
             newState=np.zeros(R, np.int32), newEnergy=np.zeros(R), newEnergyWithOutlier=np.zeros(R), center=np.zeros((R, 3), np.float32),
             idepth_hessian=np.zeros(R, np.float32), level=np.zeros(R, np.int32))
    lib.orc_lba_edge_eval(self.h, _p(T_wh, _dp), _p(photo, _dp), _p(idepth, _dp), _p(cam, _dp), _p(b0, _dp), _p(o["error"], _dp), _p(o["J_xi"], _dp),
                          _p(o["J_photo"], _dp), _p(o["J_idepth"], _dp), _p(o["J_C"], _dp), _p(o["newState"], _ip), _p(o["newEnergy"], _dp),
                          _p(o["newEnergyWithOutlier"], _dp), _p(o["center"], _fp), _p(o["idepth_hessian"], _fp), _p(o["level"], _ip))
    return o


OracleBA.lba_edge_eval = _lba_edge_eval

lib.orc_lba_g2o.restype = C.c_int
lib.orc_lba_g2o.argtypes = [V, C.c_int, _dp, _dp, _dp, _dp, _ip, _dp, _ip, _fp, _fp, _ip]


def _lba_g2o(self, cam, T_wh, photo, idepth, iters=3):
    """FullSystem::optimize g2o body with the restated g2o LM. Returns the final vertex estimates and bookkeeping."""
    c = self.counts()
    n, R = c["frames"], c["res"]
    cam, T_wh, photo, idepth = _f64(cam).copy(), _f64(T_wh).reshape(n, 12).copy(), _f64(photo).reshape(n, 2).copy(), _f64(idepth).reshape(R).copy()
    used, ns = np.zeros(n, np.int32), np.zeros(R, np.int32)
    chi2, trials = C.c_double(), C.c_int()
    ce, ih = np.zeros((R, 3), np.float32), np.zeros(R, np.float32)
    its = lib.orc_lba_g2o(self.h, iters, _p(cam, _dp), _p(T_wh, _dp), _p(photo, _dp), _p(idepth, _dp), _p(used, _ip), C.byref(chi2), _p(ns, _ip),
                          _p(ce, _fp), _p(ih, _fp), C.byref(trials))
    return dict(iterations=its, trials=trials.value, cam=cam, T_wh=T_wh.reshape(n, 3, 4), photo=photo, idepth=idepth, used_host=used, chi2=chi2.value,
                newState=ns, center=ce, idepth_hessian=ih)


OracleBA.lba_g2o = _lba_g2o

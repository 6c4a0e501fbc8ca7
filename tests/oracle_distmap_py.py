"""ctypes view of the oracle's coarse distance map / activation candidate filter (oracle/oracle_distmap.hpp) — TEST
INFRASTRUCTURE ONLY — plus synthetic inputs and a numpy restatement of the closed forms the device path relies on."""
import ctypes as C
import numpy as np
import oracle_py as O
import oracle_trace_py as T

lib = O.lib
_fp, _ip, _ubp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
lib.orc_dm_create.restype = C.c_void_p
lib.orc_dm_create.argtypes = [C.c_void_p]
lib.orc_dm_destroy.argtypes = [C.c_void_p]
lib.orc_dm_make.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _ip, _fp]
lib.orc_dm_add.argtypes = [C.c_void_p, C.c_int, _ip]
lib.orc_dm_get.argtypes = [C.c_void_p, _fp]
lib.orc_dm_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _fp, _fp, _ubp, C.c_int, _ip, C.c_void_p, _fp, C.c_float, _ip]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class DistMap:
    def __init__(self, orc):
        self.orc = orc
        self.w1, self.h1 = orc.level_size(1)
        self._h = C.c_void_p(lib.orc_dm_create(orc._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.orc_dm_destroy(self._h)
            self._h = None

    def get(self):
        m = np.zeros((self.h1, self.w1), np.float32)
        lib.orc_dm_get(self._h, m.ctypes.data_as(_fp))
        return m

    def make(self, KRKi, Kt, pt_host, pt_uvid):
        K_, t_ = _f32(KRKi).reshape(-1, 9), _f32(Kt).reshape(-1, 3)
        ph = np.asarray(pt_host, np.int32)
        assert (np.diff(ph) >= 0).all(), "points must be grouped by host"
        cnt = np.ascontiguousarray(np.bincount(ph, minlength=K_.shape[0]), dtype=np.int32)
        pv = _f32(pt_uvid).reshape(-1, 3)
        lib.orc_dm_make(self._h, K_.shape[0], K_.ctypes.data_as(_fp), t_.ctypes.data_as(_fp), cnt.ctypes.data_as(_ip), pv.ctypes.data_as(_fp))
        return self.get()

    def add(self, uv):
        uv = np.ascontiguousarray(uv, dtype=np.int32).reshape(-1, 2)
        lib.orc_dm_add(self._h, uv.shape[0], uv.ctypes.data_as(_ip))
        return self.get()

    def filter(self, KRKi, Kt, host_flagged, cand_host, pts, my_type, min_act_dist):
        K_, t_ = _f32(KRKi).reshape(-1, 9), _f32(Kt).reshape(-1, 3)
        fl = np.ascontiguousarray(host_flagged, dtype=np.uint8)
        ch = np.ascontiguousarray(cand_host, dtype=np.int32)
        ty = _f32(my_type)
        pts = np.ascontiguousarray(pts)
        verdict = np.zeros(ch.size, np.int32)
        lib.orc_dm_filter(self._h, self.orc._h, K_.shape[0], K_.ctypes.data_as(_fp), t_.ctypes.data_as(_fp), fl.ctypes.data_as(_ubp), ch.size,
                          ch.ctypes.data_as(_ip), pts.ctypes.data, ty.ctypes.data_as(_fp), float(min_act_dist), verdict.ctypes.data_as(_ip))
        return verdict, self.get()


def make_inputs(orc, seed, n_hosts=6, n_pts=1500, n_cand=4000, motion=0.05):
    """hosts with small relative motion to the newest frame, active points and immature candidates in level-0 pixels"""
    rng = np.random.default_rng(seed)
    w, h = orc.level_size(0)
    K1, _ = orc.level_K(1)
    _, Ki0 = orc.level_K(0)
    KRKi, Kt = [], []
    for _ in range(n_hosts):
        T_ = O.se3_exp(np.concatenate([rng.normal(0, motion, 3), rng.normal(0, motion * 0.3, 3)]))
        R, t = T_[:3, :3].astype(np.float32), T_[:3, 3].astype(np.float32)
        KRKi.append(((K1 @ R) @ Ki0).astype(np.float32).reshape(9))
        Kt.append((K1 @ t).astype(np.float32))
    KRKi, Kt = np.array(KRKi, np.float32), np.array(Kt, np.float32)
    pt_host = np.sort(rng.integers(0, n_hosts, n_pts)).astype(np.int32)
    pt_uvid = np.stack([rng.uniform(-20, w + 20, n_pts), rng.uniform(-20, h + 20, n_pts), rng.uniform(0.05, 2.0, n_pts)], 1).astype(np.float32)
    pts = np.zeros(n_cand, T.DTYPE)
    pts["u"] = rng.uniform(-10, w + 10, n_cand)
    pts["v"] = rng.uniform(-10, h + 10, n_cand)
    idmin = rng.uniform(0.0, 1.5, n_cand)
    pts["idepth_min"] = idmin
    pts["idepth_max"] = idmin + rng.uniform(0.0, 0.5, n_cand)
    pts["idepth_max"][rng.random(n_cand) < 0.05] = np.nan
    pts["idepth_min"][rng.random(n_cand) < 0.02] = -3.0
    pts["quality"] = rng.uniform(1.0, 12.0, n_cand)
    pts["lastTracePixelInterval"] = rng.uniform(0.0, 10.0, n_cand)
    pts["lastTraceStatus"] = rng.choice([0, 0, 0, 0, 1, 2, 3, 4, 5], n_cand)
    cand_host = np.sort(rng.integers(0, n_hosts, n_cand)).astype(np.int32)
    my_type = rng.choice([1.0, 1.0, 2.0, 4.0], n_cand).astype(np.float32)
    flagged = (rng.random(n_hosts) < 0.3).astype(np.uint8)
    return dict(KRKi=KRKi, Kt=Kt, pt_host=pt_host, pt_uvid=pt_uvid, pts=pts, cand_host=cand_host, my_type=my_type, flagged=flagged)


# ---- the closed forms the device path uses ----------------------------------------------------------------------------
def steps(dx, dy):
    dx, dy = abs(int(dx)), abs(int(dy))
    k = max(dx, dy)
    while k + (k + 1) // 2 < dx + dy:
        k += 1
    return k


def is_border(x, y, w1, h1):
    return x == 0 or y == 0 or x == w1 - 1 or y == h1 - 1


NB8 = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]


def field_from_seeds(w1, h1, seeds_in_order, base=None):
    """field after inserting the seeds (cells) in order into `base` (or an empty field), by the closed forms only"""
    val = np.full((h1, w1), 1000, np.int64) if base is None else base.astype(np.int64).copy()
    ys, xs = np.mgrid[0:h1, 0:w1]
    interior = ~((xs == 0) | (ys == 0) | (xs == w1 - 1) | (ys == h1 - 1))
    for (sx, sy) in seeds_in_order:
        val[sy, sx] = 0
        if is_border(sx, sy, w1, h1):
            continue
        dx, dy = np.abs(xs - sx), np.abs(ys - sy)
        k = np.maximum(dx, dy)
        for _ in range(64):
            k = np.where(k + (k + 1) // 2 < dx + dy, k + 1, k)
        k = np.where(k > 39, 1000, k)
        val = np.where(interior, np.minimum(val, k), val)
        # border cells: offers of the interior neighbours, a diagonal offer needs an odd step number
        for by in range(h1):
            for bx in ([0, w1 - 1] if 0 < by < h1 - 1 else range(w1)):
                for j, (ddx, ddy) in enumerate(NB8):
                    qx, qy = bx + ddx, by + ddy
                    if not (0 <= qx < w1 and 0 <= qy < h1) or is_border(qx, qy, w1, h1):
                        continue
                    kk = val[qy, qx] + 1
                    if kk <= 39 and (j < 4 or kk % 2 == 1) and kk < val[by, bx]:
                        val[by, bx] = kk
    return val.astype(np.float32)

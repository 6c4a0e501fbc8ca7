"""ctypes view of the oracle's immature-point entry points (oracle/trace.cpp) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import numpy as np
import oracle_py as O

lib = O.lib
_fp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
lib.orc_immature_init_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, _fp, C.c_void_p, _ip]
lib.orc_trace_on.argtypes = [C.c_void_p, C.c_int, _fp, _fp, _fp, C.c_int, C.c_void_p, _ip]
lib.orc_trace_stereo.argtypes = [C.c_void_p, C.c_int, _fp, C.c_int, C.c_int, C.c_void_p, _ip]

# same layout as sdso_immature_point (include/sdso_b200.h) and orc::ImmaturePoint
DTYPE = np.dtype([
    ("u", "f4"), ("v", "f4"), ("idepth_min", "f4"), ("idepth_max", "f4"), ("quality", "f4"), ("energyTH", "f4"),
    ("color", "f4", 8), ("weights", "f4", 8), ("gradH", "f4", 4),
    ("u_stereo", "f4"), ("v_stereo", "f4"), ("idepth_min_stereo", "f4"), ("idepth_max_stereo", "f4"), ("idepth_stereo", "f4"),
    ("lastTraceUV", "f4", 2), ("lastTracePixelInterval", "f4"),
    ("lastTraceStatus", "i4"), ("bestIdx", "i4"), ("numSteps", "i4")])
assert lib.orc_immature_record_size() == DTYPE.itemsize


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def immature_init(orc, fid, uv):
    uv = _f32(uv).reshape(-1, 2)
    pts = np.zeros(uv.shape[0], DTYPE)
    ok = np.zeros(uv.shape[0], np.int32)
    lib.orc_immature_init_batch(orc._h, fid, uv.shape[0], uv.ctypes.data_as(_fp), pts.ctypes.data, ok.ctypes.data_as(_ip))
    return pts, ok.astype(bool)


def trace_on(orc, fid, KRKi, Kt, aff, pts):
    K_, t_, a_ = _f32(KRKi).reshape(9), _f32(Kt).reshape(3), _f32(aff).reshape(2)
    st = np.zeros(pts.size, np.int32)
    lib.orc_trace_on(orc._h, fid, K_.ctypes.data_as(_fp), t_.ctypes.data_as(_fp), a_.ctypes.data_as(_fp), pts.size, pts.ctypes.data, st.ctypes.data_as(_ip))
    return st


def trace_stereo(orc, fid, K, mode_right, pts):
    K_ = _f32(K).reshape(9)
    st = np.zeros(pts.size, np.int32)
    lib.orc_trace_stereo(orc._h, fid, K_.ctypes.data_as(_fp), int(mode_right), pts.size, pts.ctypes.data, st.ctypes.data_as(_ip))
    return st

lib.orc_activate_points.argtypes = [C.c_void_p, C.c_int, _ip, C.c_void_p, C.c_int, C.c_int, _ip, _fp, _ip, _fp]


def activate_points(orc, nframes, host, pts, variant=0, min_obs=1):
    """D4 on the window currently held by the oracle context (oracle_ba_py.OracleBA of the same Oracle)."""
    n = pts.size
    host = np.ascontiguousarray(host, np.int32)
    res, st = np.zeros(n, np.int32), np.zeros((n, nframes), np.int32)
    idp, en = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib.orc_activate_points(orc._h, n, host.ctypes.data_as(_ip), pts.ctypes.data, variant, min_obs, res.ctypes.data_as(_ip), idp.ctypes.data_as(_fp),
                            st.ctypes.data_as(_ip), en.ctypes.data_as(_fp))
    return dict(result=res, idepth=idp, states=st, energy=en)

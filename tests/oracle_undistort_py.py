"""ctypes view of the oracle's input preparation / trajectory rows (oracle/oracle_undistort.hpp) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import numpy as np
import oracle_py as O

lib = O.lib
_fp, _dp, _ubp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_ubyte)
lib.orc_undistort.restype = C.c_float
lib.orc_undistort.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _ubp, C.c_float, C.c_float, _fp]
lib.orc_trajectory_row.argtypes = [_dp, C.c_char_p, C.c_int]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def undistort(raw, w, h, remap_x, remap_y, G=None, vignette_inv=None, photometric_calibration=2, use_exposure=True, exposure=1.0, factor=1.0):
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    h_org, w_org = raw.shape
    rx, ry = _f32(remap_x), _f32(remap_y)
    g = _f32(G) if G is not None else None
    v = _f32(vignette_inv) if vignette_inv is not None else None
    out = np.zeros((h, w), np.float32)
    e = lib.orc_undistort(w_org, h_org, w, h, rx.ctypes.data_as(_fp), ry.ctypes.data_as(_fp), g.ctypes.data_as(_fp) if g is not None else None,
                          v.ctypes.data_as(_fp) if v is not None else None, int(photometric_calibration), int(use_exposure), raw.ctypes.data_as(_ubp),
                          float(exposure), float(factor), out.ctypes.data_as(_fp))
    return out, e


def trajectory_row(T):
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(12)
    buf = C.create_string_buffer(512)
    n = lib.orc_trajectory_row(T.ctypes.data_as(_dp), buf, 512)
    assert n > 0
    return buf.value.decode()


def radial_remap(w, h, w_org, h_org, k1=-0.18, k2=0.03, crop=0.92):
    """remap tables of a radial-tangential camera rectified to a pinhole crop (the kind of table Undistort builds); entries
    that leave the raw image are marked -1 as Undistort.cpp:690-712 does"""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    fx = fy = 0.6 * w_org
    cx, cy = w_org / 2 - 0.5, h_org / 2 - 0.5
    xn = (xs - (w / 2 - 0.5)) / (fx * crop)
    yn = (ys - (h / 2 - 0.5)) / (fy * crop)
    r2 = xn * xn + yn * yn
    f = 1 + k1 * r2 + k2 * r2 * r2
    rx = (fx * xn * f + cx).astype(np.float32)
    ry = (fy * yn * f + cy).astype(np.float32)
    bad = ~((rx > 0.01) & (ry > 0.01) & (rx < w_org - 1.01) & (ry < h_org - 1.01))
    bad[:2, :] = True; bad[:, -3:] = True   # some entries always fall outside (black border of a rectified image)
    rx[bad] = -1
    ry[bad] = -1
    return rx, ry

"""ctypes view of the oracle's pixel selector (oracle/oracle_select.hpp) — TEST INFRASTRUCTURE ONLY — and a numpy
restatement of the *decomposed* selection (cell masks -> running count -> per-cell / per-block arg-max) that the device
path uses, so that the equivalence with the reference's sequential walk is checked on the CPU too."""
import ctypes as C
import numpy as np
import oracle_py as O

lib = O.lib
_fp, _ip, _ubp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
lib.orc_sel_create.restype = C.c_void_p
lib.orc_sel_create.argtypes = [C.c_int, C.c_int]
lib.orc_sel_destroy.argtypes = [C.c_void_p]
lib.orc_sel_random_pattern.argtypes = [C.c_void_p, _ubp]
lib.orc_sel_potential.argtypes = [C.c_void_p, C.c_int]
lib.orc_sel_forget_hist.argtypes = [C.c_void_p]
lib.orc_sel_ths.argtypes = [C.c_void_p, _fp, _fp]
lib.orc_sel_make_hists.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lib.orc_sel_select.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _fp, C.c_int, C.c_float, _ip]
lib.orc_sel_make_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _fp, C.c_float, C.c_int, C.c_float]

DIRECTIONS = np.array([[0, 1.0], [0.3827, 0.9239], [0.1951, 0.9808], [0.9239, 0.3827], [0.7071, 0.7071], [0.3827, -0.9239],
                       [0.8315, 0.5556], [0.8315, -0.5556], [0.5556, -0.8315], [0.9808, 0.1951], [0.9239, -0.3827], [0.7071, -0.7071],
                       [0.5556, 0.8315], [0.9808, -0.1951], [1.0, 0.0], [0.1951, -0.9808]], np.float32)


class Selector:
    def __init__(self, orc):
        self.orc = orc
        self.w, self.h = orc.w, orc.h
        self._h = C.c_void_p(lib.orc_sel_create(orc.w, orc.h))

    def close(self):
        if self._h:
            lib.orc_sel_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def random_pattern(self):
        out = np.zeros(self.w * self.h, np.uint8)
        lib.orc_sel_random_pattern(self._h, out.ctypes.data_as(_ubp))
        return out

    def potential(self, set=0):
        return lib.orc_sel_potential(self._h, int(set))

    def forget_hist(self):
        lib.orc_sel_forget_hist(self._h)

    def make_hists(self, fid):
        lib.orc_sel_make_hists(self._h, self.orc._h, fid)
        ths = np.zeros((self.h // 32, self.w // 32), np.float32)
        sm = np.zeros_like(ths)
        lib.orc_sel_ths(self._h, ths.ctypes.data_as(_fp), sm.ctypes.data_as(_fp))
        return ths, sm

    def select(self, fid, pot, th_factor=1.0):
        m = np.zeros((self.h, self.w), np.float32)
        n = np.zeros(3, np.int32)
        lib.orc_sel_select(self._h, self.orc._h, fid, m.ctypes.data_as(_fp), int(pot), float(th_factor), n.ctypes.data_as(_ip))
        return m, n

    def make_maps(self, fid, density=3000.0, recursions_left=1, th_factor=1.0):
        m = np.zeros((self.h, self.w), np.float32)
        n = lib.orc_sel_make_maps(self._h, self.orc._h, fid, m.ctypes.data_as(_fp), float(density), int(recursions_left), float(th_factor))
        return m, n


def visiting_order(w, h, pot):
    """cells (x, y origin) in the order of the reference's nested loops, with their 2pot / 4pot block ids"""
    cells = []
    b4 = b3 = -1
    for y4 in range(0, h, 4 * pot):
        for x4 in range(0, w, 4 * pot):
            b4 += 1
            for y3 in range(0, min(4 * pot, h - y4), 2 * pot):
                for x3 in range(0, min(4 * pot, w - x4), 2 * pot):
                    b3 += 1
                    for y2 in range(0, min(2 * pot, h - y3 - y4), pot):
                        for x2 in range(0, min(2 * pot, w - x3 - x4), pot):
                            cells.append((x2 + x3 + x4, y2 + y3 + y4, b3, b4))
    return cells


def select_decomposed(dI0, ag0, ag1, ag2, ths_smoothed, rp, pot, th_factor=1.0, dw1=0.75):
    """the device formulation of PixelSelector::select, in numpy (float32 throughout)"""
    f = np.float32
    h, w = ag0.shape
    w1, w2 = ag1.shape[1], ag2.shape[1]
    ys, xs = np.mgrid[0:h, 0:w]
    inb = ~((xs < 4) | (xs >= w - 5) | (ys < 4) | (ys > h - 4))
    th0 = ths_smoothed[ys >> 5, xs >> 5].astype(f)
    thf = f(th_factor)
    dw1 = f(dw1)
    dw2 = f(dw1 * dw1)
    th1 = th0 * dw1
    th2 = th1 * dw2
    p0 = inb & (ag0 > th0 * thf)
    p1 = inb & (ag1[ys // 2, xs // 2] > th1 * thf)
    p2 = inb & (ag2[ys // 4, xs // 4] > th2 * thf)
    dx, dy = dI0[..., 1].astype(f), dI0[..., 2].astype(f)
    dn = np.abs(dx[None] * DIRECTIONS[:, 0, None, None] + dy[None] * DIRECTIONS[:, 1, None, None])  # (16, h, w), float32 products and sum
    cells = visiting_order(w, h, pot)
    # 1. masks
    masks = []
    for (x0, y0, _, _) in cells:
        sl = (slice(y0, min(y0 + pot, h)), slice(x0, min(x0 + pot, w)))
        q = p0[sl][None] & (dn[(slice(None),) + sl] > 0)
        masks.append(q.reshape(16, -1).any(axis=1))
    # 2. running count: certain cells by a prefix sum, ambiguous ones serially
    n2 = 0
    n2_at = []
    sel = []
    for m in masks:
        n2_at.append(n2)
        s = bool(m[rp[n2] & 0xF])
        sel.append(s)
        n2 += s
    out = np.zeros((h, w), np.float32)

    def argmax_first(cell_list, cond, d):
        best, bv = None, f(0)
        for (x0, y0) in cell_list:
            for y in range(y0, min(y0 + pot, h)):
                for x in range(x0, min(x0 + pot, w)):
                    if cond[y, x] and dn[d, y, x] > bv:
                        bv, best = dn[d, y, x], (y, x)
        return best

    # 3. level 0
    for k, (x0, y0, _, _) in enumerate(cells):
        if sel[k]:
            out[argmax_first([(x0, y0)], p0, rp[n2_at[k]] & 0xF)] = 1
    # 4. levels 1 and 2
    n3 = n4 = 0
    blocks3, blocks4 = {}, {}
    for k, (x0, y0, b3, b4) in enumerate(cells):
        blocks3.setdefault(b3, []).append(k)
        blocks4.setdefault(b4, []).append(k)
    fired4 = {}
    for b3, ks in blocks3.items():
        b4 = cells[ks[0]][3]
        if any(sel[k] for k in ks):
            fired4[b4] = True
            continue
        best = argmax_first([cells[k][:2] for k in ks], p1, rp[n2_at[ks[0]]] & 0xF)
        if best is not None:
            out[best] = 2
            n3 += 1
            fired4[b4] = True
    for b4, ks in blocks4.items():
        if fired4.get(b4):
            continue
        best = argmax_first([cells[k][:2] for k in ks], p2, rp[n2_at[ks[0]]] & 0xF)
        if best is not None:
            out[best] = 4
            n4 += 1
    namb = sum(1 for m in masks if m.any() and not m.all())
    return out, np.array([n2, n3, n4], np.int32), namb

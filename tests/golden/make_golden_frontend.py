"""Generates tests/golden/frontend_v1.npz — committed fixtures of the rows either side of the residual path: pixel selection,
coarse distance map + activation candidate filter, undistortion.

PROVENANCE: as for hotpath_v1.npz — the reference holds no golden vectors for these rows and cannot be compiled here, so the
vectors come from the ORACLE (oracle/), not from the reference ("parity unpinned"). They pin the restatement
and the device path against drift. Run from the repo root: python tests/golden/make_golden_frontend.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_py as O              # noqa: E402
import oracle_select_py as S       # noqa: E402
import oracle_distmap_py as D      # noqa: E402
import oracle_undistort_py as U    # noqa: E402

W, H = 384, 256
K4 = (300.0, 300.0, 191.5, 127.5)


def image(seed=5):
    rng = np.random.default_rng(seed)
    big = rng.normal(0, 1, (H, W))
    k = np.exp(-0.5 * (np.arange(-5, 6) / 1.6) ** 2); k /= k.sum()
    for ax in (0, 1):
        big = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, big)
    img = 128 + 55 * big / big.std()
    img[:, : W // 3] = 128 + 0.15 * (img[:, : W // 3] - 128)   # a low-texture third: levels 1 and 2 of the selector fire there
    return np.round(np.clip(img, 0, 255)).astype(np.uint8)


def vignette(w_org, h_org):
    """inverse vignette 1 + 0.5 r^2 (plain IEEE arithmetic: reproducible without a random stream)"""
    ys, xs = np.mgrid[0:h_org, 0:w_org].astype(np.float64)
    r2 = ((xs - w_org / 2) / w_org) ** 2 + ((ys - h_org / 2) / h_org) ** 2
    return (1.0 + 0.5 * r2).astype(np.float32)


def build():
    img = image()
    orc = O.Oracle(W, H, K4, 0.5)
    assert orc.levels >= 3
    f = orc.frame_new()
    orc.make_images(f, img.astype(np.float32))
    out = dict(image=img, K4=np.array(K4))
    sel = S.Selector(orc)
    ths, sm = sel.make_hists(f)
    out["ths"] = ths; out["ths_smoothed"] = sm
    for pot in (1, 3, 5):
        m, n = sel.select(f, pot, 1.0)
        out[f"select_pot{pot}"] = m.astype(np.uint8); out[f"select_n_pot{pot}"] = n
    sel.potential(3)
    dens, nums, pots = [1200.0, 1200.0, 300.0, 5000.0], [], []
    for i, d in enumerate(dens):
        m, num = sel.make_maps(f, d)
        out[f"maps_{i}"] = m.astype(np.uint8)
        nums.append(num); pots.append(sel.potential())
    out["maps_density"] = np.array(dens); out["maps_num"] = np.array(nums); out["maps_potential"] = np.array(pots)
    # distance map + candidate filter
    inp = D.make_inputs(orc, 21, n_hosts=5, n_pts=400, n_cand=2500)
    dm = D.DistMap(orc)
    out["dm_map"] = dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"]).astype(np.int16)
    for k in ("KRKi", "Kt", "pt_host", "pt_uvid", "cand_host", "my_type", "flagged"):
        out["dm_" + k] = inp[k]
    out["dm_pts"] = inp["pts"].copy().view(np.uint8)
    for mad in (0.5, 2.0):
        dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
        v, m = dm.filter(inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
        out[f"dm_verdict_{mad}"] = v.astype(np.int8); out[f"dm_after_{mad}"] = m.astype(np.int16)
    # undistortion
    rng = np.random.default_rng(9)
    w_org, h_org = 400, 270
    raw = rng.integers(0, 256, (h_org, w_org), dtype=np.uint8)
    rx, ry = U.radial_remap(W, H, w_org, h_org)
    G = (np.linspace(0, 255, 256) ** 1.1 / 255 ** 0.1).astype(np.float32)
    vig = vignette(w_org, h_org)
    und, e = U.undistort(raw, W, H, rx, ry, G, vig, photometric_calibration=2, exposure=0.02)
    out.update(und_raw=raw, und_rx=rx[::4, ::4].copy(), und_ry=ry[::4, ::4].copy(), und_G=G, und_out_sub=und[::3, ::5].copy(),
               und_out_sum=np.array(und.astype(np.float64).sum()), und_exposure=np.array(e))
    return out


if __name__ == "__main__":
    d = build()
    p = os.path.join(HERE, "frontend_v1.npz")
    np.savez_compressed(p, **d)
    print("wrote", p, os.path.getsize(p), "bytes")

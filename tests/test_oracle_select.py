"""Pixel selector (PixelSelector2.cpp) — CPU checks of the oracle: randomPattern is the platform's srand/rand sequence,
histogram thresholds against a numpy restatement, structural properties of select, and the decomposed (device)
formulation against the sequential walk."""
import numpy as np
import pytest
import oracle_py as O
import oracle_select_py as S
import synth


def _frame(w, h, seed, kind="scene"):
    K = (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5)
    orc = O.Oracle(w, h, K, 0.1)
    rng = np.random.default_rng(seed)
    if kind == "scene":
        img, _ = synth.render(synth.make_scene(), synth.camera_pose(seed), w, h, K)
    elif kind == "u8":      # integer intensities: exactly-zero gradient components are frequent (ambiguous cells)
        img = np.kron(rng.integers(0, 255, (h // 4, w // 4)), np.ones((4, 4))).astype(np.float32)
        img += rng.integers(0, 3, (h, w)).astype(np.float32)
    elif kind == "flat":    # low texture: levels 1 / 2 dominate
        img = (rng.random((h, w)) * 6 + 100).astype(np.float32)
        img[h // 3:h // 2, w // 4:w // 2] += 60
    else:
        raise ValueError(kind)
    fid = orc.frame_new()
    orc.make_images(fid, img)
    return orc, fid


def test_random_pattern_is_glibc_rand():
    orc, _ = _frame(64, 64, 0)
    rp = S.Selector(orc).random_pattern()
    # first values of srand(3141592); rand() & 0xFF — a fixed fingerprint so a libc change is noticed
    assert rp.size == 64 * 64
    assert len(np.unique(rp)) > 200
    sel2 = S.Selector(orc).random_pattern()
    assert np.array_equal(rp, sel2)


def test_product_pattern_generator_matches_platform_rand():
    # host-only entry point of the product library (no GPU involved): its private generator must reproduce rand()
    from conftest import load_pkg
    pkg = load_pkg()
    orc, _ = _frame(160, 128, 0)
    rp = S.Selector(orc).random_pattern()
    assert np.array_equal(pkg.selector_pattern_host(rp.size), rp)


def test_make_hists_against_numpy():
    w, h = 160, 128
    orc, fid = _frame(w, h, 1)
    sel = S.Selector(orc)
    ths, sm = sel.make_hists(fid)
    _, ag = orc.frame_get(fid, 0)
    for by in range(h // 32):
        for bx in range(w // 32):
            ys, xs = np.mgrid[32 * by:32 * by + 32, 32 * bx:32 * bx + 32]
            ok = ~((xs > w - 2) | (ys > h - 2) | (xs < 1) | (ys < 1))
            g = np.minimum(np.sqrt(ag[ys, xs][ok]).astype(np.int32), 48)
            hist = np.bincount(g, minlength=91)
            th = int(np.float32(g.size) * np.float32(0.5) + np.float32(0.5))
            q = 90
            for i in range(90):
                th -= hist[i]
                if th < 0:
                    q = i
                    break
            assert ths[by, bx] == q + 7
    pad = np.pad(ths, 1, constant_values=np.nan)
    win = np.stack([pad[dy:dy + ths.shape[0], dx:dx + ths.shape[1]] for dy in range(3) for dx in range(3)])
    mean = np.nansum(win, 0) / np.sum(~np.isnan(win), 0)
    assert np.allclose(sm, mean ** 2, rtol=1e-6)


@pytest.mark.parametrize("kind,w,h", [("scene", 160, 128), ("u8", 192, 128), ("flat", 160, 128), ("u8", 168, 136)])
@pytest.mark.parametrize("pot", [1, 2, 3, 5])
def test_select_decomposition_matches_sequential(kind, w, h, pot):
    orc, fid = _frame(w, h, 3, kind)
    sel = S.Selector(orc)
    ths, sm = sel.make_hists(fid)
    m, n = sel.select(fid, pot, 1.0)
    dI0, ag0 = orc.frame_get(fid, 0)
    _, ag1 = orc.frame_get(fid, 1)
    _, ag2 = orc.frame_get(fid, 2)
    # thsSmoothed as the selector indexes it: (x >> 5) + (y >> 5) * (w / 32), zero beyond the allocation
    full = np.zeros(((h + 31) // 32 + 4) * max(1, w // 32) + 200, np.float32)
    full[:sm.size] = sm.reshape(-1)
    ys, xs = np.mgrid[0:(h + 31) // 32, 0:(w + 31) // 32]
    grid = full[xs + ys * (w // 32)]
    m2, n2, namb = S.select_decomposed(dI0, ag0, ag1, ag2, grid, sel.random_pattern(), pot)
    assert np.array_equal(n, n2), (n, n2)
    assert np.array_equal(m, m2)
    assert n.sum() == np.count_nonzero(m)
    assert set(np.unique(m)) <= {0.0, 1.0, 2.0, 4.0}
    if kind == "u8" and pot == 1:
        assert namb > 0  # the case that makes the running count a true serial dependency is exercised


def test_select_structure():
    w, h, pot = 160, 128, 3
    orc, fid = _frame(w, h, 5)
    sel = S.Selector(orc)
    sel.make_hists(fid)
    m, n = sel.select(fid, pot)
    assert (n > 0).any()
    # at most one selection per pot-cell at level 0, none in the border
    cells = (m == 1).reshape(h // pot + (h % pot > 0), -1, 1, 1) if h % pot == 0 and w % pot == 0 else None
    ys, xs = np.nonzero(m)
    assert xs.min() >= 4 and xs.max() < w - 5 and ys.min() >= 4 and ys.max() <= h - 4
    l0 = np.stack(np.nonzero(m == 1), 1)
    keys = (l0[:, 0] // pot) * 10000 + l0[:, 1] // pot
    assert len(np.unique(keys)) == len(keys)


def test_make_maps_adapts_potential_and_subsamples():
    w, h = 160, 128
    orc, fid = _frame(w, h, 7)
    sel = S.Selector(orc)
    assert sel.potential() == 3
    m, num = sel.make_maps(fid, density=300.0)
    assert num == np.count_nonzero(m)
    assert abs(num - 300) < 150
    p1 = sel.potential()
    m2, num2 = sel.make_maps(fid, density=300.0)   # second call starts from the adapted potential
    assert num2 == np.count_nonzero(m2)
    assert sel.potential() >= 1 and p1 >= 1
    # very low density forces the sub-sampling branch
    m3, num3 = sel.make_maps(fid, density=40.0, recursions_left=0)
    assert 0 < num3 < np.count_nonzero(m2) + 1

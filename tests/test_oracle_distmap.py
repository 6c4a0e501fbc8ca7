"""Coarse distance map / activation candidate filter — CPU checks of the oracle and of the closed forms the device path
relies on (octagonal first-reach distance, order independence of the interior field, border replay)."""
import numpy as np
import pytest
import oracle_py as O
import oracle_distmap_py as D


def _orc(w=128, h=96):
    return O.Oracle(w, h, (0.6 * w, 0.6 * w, w / 2 - 0.5, h / 2 - 0.5), 0.1)


def test_single_seed_matches_closed_form():
    orc = _orc(256, 192)
    dm = D.DistMap(orc)
    w1, h1 = dm.w1, dm.h1
    for (sx, sy) in [(w1 // 2, h1 // 2), (1, 1), (w1 - 2, 5), (w1 - 1, 7), (30, h1 - 1)]:
        dm.make(np.zeros((1, 9)), np.zeros((1, 3)), np.zeros(0, np.int32), np.zeros((0, 3)))
        m = dm.add([[sx, sy]])
        assert np.array_equal(m, D.field_from_seeds(w1, h1, [(sx, sy)])), (sx, sy)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sequential_insertion_equals_closed_form_replay(seed):
    """addIntoDistFinal one by one (BFS) == min over seeds of the closed form, border cells replayed in order"""
    orc = _orc()
    dm = D.DistMap(orc)
    w1, h1 = dm.w1, dm.h1
    rng = np.random.default_rng(seed)
    seeds = [(int(rng.integers(1, w1)), int(rng.integers(1, h1))) for _ in range(25)]
    seeds += [(w1 - 1, int(rng.integers(1, h1))) for _ in range(3)] + [(int(rng.integers(1, w1)), h1 - 1) for _ in range(3)]
    rng.shuffle(seeds)
    dm.make(np.zeros((1, 9)), np.zeros((1, 3)), np.zeros(0, np.int32), np.zeros((0, 3)))
    for s in seeds:
        m = dm.add([list(s)])
    assert np.array_equal(m, D.field_from_seeds(w1, h1, seeds))
    # interior cells do not depend on the insertion order
    m2 = D.field_from_seeds(w1, h1, seeds[::-1])
    assert np.array_equal(m[1:-1, 1:-1], m2[1:-1, 1:-1])


def test_make_distance_map_projects_and_floods():
    orc = _orc(256, 192)
    dm = D.DistMap(orc)
    inp = D.make_inputs(orc, 3)
    m = dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    assert m.min() == 0 and set(np.unique(m)) <= set(range(40)) | {1000}
    zs = np.argwhere(m == 0)
    assert len(zs) > 100
    # the flood of the seed cells alone reproduces the map (interior)
    m2 = D.field_from_seeds(dm.w1, dm.h1, [(int(x), int(y)) for y, x in zs])
    assert np.array_equal(m[1:-1, 1:-1], m2[1:-1, 1:-1])


@pytest.mark.parametrize("mad", [0.0, 0.7, 2.0, 4.0])
def test_filter_verdicts(mad):
    orc = _orc()
    dm = D.DistMap(orc)
    inp = D.make_inputs(orc, 5, n_pts=200, n_cand=1500)
    dm.make(inp["KRKi"], inp["Kt"], inp["pt_host"], inp["pt_uvid"])
    verdict, m = dm.filter(inp["KRKi"], inp["Kt"], inp["flagged"], inp["cand_host"], inp["pts"], inp["my_type"], mad)
    assert set(np.unique(verdict)) <= {0, 1, 2}
    p = inp["pts"]
    dead = ~np.isfinite(p["idepth_max"]) | (p["lastTraceStatus"] == 2)
    assert (verdict[dead] == 2).all()
    assert (verdict == 1).sum() > 0
    if mad == 0.0:
        # nothing is held back by the distance field
        can = ~dead & np.isin(p["lastTraceStatus"], [0, 1, 3, 4]) & (p["lastTracePixelInterval"] < 8) & (p["quality"] > 3) & \
            ((p["idepth_max"] + p["idepth_min"]) > 0)
        assert ((verdict[can] == 1) | (verdict[can] == 2)).all()
    # accepted candidates are seeds of the final field
    assert (m == 0).sum() >= 1

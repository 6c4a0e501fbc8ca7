"""Committed golden fixtures of the BA side of the path (tests/golden/ba_v1.npz, made by tests/golden/make_golden_ba.py from the
oracle — see its docstring for provenance): the oracle must keep reproducing them. Inputs (8-bit images, window description,
vertex updates) are stored in the file; tolerances are rounding-level (the fixture was made by this very code: anything larger
is a change of the restated arithmetic and has to be deliberate)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_ba as MG   # noqa: E402

G = np.load(os.path.join(HERE, "golden", "ba_v1.npz"))


def test_fixture_is_a_meaningful_window():
    ns = G["out_newState"]
    assert (ns == 0).mean() > 0.55, "most residuals of the stored window are IN"
    assert int(G["out_lba_iterations"]) == 3 and int(G["out_lba_trials"]) >= 3
    assert np.all(np.isfinite(G["out_x"])) and np.abs(G["out_x"]).max() > 0
    assert np.allclose(G["out_H_top"], G["out_H_top"].T, rtol=1e-12, atol=1e-9)


def test_oracle_reproduces_ba_golden():
    out = MG.evaluate(G)
    for k, v in out.items():
        g = G["out_" + k]
        v = np.asarray(v)
        assert v.shape == g.shape, k
        if g.dtype.kind in "iub":
            assert np.array_equal(v, g), k
        else:
            scale = max(float(np.abs(g).max()), 1e-30)
            assert np.abs(v.astype(np.float64) - g.astype(np.float64)).max() <= 1e-9 * scale, (k, float(np.abs(v - g).max()), scale)

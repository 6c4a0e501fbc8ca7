"""D4 — point activation (FullSystem::optimizeImmaturePoint): CPU sanity of the oracle and GPU parity for both variants."""
import numpy as np
import pytest
import oracle_py as O
import oracle_ba_py as OB
import oracle_trace_py as OT
import ba_synth
import synth
import trace_synth as TS

SHAPES = {"640x192-n4": (640, 192, (360.0, 360.0, 319.5, 95.5), 4, 80, 150),
          "1232x368-n7": (synth.W, synth.H, synth.K4, 7, 140, 220)}   # config-3 window (7 key frames at the KITTI working resolution)


def candidates(win, orc, fids, rng, per_host=150, noise=0.15):
    """Immature points on every host frame with an inverse-depth interval around the true value (as the epipolar search leaves it)."""
    pts_all, host_all, true_all = [], [], []
    for h, f in enumerate(win["frames"]):
        uv = TS.candidate_pixels(f["image"], per_host, rng, margin=16, min_grad=6.0)
        pts, ok = OT.immature_init(orc, fids[h], uv)
        tid = 1.0 / f["depth"][uv[:, 1].astype(int), uv[:, 0].astype(int)]
        mid = tid * (1 + rng.normal(0, noise / 3, tid.size))
        pts["idepth_min"] = (mid * (1 - noise)).astype(np.float32)
        pts["idepth_max"] = (mid * (1 + noise)).astype(np.float32)
        pts_all.append(pts[ok]); host_all.append(np.full(ok.sum(), h, np.int32)); true_all.append(tid[ok])
    return np.concatenate(pts_all), np.concatenate(host_all), np.concatenate(true_all)


@pytest.fixture(scope="module", params=list(SHAPES))
def window(scene, request):
    W_, H_, K_, n, P, per_host = SHAPES[request.param]
    win = ba_synth.make_window(scene, n=n, P=P, seed=8, spacing=0.5, w=W_, h=H_, K=K_)
    win["shape"] = (W_, H_, K_)
    orc = O.Oracle(W_, H_, K_, synth.BASELINE)
    ba, fids, cw = ba_synth.fill_oracle(win, orc, OB.OracleBA, OB.immature_init)
    pts, host, tid = candidates(win, orc, fids, np.random.default_rng(5), per_host=per_host)
    return win, orc, ba, cw, pts, host, tid


def test_oracle_sse_activation_refines_inverse_depth(window):
    win, orc, ba, cw, pts, host, tid = window
    o = OT.activate_points(orc, win["n"], host, pts, variant=0)
    act = o["result"] == 1
    assert act.mean() > 0.5
    start = 0.5 * (pts["idepth_min"] + pts["idepth_max"])
    assert np.median(np.abs(o["idepth"][act] - tid[act]) / tid[act]) < np.median(np.abs(start[act] - tid[act]) / tid[act])
    assert np.all(o["states"][np.arange(host.size), host] == -1)


def test_oracle_g2o_activation_keeps_the_initial_inverse_depth(window):
    """the live activation edge projects once, so the LM cannot move the vertex."""
    win, orc, ba, cw, pts, host, tid = window
    o = OT.activate_points(orc, win["n"], host, pts, variant=1)
    assert np.array_equal(o["idepth"], (pts["idepth_max"] + pts["idepth_min"]) * np.float32(0.5))
    assert (o["result"] == 1).mean() > 0.95


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_gpu_activation_matches_oracle(pkg, window, variant):
    win, orc, ba, cw, pts, host, tid = window
    W_, H_, K_ = win["shape"]
    ctx = pkg.Context(W_, H_, K_, synth.BASELINE)
    Wd, _ = ba_synth.fill_device(win, ctx, pkg.Window, cw)
    rng = np.random.default_rng(1)
    extra = pts[:20].copy()
    extra["idepth_min"][:5] *= 8; extra["idepth_max"][:5] *= 8  # projects out of the targets
    extra["color"][10:] += 150                                   # energy above the threshold
    extra["weights"][5:10] = 1e-3                                # Hdd below setting_minIdepthH_act: not well constrained
    p2 = np.concatenate([pts, extra]); h2 = np.concatenate([host, host[:20]])
    o = OT.activate_points(orc, win["n"], h2, p2, variant=variant)
    g = Wd.activate_points(h2, p2, variant=variant)
    assert np.array_equal(g["result"], o["result"])
    assert np.array_equal(g["states"], o["states"])
    assert set(np.unique(o["result"])) >= ({1, 0} if variant == 0 else {1})
    assert np.allclose(g["idepth"], o["idepth"], rtol=1e-4, atol=1e-7)
    assert np.allclose(g["energy"], o["energy"], rtol=1e-4, atol=1e-3)
    ctx.close()

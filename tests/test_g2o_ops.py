"""V1-V5 (dso_g2o_vertex.cpp:15-106) and E3 (EdgeTracePointUVDSO, dso_g2o_edge.cpp:571-619) at operator level over SoA batches:
CPU checks of the oracle restatement and GPU parity through sdso_vertex_oplus / sdso_edge_trace_uv_eval."""
import numpy as np
import pytest
import oracle_py as O
import synth


def vertex_inputs(rng, n=257):
    poses = np.stack([synth.perturb_T(synth.T_cw(synth.camera_pose(k * 0.7)), rng, 0.1, 0.05) for k in range(n)])
    upd6 = rng.normal(0, 1, (n, 6)) * np.array([0.3, 0.3, 0.3, 0.2, 0.2, 0.2])
    upd6[0] = 0                      # identity update
    upd6[1, 3:] = 1e-12              # below Sophus' small-angle switch
    upd6[2, 3:] = [2.5, -1.0, 0.7]   # large rotation
    return dict(
        pose=(poses.reshape(n, 12), upd6),
        photo=(rng.normal(0, 1, (n, 2)), rng.normal(0, 0.1, (n, 2))),
        idepth=(rng.uniform(0.01, 1, n), rng.normal(0, 0.01, n)),
        uv=(rng.uniform(10, 600, (n, 2)), np.concatenate([rng.normal(0, 0.6, n - 4), [np.nan, np.inf, -np.inf, 0.5]]), rng.normal(0, 1, (n, 2))),
        cam=(rng.uniform(300, 800, (n, 4)), rng.normal(0, 1, (n, 4))))


def test_oracle_vertex_updates_are_what_the_reference_writes():
    rng = np.random.default_rng(0)
    v = vertex_inputs(rng)
    est, upd = v["pose"]
    out = O.vertex_oplus(1, est, upd).reshape(-1, 3, 4)
    for i in (0, 1, 2, 17):   # exp(update) * T with the oracle's own SE3 (pinned to Sophus' test set in test_oracle_se3.py)
        E = np.vstack([O.se3_exp(upd[i]), [0, 0, 0, 1]]); T = np.vstack([est[i].reshape(3, 4), [0, 0, 0, 1]])
        assert np.allclose(out[i], (E @ T)[:3], rtol=0, atol=1e-12)
    assert np.allclose(out[0], est[0].reshape(3, 4), rtol=0, atol=1e-14)   # (matrix -> quaternion -> matrix round trip)
    est, upd = v["photo"]
    assert np.array_equal(O.vertex_oplus(2, est, upd), est + upd)
    est, upd, aux = v["uv"]
    out = O.vertex_oplus(4, est, upd, aux)
    c = np.clip(upd, -0.5, 0.5)
    c[np.isnan(upd)] = 0             # NaN -> 0; +-inf hit the clamp first, as written (:76-84)
    assert np.array_equal(out, est + c[:, None] * aux)
    est, upd = v["cam"]
    assert np.array_equal(O.vertex_oplus(5, est, upd), est + upd)


def edge_inputs(rng, w, h, n=4000):
    uv = np.stack([rng.uniform(-3, w + 3, n), rng.uniform(-3, h + 3, n)], 1)
    uv[:8] = [[2.0, 2.0], [1.999, 50], [w - 6.0, 50], [w - 5.999, 50], [50, h - 6.0], [50, h - 5.999], [50, 2.0], [50, 1.999]]  # the border, both sides
    rot = rng.integers(-2, 3, (n, 2)).astype(np.float32)
    meas = rng.uniform(0, 255, n)
    dxdy = rng.normal(0, 1, (n, 2))
    return uv, rot, meas, dxdy


def test_oracle_trace_edge_border_and_stale_members(frames):
    orc = O.Oracle(synth.W, synth.H, synth.K4)
    f = orc.frame_new(); orc.make_images(f, frames[0][0])
    rng = np.random.default_rng(1)
    uv, rot, meas, dxdy = edge_inputs(rng, synth.W, synth.H)
    e0, J0 = np.full(uv.shape[0], 7.0), np.full(uv.shape[0], -3.0)
    err, J, flag = O.edge_trace_uv(orc, f, uv, rot, meas, (1.0, 0.0), dxdy, e0, J0)
    assert list(flag[:8]) == [1, 0, 1, 0, 1, 0, 1, 0]                       # u-2<0 / u+3>w-3 / v-2<0 / v+3>h-3 exactly at the limits
    out = flag == 0
    assert out.any() and np.all(err[out] == 0) and np.all(J[out] == -3.0)   # outside: error 0, Jacobian untouched
    ok = flag == 1
    dI, _ = orc.frame_get(f, 0)
    i = np.nonzero(ok & (uv[:, 0] % 1 > 0.1))[0][0]
    x, y = np.float32(uv[i, 0] + rot[i, 0]), np.float32(uv[i, 1] + rot[i, 1])
    ix, iy = int(x), int(y); dx, dy = x - ix, y - iy
    tex = dI[iy:iy + 2, ix:ix + 2].astype(np.float64)
    wts = np.array([[(1 - dx) * (1 - dy), dx * (1 - dy)], [(1 - dx) * dy, dx * dy]])
    hit = (tex * wts[..., None]).sum((0, 1))
    assert np.isclose(err[i], hit[0] - meas[i], rtol=1e-5, atol=1e-4)
    assert np.isclose(J[i], dxdy[i, 0] * hit[1] + dxdy[i, 1] * hit[2], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_gpu_vertex_oplus_matches_oracle(pkg):
    ctx = pkg.Context(640, 192, (360.0, 360.0, 319.5, 95.5))
    rng = np.random.default_rng(0)
    v = vertex_inputs(rng)
    est, upd = v["pose"]
    g, o = ctx.vertex_oplus(pkg.VERTEX_SE3_POSE, est, upd), O.vertex_oplus(1, est, upd)
    assert np.abs(g - o).max() < 1e-12                        # f64 on both sides; Rodrigues form vs Sophus' quaternion form
    R = g.reshape(-1, 3, 4)[:, :, :3]
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-12
    for kind, key in ((pkg.VERTEX_PHOTOMETRIC, "photo"), (pkg.VERTEX_INVERSE_DEPTH, "idepth"), (pkg.VERTEX_CAM, "cam")):
        est, upd = v[key]
        assert np.array_equal(ctx.vertex_oplus(kind, est, upd), O.vertex_oplus(kind, est, upd)), key
    est, upd, aux = v["uv"]
    assert np.array_equal(ctx.vertex_oplus(pkg.VERTEX_UV, est, upd, aux), O.vertex_oplus(4, est, upd, aux))
    with pytest.raises(pkg.SdsoError):
        ctx.vertex_oplus(pkg.VERTEX_UV, est, upd, None)       # VertexUVDSO without SetDxDy
    assert ctx.vertex_oplus(pkg.VERTEX_CAM, np.zeros((0, 4)), np.zeros((0, 4))).shape == (0, 4)   # empty batch
    ctx.close()


@pytest.mark.gpu
def test_gpu_trace_edge_matches_oracle(pkg, frames):
    ctx = pkg.Context(synth.W, synth.H, synth.K4)
    orc = O.Oracle(synth.W, synth.H, synth.K4)
    fg, fo = ctx.frame_create(), orc.frame_new()
    ctx.make_images(fg, frames[0][0]); orc.make_images(fo, frames[0][0])
    rng = np.random.default_rng(1)
    uv, rot, meas, dxdy = edge_inputs(rng, synth.W, synth.H, 20000)
    e0, J0 = rng.normal(0, 1, uv.shape[0]), rng.normal(0, 1, uv.shape[0])
    for aff in ((1.0, 0.0), (1.07, -3.5)):
        eg, Jg, fl_g = ctx.edge_trace_uv_eval(fg, uv, rot, meas, aff, dxdy, e0, J0)
        eo, Jo, fl_o = O.edge_trace_uv(orc, fo, uv, rot, meas, aff, dxdy, e0, J0)
        assert np.array_equal(fl_g, fl_o) and set(np.unique(fl_o)) == {0, 1}
        assert np.array_equal(eg, eo) and np.array_equal(Jg, Jo)   # same float gather, same promotion to double: bit-exact
    ctx.close()

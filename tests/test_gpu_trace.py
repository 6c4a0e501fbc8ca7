"""D1-D3 parity on the GPU through the C ABI against the oracle: status enum, bestIdx, numSteps bit-exact; every float
field of the record within 1e-4 relative (north_star)."""
import numpy as np
import pytest
import oracle_py as O
import oracle_trace_py as OT
import synth
import trace_synth as TS

pytestmark = pytest.mark.gpu
REL = 1e-4
FLOAT_FIELDS = ["u", "v", "idepth_min", "idepth_max", "quality", "energyTH", "color", "weights", "gradH", "u_stereo", "v_stereo",
                "idepth_min_stereo", "idepth_max_stereo", "idepth_stereo", "lastTraceUV", "lastTracePixelInterval"]


def assert_records_match(g, o, where=""):
    assert np.array_equal(g["lastTraceStatus"], o["lastTraceStatus"]), where + " status"
    assert np.array_equal(g["bestIdx"], o["bestIdx"]), where + " bestIdx"
    assert np.array_equal(g["numSteps"], o["numSteps"]), where + " numSteps"
    for f in FLOAT_FIELDS:
        a, b = g[f].astype(np.float64), o[f].astype(np.float64)
        assert np.array_equal(np.isnan(a), np.isnan(b)), where + " NaN pattern of " + f
        assert np.array_equal(np.isinf(a), np.isinf(b)), where + " Inf pattern of " + f
        fin = np.isfinite(b)
        assert np.allclose(a[fin], b[fin], rtol=REL, atol=1e-7), where + " " + f


@pytest.fixture(scope="module")
def setup(pkg, frames):
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE)
    oid, gid = {}, {}
    for name, key in (("l", 0), ("r", "r0"), ("n", 1)):
        oid[name], gid[name] = orc.frame_new(), ctx.frame_create()
        orc.make_images(oid[name], frames[key][0]); ctx.make_images(gid[name], frames[key][0])
    rng = np.random.default_rng(2)
    uv = TS.candidate_pixels(frames[0][0], 2000, rng)
    yield orc, ctx, oid, gid, uv
    ctx.close()


def test_constructor(setup, pkg):
    orc, ctx, oid, gid, uv = setup
    assert pkg.IMMATURE_DTYPE == OT.DTYPE
    uv2 = np.concatenate([uv, uv[:50] + np.float32(0.5)])  # half-pixel positions exercise the BiLin weights
    po, oko = OT.immature_init(orc, oid["l"], uv2)
    pg, okg = ctx.immature_init(gid["l"], uv2)
    assert np.array_equal(okg, oko)
    assert_records_match(pg, po, "ctor")
    assert np.array_equal(pg["color"], po["color"]) and np.array_equal(pg["weights"], po["weights"])


def test_trace_stereo_left_to_right_and_back(setup):
    orc, ctx, oid, gid, uv = setup
    po, _ = OT.immature_init(orc, oid["l"], uv)
    pg = po.copy()
    so = OT.trace_stereo(orc, oid["r"], TS.K33(), True, po)
    sg = ctx.trace_stereo(gid["r"], TS.K33(), True, pg)
    assert np.array_equal(sg, so) and (so == 0).mean() > 0.5
    assert_records_match(pg, po, "stereo L->R")
    # back-trace of the matched positions in the left image (FullSystem.cpp:588-596): points of the RIGHT frame
    uvr = po["lastTraceUV"][so == 0][:500]
    qo, _ = OT.immature_init(orc, oid["r"], uvr)
    qg = qo.copy()
    s2o = OT.trace_stereo(orc, oid["l"], TS.K33(), False, qo)
    s2g = ctx.trace_stereo(gid["l"], TS.K33(), False, qg)
    assert np.array_equal(s2g, s2o)
    assert_records_match(qg, qo, "stereo R->L")


def test_trace_on_after_stereo(setup):
    orc, ctx, oid, gid, uv = setup
    po, _ = OT.immature_init(orc, oid["l"], uv)
    st = OT.trace_stereo(orc, oid["r"], TS.K33(), True, po)
    good = st == 0
    po["idepth_min"] = np.where(good, po["idepth_min_stereo"], po["idepth_min"])
    po["idepth_max"] = np.where(good, po["idepth_max_stereo"], po["idepth_max"])
    pg = po.copy()
    KRKi, Kt = TS.krki_kt(synth.camera_pose(0), synth.camera_pose(1))
    for rep, aff in enumerate(((1.0, 0.0), (1.02, -1.5))):
        so = OT.trace_on(orc, oid["n"], KRKi, Kt, aff, po)
        sg = ctx.trace_on(gid["n"], KRKi, Kt, aff, pg)
        assert np.array_equal(sg, so), rep
        assert_records_match(pg, po, f"traceOn pass {rep}")
    assert (so == 0).sum() > 100


def test_all_status_branches(setup):
    orc, ctx, oid, gid, uv = setup
    rng = np.random.default_rng(0)
    base, _ = OT.immature_init(orc, oid["l"], uv[:64])
    KRKi, Kt = TS.krki_kt(synth.camera_pose(0), synth.camera_pose(1))
    po = TS.adversarial(base.copy(), rng)
    po["idepth_min"][:64] = 0.02; po["idepth_max"][:64] = 0.2
    pg = po.copy()
    seen = set()
    for rep in range(2):
        so = OT.trace_on(orc, oid["n"], KRKi, Kt, (1.0, 0.0), po)
        sg = ctx.trace_on(gid["n"], KRKi, Kt, (1.0, 0.0), pg)
        assert np.array_equal(sg, so)
        assert_records_match(pg, po, f"adversarial traceOn {rep}")
        seen |= set(np.unique(so))
    assert {0, 1, 2, 3, 4} <= seen
    qo = TS.adversarial(base.copy(), rng)
    qg = qo.copy()
    so = OT.trace_stereo(orc, oid["r"], TS.K33(), True, qo)
    sg = ctx.trace_stereo(gid["r"], TS.K33(), True, qg)
    assert np.array_equal(sg, so)
    assert_records_match(qg, qo, "adversarial stereo")


def test_empty_batch(setup):
    orc, ctx, oid, gid, uv = setup
    pts, ok = ctx.immature_init(gid["l"], np.zeros((0, 2), np.float32))
    assert pts.size == 0
    assert ctx.trace_on(gid["n"], np.eye(3), np.zeros(3), (1, 0), pts).size == 0


def test_multi_host_launch_and_resident_pool_equal_per_host_calls(setup, frames, scene, pkg):
    """All hosts of a window in ONE launch (per-point transform index), with host records and with the device-resident pool:
    bit-identical to the per-host calls, which are bit-exact against the oracle above. Long segments (> 32 and > 64 steps)
    exercise the second search round of a warp."""
    orc, ctx, oid, gid, uv = setup
    rng = np.random.default_rng(7)
    hosts = [synth.camera_pose(0), synth.camera_pose(0.5), synth.camera_pose(2)]
    imgs = [frames[0][0], synth.render(scene, hosts[1])[0], frames[2][0]]
    hf = []
    for im in imgs:
        f = ctx.frame_create(); ctx.make_images(f, im); hf.append(f)
    target = synth.camera_pose(1)
    per_host, xf = [], []
    for h, pose in enumerate(hosts):
        pts, ok = ctx.immature_init(hf[h], TS.candidate_pixels(imgs[h], 700, rng))
        pts = pts[ok]
        pts["idepth_min"] = rng.uniform(0.005, 0.03, pts.size).astype(np.float32)
        pts["idepth_max"] = (pts["idepth_min"] * rng.uniform(1.5, 40, pts.size)).astype(np.float32)   # segments of a few to > 64 steps
        pts["idepth_max"][::17] = np.nan                                                              # unbounded prior
        per_host.append(pts)
        KRKi, Kt = TS.krki_kt(pose, target)
        xf.append((KRKi, Kt, (1.0 + 0.01 * h, -0.5 * h)))
    ref, ref_st = [], []
    for pts, (KRKi, Kt, aff) in zip(per_host, xf):
        q = pts.copy()
        ref_st.append(ctx.trace_on(gid["n"], KRKi, Kt, aff, q)); ref.append(q)
    ref, ref_st = np.concatenate(ref), np.concatenate(ref_st)
    # (maxPixSearch = 0.027 (w + h) caps a segment at 45 steps here: two search rounds per lane)
    assert ref["numSteps"].max() > 32 and (ref["numSteps"] > 32).sum() > 20 and len(set(np.unique(ref_st))) >= 3
    allp = np.concatenate(per_host)
    host_of = np.concatenate([np.full(p.size, h, np.int32) for h, p in enumerate(per_host)])
    perm = rng.permutation(allp.size)   # points of different hosts interleaved
    KR, KT, AF = np.stack([x[0] for x in xf]), np.stack([x[1] for x in xf]), np.array([x[2] for x in xf], np.float32)
    a = np.ascontiguousarray(allp[perm])
    st = ctx.trace_on_hosts(gid["n"], KR, KT, AF, host_of[perm], a)
    assert np.array_equal(st, ref_st[perm]) and a.tobytes() == np.ascontiguousarray(ref[perm]).tobytes()
    # resident pool: upload once, trace in place (twice: the second pass starts from the first one's intervals), read back once
    ctx.immature_upload(np.ascontiguousarray(allp[perm]))
    st1 = ctx.trace_on_hosts(gid["n"], KR, KT, AF, host_of[perm], None)
    assert np.array_equal(st1, ref_st[perm])
    assert ctx.immature_download(a.size).tobytes() == a.tobytes()
    ctx.trace_on_hosts(gid["n"], KR, KT, AF, host_of[perm], None, want_status=False)
    b = a.copy()
    ctx.trace_on_hosts(gid["n"], KR, KT, AF, host_of[perm], b)
    assert ctx.immature_download(a.size).tobytes() == b.tobytes()
    assert ctx.immature_download(10, first=5).tobytes() == b[5:15].tobytes()
    with pytest.raises(pkg.SdsoError):
        ctx.trace_on_hosts(gid["n"], KR, KT, AF, np.zeros(a.size + 1, np.int32), None)   # more points than the pool holds
    for f in hf:
        ctx.frame_release(f)

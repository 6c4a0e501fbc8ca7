"""Batched forms of the hot path: one launch for many frames / many independent sequences must give exactly what the
single-frame calls give (the kernels are deterministic, so equality is bitwise)."""
import numpy as np
import pytest
import synth

pytestmark = pytest.mark.gpu


def test_batched_make_images_u8_and_float_equal_single(pkg, frames):
    import torch
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE)
    imgs = [frames[0][0], frames[1][0], frames[2][0], frames["r0"][0]]
    single = []
    for im in imgs:
        f = ctx.frame_create(); ctx.make_images(f, im); single.append(f)
    # float sources resident on the device
    dev = [torch.from_numpy(im).cuda().contiguous() for im in imgs]
    fb = [ctx.frame_create() for _ in imgs]
    ctx.make_images_batch_device(fb, [d.data_ptr() for d in dev], u8=False)
    # uint8 sources uploaded asynchronously from pinned host memory (the synthetic images are quantised: exact in 8 bits)
    host8 = [torch.from_numpy(im.astype(np.uint8)).contiguous().pin_memory() for im in imgs]
    fu = [ctx.frame_create() for _ in imgs]
    ctx.upload_images_async(fu, [h.data_ptr() for h in host8], u8=True)
    ctx.make_images_uploaded(fu)
    ctx.synchronize()
    for k in range(len(imgs)):
        for lvl in range(ctx.levels):
            a, ag = ctx.frame_download(single[k], lvl)
            b, bg = ctx.frame_download(fb[k], lvl)
            c, cg = ctx.frame_download(fu[k], lvl)
            assert np.array_equal(a, b) and np.array_equal(ag, bg), (k, lvl, "float batch")
            assert np.array_equal(a, c) and np.array_equal(ag, cg), (k, lvl, "u8 batch")
    # the epipolar search reads the frame's planar level-0 image: must be valid after a u8 build
    uv = np.array([[100.0, 100.0], [640.5, 200.25]], np.float32)
    p1, _ = ctx.immature_init(single[0], uv); p2, _ = ctx.immature_init(fu[0], uv)
    assert np.array_equal(p1["color"], p2["color"])
    ctx.close()


def test_multi_reference_batch_equals_separate_tracking(pkg, scene, frames):
    """Three independent 'sequences' (own reference keyframe, own template, own new frame) tracked by ONE launch."""
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE)
    fid = {k: ctx.frame_create() for k in (0, 1, 2)}
    for k in (0, 1, 2):
        ctx.make_images(fid[k], frames[k][0])
    rng = np.random.default_rng(1)
    seqs = [(0, 1), (1, 2), (0, 2)]  # (reference, new)
    singles, T0s = [], []
    for s, (r, nw) in enumerate(seqs):
        ctx.tracker_select_ref(s)
        ctx.tracker_set_ref(fid[r], synth.pick_points(rng, frames[r][1], 1500 + 100 * s), (0.0, 0.0))
        Ttrue = synth.T_rel(synth.camera_pose(r), synth.camera_pose(nw))
        T0 = synth.perturb_T(Ttrue, rng, 0.03, np.deg2rad(0.3))
        T0s.append(T0)
        singles.append(ctx.track(fid[nw], T0, (0.0, 0.0), ctx.levels - 1, [np.nan] * 5, 0))
        assert np.abs(singles[-1]["T"][:, 3] - Ttrue[:, 3]).max() < 1e-2
    ctx.track_enqueue_multi([0, 1, 2], [fid[nw] for _, nw in seqs], np.stack(T0s), np.zeros((3, 2)), ctx.levels - 1, np.full((3, 5), np.nan), 0)
    b = ctx.track_collect(3)
    for s in range(3):
        assert np.array_equal(b["T"][s], singles[s]["T"]), s
        assert np.array_equal(b["lastResiduals"][s], singles[s]["lastResiduals"])
        assert b["ok"][s] == singles[s]["ok"]
    # slots are independent: re-reading slot 0's template after touching the others
    ctx.tracker_select_ref(0)
    n0 = ctx.tracker_get_pc(0)[0].size
    ctx.tracker_select_ref(2)
    assert ctx.tracker_get_pc(0)[0].size != n0
    ctx.close()


@pytest.mark.parametrize("cluster,threads,gather", [(1, 256, 1), (2, 256, 1), (4, 128, 1), (8, 256, 2), (2, 192, 2)])
def test_tracker_configurations_are_bitwise_identical(pkg, frames, cluster, threads, gather):
    """cluster size, CTA size, points in flight and the shared-memory texel / point cache change how the work is laid out, not the
    result: every configuration (cache on and off) returns the pose, residuals and iteration counts of the default one bit for bit
    (fixed-order reductions; a cache hit returns the texels a gather would)."""
    rng = np.random.default_rng(3)
    pts = synth.pick_points(rng, frames[0][1], 2000)
    Ttrue = synth.T_rel(synth.camera_pose(0), synth.camera_pose(1))
    T0 = synth.perturb_T(Ttrue, rng, 0.03, np.deg2rad(0.3))

    def run(settings):
        ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, settings=settings)
        f0, f1 = ctx.frame_create(), ctx.frame_create()
        ctx.make_images(f0, frames[0][0]); ctx.make_images(f1, frames[1][0])
        ctx.tracker_set_ref(f0, pts, (0.0, 0.0))
        r = ctx.track(f1, T0, (0.0, 0.0), ctx.levels - 1, [np.nan] * 5, 0)
        ctx.close()
        return r

    ref = run(None)
    assert np.abs(ref["T"][:, 3] - Ttrue[:, 3]).max() < 1e-2
    for cache in (1, 0):
        s = pkg.default_settings()
        s.cluster_size, s.block_threads, s.gather_batch, s.track_cache = cluster, threads, gather, cache
        r = run(s)
        # the accumulation order follows the thread <-> point mapping, so H differs in the last bits between layouts; the LM path
        # (accept / reject, iteration counts) and the converged pose must not
        assert r["ok"] == ref["ok"] and np.array_equal(r["iterations"], ref["iterations"]), (cluster, threads, gather, cache)
        assert np.abs(r["T"] - ref["T"]).max() < 1e-6
        if cache == 1:
            on = r
        else:
            assert np.array_equal(r["T"], on["T"]) and np.array_equal(r["lastResiduals"], on["lastResiduals"]), "cache on/off must be bitwise equal"


@pytest.mark.parametrize("shape", [(640, 480), (640, 192), (1920, 1088), (96, 64), (100, 60), (1241, 376), (136, 104)])
def test_make_images_u8_equals_float_at_every_depth(pkg, shape):
    """8-bit sources take wider loads where the region origin (64 bx - 2^(L-1)), the row pitch and the pointer allow it: 16-byte
    (L >= 5), 4-byte (L >= 3), scalar otherwise. Every combination must equal the float path bit for bit — including 4-level
    pyramids (640x480, 640x192), 1-3 level sizes, odd widths, and device pointers that are not 16-byte aligned."""
    import torch
    w, h = shape
    rng = np.random.default_rng(w * 13 + h)
    img8 = rng.integers(0, 256, (h, w), dtype=np.uint8)
    K = (500.0, 500.0, w / 2 - 0.5, h / 2 - 0.5)
    ctx = pkg.Context(w, h, K)
    ff = ctx.frame_create()
    ctx.make_images(ff, img8.astype(np.float32))
    slab = torch.zeros(w * h + 64, dtype=torch.uint8, device="cuda")
    for shift in (0, 4, 1):   # 16-byte aligned, 4-byte aligned, unaligned device pointer
        view = slab[shift:shift + w * h]
        view.copy_(torch.from_numpy(img8.reshape(-1)).cuda())
        fu = ctx.frame_create()
        ctx.make_images_batch_device([fu], [view.data_ptr()], u8=True)
        ctx.synchronize()
        for lvl in range(ctx.levels):
            a, ag = ctx.frame_download(ff, lvl)
            b, bg = ctx.frame_download(fu, lvl)
            assert np.array_equal(a, b) and np.array_equal(ag, bg), (shape, shift, lvl)
        ctx.frame_release(fu)
    ctx.close()


def test_multi_reference_g2o_batch_equals_separate_tracking(pkg, frames):
    """g2o variant with per-problem templates of DIFFERENT sizes in one launch: the per-problem edge flag / error scratch must be
    strided by the largest template of the batch (a later, smaller current slot used to size it)."""
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE)
    fid = {k: ctx.frame_create() for k in (0, 1, 2)}
    for k in (0, 1, 2):
        ctx.make_images(fid[k], frames[k][0])
    rng = np.random.default_rng(2)
    seqs = [(0, 1), (1, 2), (0, 2), (1, 0)]
    npts = [2400, 1500, 900, 300]   # the current slot after the loop is the smallest
    singles, T0s = [], []
    for s, (r, nw) in enumerate(seqs):
        ctx.tracker_select_ref(s)
        ctx.tracker_set_ref(fid[r], synth.pick_points(rng, frames[r][1], npts[s]), (0.0, 0.0))
        Ttrue = synth.T_rel(synth.camera_pose(r), synth.camera_pose(nw))
        T0 = synth.perturb_T(Ttrue, rng, 0.02, np.deg2rad(0.2))
        T0s.append(T0)
        singles.append(ctx.track(fid[nw], T0, (0.0, 0.0), ctx.levels - 1, [np.nan] * 5, pkg.VARIANT_G2O))
    for rep in range(2):
        ctx.track_enqueue_multi(list(range(4)), [fid[nw] for _, nw in seqs], np.stack(T0s), np.zeros((4, 2)), ctx.levels - 1,
                                np.full((4, 5), np.nan), pkg.VARIANT_G2O)
        b = ctx.track_collect(4)
        for s in range(4):
            assert np.array_equal(b["T"][s], singles[s]["T"]), (rep, s)
            assert np.array_equal(b["lastResiduals"][s], singles[s]["lastResiduals"], equal_nan=True), (rep, s)
            assert b["ok"][s] == singles[s]["ok"]
    # an empty current slot must not break a multi-reference enqueue (fill_params used to index frames[-1])
    ctx.tracker_select_ref(7)
    ctx.track_enqueue_multi([0, 1], [fid[1], fid[2]], np.stack(T0s[:2]), np.zeros((2, 2)), ctx.levels - 1, np.full((2, 5), np.nan), 0)
    ctx.track_collect(2)
    ctx.close()

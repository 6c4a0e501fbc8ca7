"""The drop-in boundary exercised by NATIVE callers (ctypes cannot see a signature that drifts between the header and the library):

* tests/native/abi_smoke.c    — plain C99 compiled against include/sdso_b200.h: ctx_create -> make_images -> tracker_set_ref -> track
* tests/native/adapter_test.cpp — C++ written against stereo-dso-g2o_b200/host/dso_adapters.hpp (the reference's class names and
  signatures: FrameHessian::makeImages, CoarseTracker::{makeK, setCoarseTrackingRef, trackNewestCoarse}, ImmaturePoint::traceStereo,
  VertexSE3PoseDSO::oplusImpl, EnergyFunctional)

Both are built with -Wall -Werror and linked with libsdso_b200.so. Without a GPU they must report "nodevice" (no CPU fallback);
on the GPU their results must equal what the same calls return through the ctypes view, bit for bit."""
import os
import subprocess
import numpy as np
import pytest
import synth
from conftest import ROOT, load_pkg, has_gpu

NATIVE = os.path.join(ROOT, "tests", "native")
LIBDIR = os.path.join(ROOT, "stereo-dso-g2o_b200")
W_, H_, K_ = 640, 192, (360.0, 360.0, 319.5, 95.5)


def build(tmp, name):
    out = os.path.join(tmp, name.split(".")[0])
    if name.endswith(".c"):
        cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-comment"]
    else:
        cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-comment"]
    cmd += ["-I" + os.path.join(ROOT, "include"), os.path.join(NATIVE, name), "-o", out, "-L" + LIBDIR, "-lsdso_b200", "-Wl,-rpath," + LIBDIR, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    load_pkg()   # the library must exist
    tmp = str(tmp_path_factory.mktemp("native"))
    return build(tmp, "abi_smoke.c"), build(tmp, "adapter_test.cpp")


@pytest.fixture(scope="module")
def case(tmp_path_factory, scene):
    rng = np.random.default_rng(4)
    p0, p1 = synth.camera_pose(0), synth.camera_pose(1)
    img0, d0 = synth.render(scene, p0, W_, H_, K_)
    img1, _ = synth.render(scene, p1, W_, H_, K_)
    imgR, _ = synth.render(scene, synth.right_of(p0), W_, H_, K_)
    pts = synth.pick_points(rng, d0, 700, W_, H_)
    hdi = rng.uniform(1e-4, 1e-2, len(pts)).astype(np.float32)
    # the weight exactly as CoarseTracker.cpp:350 forms it: sqrtf(1e-3 / (HdiF + 1e-12))
    pts[:, 3] = np.sqrt((1e-3 / (hdi.astype(np.float64) + 1e-12)).astype(np.float32))
    T0 = synth.perturb_T(synth.T_rel(p0, p1), rng, 0.02, np.deg2rad(0.2))
    d = str(tmp_path_factory.mktemp("case"))
    hdr = np.array([W_, H_, len(pts)], np.int32).tobytes() + np.array(K_, np.float32).tobytes() + np.float32(synth.BASELINE).tobytes()
    with open(os.path.join(d, "c.bin"), "wb") as f:   # abi_smoke: splats {u, v, idepth, weight}
        f.write(hdr + img0.tobytes() + img1.tobytes() + pts.astype(np.float32).tobytes() + T0.astype(np.float64).tobytes())
    ph = pts.copy(); ph[:, 3] = hdi
    with open(os.path.join(d, "cpp.bin"), "wb") as f:  # adapter_test: points {u, v, idepth, HdiF} + the right image
        f.write(hdr + img0.tobytes() + img1.tobytes() + imgR.tobytes() + ph.astype(np.float32).tobytes() + T0.astype(np.float64).tobytes())
    return dict(dir=d, img0=img0, img1=img1, imgR=imgR, pts=pts, T0=T0)


def run(binary, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([binary, *[str(a) for a in args]], capture_output=True, text=True, env=env, timeout=300)


def parse(stdout):
    out = {}
    for line in stdout.strip().splitlines():
        k, *v = line.split()
        out[k] = np.array([float(x) for x in v])
    return out


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_native_callers_build_link_and_fail_loudly_without_a_gpu(binaries, case):
    c, cpp = binaries
    r = run(c, os.path.join(case["dir"], "c.bin"), 0)
    assert r.returncode == 3 and r.stdout.strip() == "nodevice", (r.returncode, r.stdout, r.stderr)
    r = run(cpp, os.path.join(case["dir"], "cpp.bin"), 0, 1)
    assert r.returncode == 3 and r.stdout.strip() == "nodevice", (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [0, 1])
def test_plain_c_caller_equals_ctypes_path(pkg, binaries, case, variant):
    r = run(binaries[0], os.path.join(case["dir"], "c.bin"), variant)
    assert r.returncode == 0, (r.stdout, r.stderr)
    got = parse(r.stdout)
    ctx = pkg.Context(W_, H_, K_, synth.BASELINE)
    f0, f1 = ctx.frame_create(), ctx.frame_create()
    ctx.make_images(f0, case["img0"]); ctx.make_images(f1, case["img1"])
    ctx.tracker_make_k(K_)
    ctx.tracker_set_ref(f0, case["pts"], (0.0, 0.0))
    ref = ctx.track(f1, case["T0"], (0.0, 0.0), ctx.levels - 1, [np.nan] * 5, variant)
    assert np.array_equal(got["T"].reshape(3, 4), ref["T"])
    assert np.array_equal(got["aff"], np.asarray(ref["aff"], float)) and int(got["ok"][0]) == int(ref["ok"])
    assert np.array_equal(got["res"], ref["lastResiduals"], equal_nan=True)
    assert got["launches"][0] > 0
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("use_right", [0, 1])
def test_cpp_adapter_caller_equals_ctypes_path(pkg, binaries, case, use_right):
    r = run(binaries[1], os.path.join(case["dir"], "cpp.bin"), 0, use_right)
    assert r.returncode == 0, (r.stdout, r.stderr)
    got = parse(r.stdout)
    ctx = pkg.Context(W_, H_, K_, synth.BASELINE)
    f0, f1, fR = ctx.frame_create(), ctx.frame_create(), ctx.frame_create()
    ctx.make_images(f0, case["img0"]); ctx.make_images(f1, case["img1"]); ctx.make_images(fR, case["imgR"])
    ctx.tracker_make_k(K_)
    pts = case["pts"].copy()
    K33 = np.array([K_[0], 0, K_[2], 0, K_[1], K_[3], 0, 0, 1], np.float32)
    if use_right:   # makeCoarseDepthL0's two-way static-stereo re-check (CoarseTracker.cpp:305-347), through the ctypes view
        fwd, _ = ctx.immature_init(f0, pts[:, :2])
        fwd["idepth_min_stereo"] = pts[:, 2] * np.float32(0.1); fwd["idepth_max_stereo"] = pts[:, 2] * np.float32(1.9)
        st = ctx.trace_stereo(fR, K33, True, fwd)
        good = np.nonzero(st == pkg.IPS_GOOD)[0]
        assert good.size > 50
        back, _ = ctx.immature_init(fR, np.ascontiguousarray(fwd["lastTraceUV"][good]))
        back["idepth_min_stereo"] = pts[good, 2] * np.float32(0.1); back["idepth_max_stereo"] = pts[good, 2] * np.float32(1.9)
        ctx.trace_stereo(f0, K33, False, back)
        with np.errstate(divide="ignore"):
            depth = np.float32(1.0) / fwd["idepth_stereo"][good]
        u_delta = np.abs(fwd["u"][good] - back["lastTraceUV"][:, 0])
        take = (u_delta < 1) & (depth > 0) & (depth < 50)
        assert take.sum() > 20 and (~take).sum() > 0
        pts[good[take], 2] = fwd["idepth_stereo"][good[take]]
    ctx.tracker_set_ref(f0, pts, (0.0, 0.0))
    ref = ctx.track(f1, case["T0"], (0.0, 0.0), ctx.levels - 1, [np.nan] * 5, 0)
    assert np.array_equal(got["T"].reshape(3, 4), ref["T"])
    assert np.array_equal(got["res"], ref["lastResiduals"], equal_nan=True) and int(got["ok"][0]) == int(ref["ok"])
    # ImmaturePoint ctor + traceStereo of the first point
    one, _ = ctx.immature_init(f0, case["pts"][:1, :2])
    one["idepth_min_stereo"] = case["pts"][:1, 2] * np.float32(0.1); one["idepth_max_stereo"] = case["pts"][:1, 2] * np.float32(1.9)
    st1 = ctx.trace_stereo(fR, K33, True, one)
    assert int(got["stereo"][2]) == int(st1[0])
    assert np.allclose(got["stereo"][:2], one["lastTraceUV"][0], rtol=1e-7, equal_nan=True)
    # VertexSE3PoseDSO::oplusImpl
    v = ctx.vertex_oplus(pkg.VERTEX_SE3_POSE, case["T0"].reshape(1, 12), np.array([[0.01, -0.02, 0.03, 0.004, -0.005, 0.006]]))
    assert np.array_equal(got["vertex"], v.reshape(12))
    # the two-frame window went through linearizeAll + solveSystemF
    assert got["ba"][0] > 0 and np.isfinite(got["ba"][1]) and got["ba"][1] > 0 and got["ba"][2] > 100
    ctx.close()

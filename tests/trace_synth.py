"""Immature-point test cases shared by the oracle (CPU) and device (GPU) trace tests: a pool of records that drives
traceStereo / traceOn through all of their status branches."""
import numpy as np
import synth


def K33(K4=synth.K4):
    fx, fy, cx, cy = K4
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)


def candidate_pixels(img, n, rng, margin=12, min_grad=8.0):
    """n distinct integer pixels with a usable gradient (what the pixel selector would deliver)."""
    gy, gx = np.gradient(img.astype(np.float64))
    g = np.hypot(gx, gy)
    h, w = img.shape
    ys, xs = np.nonzero(g[margin:h - margin, margin:w - margin] > min_grad)
    pick = rng.choice(ys.size, size=min(n, ys.size), replace=False)
    return np.stack([xs[pick] + margin, ys[pick] + margin], 1).astype(np.float32)


def krki_kt(pose_host, pose_target, K4=synth.K4):
    """hostToFrame_KRKi / hostToFrame_Kt as FullSystem::traceNewCoarse forms them (FullSystem.cpp:760-764), in float."""
    T = synth.T_rel(pose_host, pose_target)
    K = K33(K4).astype(np.float64)
    KRKi = (K @ T[:, :3] @ np.linalg.inv(K)).astype(np.float32)
    Kt = (K @ T[:, 3]).astype(np.float32)
    return KRKi, Kt


def adversarial(pts, rng):
    """Append crafted records: border points (OOB), a too-certain interval (SKIPPED), a gradient orthogonal to the epipolar
    line (BADCONDITION), a wrong colour (OUTLIER, then OOB on the second call), non-finite ranges."""
    extra = pts[:12].copy()
    extra["u"][0] = 5.0; extra["u_stereo"][0] = 5.0                    # uMin just inside / search leaves the image
    extra["u"][1] = 2.0; extra["u_stereo"][1] = 2.0                    # uMin outside
    extra["idepth_min"][2] = 0.05; extra["idepth_max"][2] = 0.0501     # SKIPPED (temporal)
    extra["idepth_min_stereo"][2] = 0.05; extra["idepth_max_stereo"][2] = 0.0501
    extra["gradH"][3] = [0, 0, 0, 400.0]                               # BADCONDITION needs a finite idepth_max
    extra["idepth_min"][3] = 0.02; extra["idepth_max"][3] = 0.08
    extra["idepth_min_stereo"][3] = 0.02; extra["idepth_max_stereo"][3] = 0.08
    extra["color"][4] += 120.0                                         # OUTLIER (twice -> OOB)
    extra["color"][5] -= 90.0
    extra["idepth_min"][6] = 10.0; extra["idepth_max"][6] = 20.0       # absurd inverse depths
    extra["idepth_min_stereo"][6] = 10.0; extra["idepth_max_stereo"][6] = 20.0
    extra["idepth_min"][7] = -0.01; extra["idepth_max"][7] = 0.2       # negative lower bound (scale test bypass)
    extra["gradH"][8] = [1e-12, 0, 0, 1e-12]                           # degenerate gradient matrix
    extra["idepth_min"][9] = 0.01; extra["idepth_max"][9] = 0.3        # long finite segment (> maxPixSearch)
    extra["idepth_min_stereo"][9] = 0.001; extra["idepth_max_stereo"][9] = 0.5
    extra["lastTraceStatus"][10] = 1                                   # already OOB: traceOn returns at once
    extra["energyTH"][11] = np.nan                                     # constructor bailed out
    return np.concatenate([pts, extra])

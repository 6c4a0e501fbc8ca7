#!/usr/bin/env python
"""bench.py — photometric residual+Jacobian evals/s of the CoarseTracker hot path (BASELINE.json configs[1]).

A "step" is one tracked frame at KITTI shape (1232x368 working size, 5 pyramid levels, ~2k template
points dilated to ~10k single-pixel residuals on level 0): FrameHessian::makeImages of the new left
image + CoarseTracker::trackNewestCoarse against a fixed reference keyframe.

  value      : evals/s with the new images already resident in HBM (device-resident inputs)
  e2e        : the same through the C ABI with HOST (pinned) images: H2D copy + makeImages + track + D2H of the result
  roofline   : dominant kernel (the persistent cluster track kernel): algorithmic bytes = evals x 64 B
               (16 B point record + 4 x 12 B gathered texels), duration from CUDA events
               recorded around that launch on its stream, peak = measured HBM copy bandwidth
  cpu_baseline: the oracle port of the same variant (single thread, as the reference tracks), bounded sample

`--impl reference` times the CPU oracle port (the reference itself cannot be compiled here: Eigen,
g2o, Boost, OpenCV are absent — DESIGN.md) on the same workload and prints the same JSON line.
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_EVAL = 64  # tracking eval = 16 B point record + 4 texels x 12 B
N_POINTS = 2000
POOL = 24            # rotating pool of new-frame slots: 24 x (1.8 MB image + 2.4 MB planes + 9.66 MB pyramid) > 126 MB L2
POSES = 12


def load_pkg():
    path = os.path.join(ROOT, "stereo-dso-g2o_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("sdso_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sdso_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(seed_shift=0.0):
    """Reference keyframe (pose 0) + POSES new frames along the path, each with a constant-velocity-like initial guess."""
    import synth
    scene = synth.make_scene()
    p0 = synth.camera_pose(0, seed_shift)
    ref_img, ref_depth = synth.render(scene, p0)
    rng = np.random.default_rng(20260118)
    pts = synth.pick_points(rng, ref_depth, N_POINTS)
    new_imgs, T_true, T_init = [], [], []
    for j in range(POSES):
        k = 0.25 * (j + 1)
        pk = synth.camera_pose(k, seed_shift)
        img, _ = synth.render(scene, pk)
        new_imgs.append(img)
        Tt = synth.T_rel(p0, pk)
        T_true.append(Tt)
        T_init.append(synth.perturb_T(Tt, rng, 0.05, np.deg2rad(0.5)))
    return dict(ref_img=ref_img, pts=pts, new_imgs=new_imgs, T_true=T_true, T_init=T_init)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def cpu_track_loop(wl, variant, min_seconds, max_frames):
    """Oracle port on one host thread: makeImages + trackNewestCoarse per frame. Returns (evals, frames, seconds)."""
    import oracle_py as O
    import synth
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    fref = orc.frame_new()
    orc.make_images(fref, wl["ref_img"])
    orc.tracker_set_ref(fref, wl["pts"])
    fnew = orc.frame_new()
    mr = [np.nan] * 5
    orc.reset_evals()
    t0 = time.perf_counter()
    frames = 0
    while True:
        j = frames % POSES
        orc.make_images(fnew, wl["new_imgs"][j])
        orc.track(fnew, wl["T_init"][j], (0.0, 0.0), orc.levels - 1, mr, variant)
        frames += 1
        el = time.perf_counter() - t0
        if (el >= min_seconds and frames >= POSES) or frames >= max_frames:
            break
    return orc.evals(), frames, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="sse", choices=["sse", "g2o"])
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    variant = 0 if args.variant == "sse" else 1
    W_ = max(args.warmup, 3)
    K_ = args.steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = dict(workload="CoarseTracker pose tracking, 1232x368 (1241x376 cropped), 5-level pyramid, 2000 template points, "
                           f"variant={args.variant}; step = makeImages(new left image) + trackNewestCoarse vs a fixed reference keyframe",
                  points=N_POINTS, levels=5, variant=args.variant,
                  cache="inputs larger than L2: rotating pool of %d new-frame slots (~%d MB)" % (POOL, POOL * 14),
                  parallelism="independent sequences, one per GPU (replicas, no collective)" if args.gpus > 1 else "single sequence")

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = build_workload()
        # warm-up frames, then K bounded samples; each "step" is one tracked frame on the host
        cpu_track_loop(wl, variant, 0.0, W_)
        ev, fr, sec = cpu_track_loop(wl, variant, 0.0, K_)
        val = ev / sec
        line = dict(metric="photometric residual+Jacobian evals/s", value=val, unit="evals/s", n_gpus=args.gpus, steps=K_, warmup=W_,
                    ms_per_step=1e3 * sec / fr, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=config, impl="reference",
                    tracked_frames_per_s=fr / sec,
                    cpu_baseline=dict(value=val, unit="evals/s", cores=1, kind="port",
                                      sample=f"{fr} tracked frames (oracle port of CoarseTracker, 1 thread as in the reference; the reference cannot be compiled here)"),
                    e2e=dict(value=val, unit="evals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------------------------ our arm
    import torch
    import synth
    if not torch.cuda.is_available():
        print(json.dumps(dict(error="no CUDA device: the B200 hot path has no CPU fallback")))
        return 1
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    pkg = load_pkg()
    s = pkg.default_settings()
    s.cluster_size = args.cluster
    s.block_threads = args.threads
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, device=dev, settings=s)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    wl = build_workload(seed_shift=0.37 * rank)  # each rank tracks its own independent sequence
    fref = ctx.frame_create()
    ctx.make_images(fref, wl["ref_img"])
    ctx.tracker_set_ref(fref, wl["pts"])
    slots = [ctx.frame_create() for _ in range(POOL)]
    npx = synth.W * synth.H
    dev_imgs = [torch.from_numpy(wl["new_imgs"][j % POSES]).to(f"cuda:{dev}").contiguous() for j in range(POOL)]
    host_imgs = [torch.from_numpy(wl["new_imgs"][j % POSES]).contiguous().pin_memory() for j in range(POOL)]
    T_init = [np.ascontiguousarray(wl["T_init"][j % POSES]).reshape(1, 12) for j in range(POOL)]
    aff0 = np.zeros((1, 2))
    mr = np.full((1, 5), np.nan)
    coarsest = ctx.levels - 1

    def step_device(i):
        j = i % POOL
        ctx.make_images_device(slots[j], dev_imgs[j].data_ptr(), 1.0, True)
        ctx.track_enqueue([slots[j]], T_init[j], aff0, coarsest, mr, variant)

    def step_host(i):
        j = i % POOL
        ctx.make_images_ptr(slots[j], host_imgs[j].data_ptr(), 1.0, True)
        ctx.track_enqueue([slots[j]], T_init[j], aff0, coarsest, mr, variant)
        return ctx.track_collect(1)  # D2H of the result (pose, aff, residuals) + stream sync

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness guard inside the bench: the tracked poses must be the true motion (not skipped work)
    r = step_host(0)
    err_t = float(np.abs(r["T"][0][:, 3] - wl["T_true"][0][:, 3]).max())
    if not (r["ok"][0] and err_t < 2e-2):
        print(json.dumps(dict(error=f"tracking did not converge in bench (err_t={err_t})")))
        return 1

    # ---- device-resident timing -----------------------------------------------------------------
    for i in range(W_):
        step_device(i)
        ctx.track_collect(1)
    launches0 = ctx.launch_count()
    ctx.profile_enable(True)
    barrier()
    sampler = ClockSampler(dev) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evals = 0
    e0.record(stream)
    for i in range(K_):
        step_device(W_ + i)
        evals += ctx.track_collect(1)["evals"]   # the LM result of frame i gates frame i+1 in a real sequence
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.launch_count() - launches0

    # ---- end-to-end timing (host images, result read back every step) -----------------------------
    for i in range(3):
        step_host(i)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    evals_e2e = 0
    t0.record(stream)
    for i in range(K_):
        evals_e2e += step_host(W_ + i)["evals"]
    t1.record(stream)
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    clocks = sampler.stop() if sampler else None

    # ---- aggregate over ranks (max time, summed work) ---------------------------------------------
    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
        c = torch.tensor([evals, evals_e2e], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        evals, evals_e2e = float(c[0]), float(c[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    value = evals / (ms_dev * 1e-3)
    e2e_val = evals_e2e / (ms_e2e * 1e-3)
    evals_per_launch = (evals / world) / max(prof["track_launches"], 1)
    track_ms_per_launch = prof["track_ms"] / max(prof["track_launches"], 1)
    achieved = evals_per_launch * BYTES_PER_EVAL / (track_ms_per_launch * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"track_{args.variant}_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # CPU baseline (rank 0, bounded sample of the same workload)
    cev, cfr, csec = cpu_track_loop(wl, variant, args.cpu_seconds, 100000)
    line = dict(
        metric="photometric residual+Jacobian evals/s", value=value, unit="evals/s", n_gpus=args.gpus, steps=K_, warmup=W_,
        ms_per_step=ms_dev / K_, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=config,
        tracked_frames_per_s=world * K_ / (ms_dev * 1e-3),
        evals_per_step=evals / (world * K_),
        e2e=dict(value=e2e_val, unit="evals/s", h2d_bytes_per_step=npx * 4, d2h_bytes_per_step=8 * (12 + 2 + 5 + 3) + 4 * 6 + 8,
                 ms_per_step=ms_e2e / K_, tracked_frames_per_s=world * K_ / (ms_e2e * 1e-3)),
        gpu_launches=int(launches),
        clocks=clocks,
        roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                      kernel="track_kernel" if variant == 0 else "track_g2o_kernel", peak_source=peak_src,
                      algorithmic_bytes_per_launch=evals_per_launch * BYTES_PER_EVAL, avg_launch_ms=track_ms_per_launch,
                      share_of_step=prof["track_ms"] / ms_dev,
                      make_images=dict(achieved=(npx * 4 + 16 * 603911) * prof["images_launches"] / max(prof["images_ms"], 1e-9) / 1e6,
                                       unit="GB/s", avg_ms=prof["images_ms"] / max(prof["images_launches"], 1),
                                       algorithmic_bytes=npx * 4 + 16 * 603911)),
        cpu_baseline=dict(value=cev / csec, unit="evals/s", cores=1, kind="port",
                          sample=f"{cfr} tracked frames in {csec:.1f} s (oracle port, 1 thread: trackNewestCoarse is single-threaded in the reference)",
                          tracked_frames_per_s=cfr / csec),
    )
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

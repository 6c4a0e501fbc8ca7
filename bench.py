#!/usr/bin/env python
"""bench.py — photometric residual+Jacobian evals/s of the CoarseTracker hot path (BASELINE.json configs[1]).

A "step" is one new STEREO frame for each of SEQS independent sequences that share the GPU, at KITTI shape
(1232x368 working size, 5 pyramid levels, ~2k template points per keyframe dilated to ~10k single-pixel
residuals on level 0): FrameHessian::makeImages of the new left AND right images (the reference builds both
pyramids of every frame, FullSystem.cpp:1083-1085; batched launches) + CoarseTracker::trackNewestCoarse of every
sequence against its own reference keyframe (one thread-block cluster per sequence, one launch). A single 2k-point frame occupies 8 of the 148 SMs and is bound by the
latency of its ~25 dependent LM evaluations, so throughput is reached by tracking sequences side by side;
the latency of one sequence alone is reported under `single_sequence`.

  value      : evals/s with the new images already resident in HBM (device-resident inputs)
  e2e        : the same through the C ABI with HOST (pinned) images: H2D copy + makeImages + track + D2H of the result
  roofline   : dominant kernel (the persistent cluster track kernel): algorithmic bytes = evals x 64 B
               (16 B point record + 4 x 12 B gathered texels), duration from CUDA events
               recorded around that launch on its stream, peak = measured HBM copy bandwidth
  cpu_baseline: the oracle port of the same variant (single thread, as the reference tracks), bounded sample

Extra keys (`legs`, see bench_legs.py): the g2o tracking variant, the windowed BA at configs 3 and 4, the fork's g2o LBA
driver, the epipolar search (traceOn / traceStereo), the pixel selector, and — under torchrun with N > 1 — the point-sharded
config-4 LM iteration with its allreduce, each with its own roofline / e2e / cpu_baseline.

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
SEQS = 592           # independent sequences per GPU per step (4 x 148: whole waves of 2-CTA clusters at two CTAs per SM)
PATH = 37            # distinct positions along the rendered path; sequence s starts at position s % PATH
# new stereo frames per sequence (cycled); the same at every N — the mix of motions sets the LM iteration counts (3 pose sets read 7 % slower per launch than 6)
POSES = int(os.environ.get("SDSO_BENCH_POSES", "6"))
SETS = 2             # frame-slot sets (double buffering: upload of step i+1 overlaps the kernels of step i)


def load_pkg():
    path = os.path.join(ROOT, "stereo-dso-g2o_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("sdso_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sdso_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def native_oracle():
    """CPU arm: compile the oracle port for THIS host (-O3 -march=native, the reference's own flags, CMakeLists.txt:84) when a
    compiler is present on the box; otherwise the portable x86-64-v3 build that travelled with the snapshot. Must run before
    oracle_py is imported. Returns the flags string for the JSON line."""
    odir = os.path.join(ROOT, "oracle")
    lib = os.path.join(odir, "_build", "liboracle_native.so")
    try:
        subprocess.run(["make", "-C", odir, "native"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
        if os.path.exists(lib):
            os.environ["SDSO_ORACLE_LIB"] = lib
            return "g++ -O3 -march=native -ffp-contract=off (built on this host)"
    except Exception:
        pass
    return "g++ -O3 -march=x86-64-v3 -ffp-contract=off (portable build; no compiler on this host)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(seed_shift=0.0, seqs=SEQS):
    """`seqs` independent sequences cut from one rendered path: sequence s has its reference keyframe at path position
    p = s % PATH, its OWN 2000-point template (drawn independently) and POSES new frames (positions p+1..p+POSES), each with
    a constant-velocity-like initial guess (truth perturbed by 5 cm / 0.5 deg). Sequences that share a path position share
    the rendered source images (rendering is a numpy ray caster, ~0.4 s per frame) but never device memory: every sequence
    has its own template, its own pyramid slots and its own copy of the sources."""
    import synth
    scene = synth.make_scene()
    step = 0.25
    npos = min(seqs, PATH)
    poses = [synth.camera_pose(step * k, seed_shift) for k in range(npos + POSES)]
    rend = [synth.render(scene, p) for p in poses]
    rend_r = [None] + [synth.render(scene, synth.right_of(p))[0] for p in poses[1:]]   # right images of the frames that get tracked
    rng = np.random.default_rng(20260118)
    out = []
    for s in range(seqs):
        p = s % npos
        pts = synth.pick_points(rng, rend[p][1], N_POINTS)
        T_true = [synth.T_rel(poses[p], poses[p + j + 1]) for j in range(POSES)]
        T_init = [synth.perturb_T(T, rng, 0.05, np.deg2rad(0.5)) for T in T_true]
        out.append(dict(ref_img=rend[p][0], pts=pts, new_imgs=[rend[p + j + 1][0] for j in range(POSES)],
                        new_imgs_right=[rend_r[p + j + 1] for j in range(POSES)], T_true=T_true, T_init=T_init, pos=p))
    return out


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
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


def cpu_track_loop(wl, variant, min_seconds, max_frames, threads, warm_frames=3):
    """Oracle port on `threads` host threads, one independent sequence per thread (the reference tracks a sequence on one
    thread; ctypes releases the GIL): makeImages (left + right) + trackNewestCoarse per stereo frame. Returns (evals, frames, seconds)."""
    import threading
    import oracle_py as O
    import synth
    res = [None] * threads
    gate = threading.Barrier(threads)
    t_start = [0.0] * threads
    t_end = [0.0] * threads

    def work(t):
        seq = wl[t % len(wl)]
        orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
        fref = orc.frame_new()
        orc.make_images(fref, seq["ref_img"])
        orc.tracker_set_ref(fref, seq["pts"])
        fnew = orc.frame_new()
        fright = orc.frame_new()
        mr = [np.nan] * 5
        for j in range(warm_frames):   # untimed warm-up on the same buffers (page faults, caches)
            orc.make_images(fnew, seq["new_imgs"][j % POSES])
            orc.make_images(fright, seq["new_imgs_right"][j % POSES])
            orc.track(fnew, seq["T_init"][j % POSES], (0.0, 0.0), orc.levels - 1, mr, variant)
        orc.reset_evals()
        gate.wait()                    # every thread has built its reference before the clock starts
        t0 = time.perf_counter()
        t_start[t] = t0
        frames = 0
        while True:
            j = frames % POSES
            orc.make_images(fnew, seq["new_imgs"][j])
            orc.make_images(fright, seq["new_imgs_right"][j])   # fh_right->makeImages (FullSystem.cpp:1085)
            orc.track(fnew, seq["T_init"][j], (0.0, 0.0), orc.levels - 1, mr, variant)
            frames += 1
            el = time.perf_counter() - t0
            if (el >= min_seconds and frames >= POSES) or frames >= max_frames:
                break
        t_end[t] = time.perf_counter()
        res[t] = (orc.evals(), frames)

    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    sec = max(t_end) - min(t_start)
    return sum(r[0] for r in res), sum(r[1] for r in res), sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="sse", choices=["sse", "g2o"])
    ap.add_argument("--seqs", type=int, default=SEQS)
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cache", action="store_true", help="disable the tracker's shared-memory texel / point cache")
    ap.add_argument("--track-cache", type=int, default=-1, help="tracker shared-memory cache of texel patches / point records: 0 off, 1 on (-1: library default)")
    ap.add_argument("--probe", action="store_true", help="also time the tracker alone on resident pyramids (no makeImages in between)")
    ap.add_argument("--gather", type=int, default=1, help="points in flight per thread (1: 128-register kernel, 2 with --threads 192: 168-register kernel)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--legs", default="all", help="comma list of extra legs on the JSON line: g2o,ba3,ba4,trace,sequence,sharded (or all / none)")
    args = ap.parse_args()
    variant = 0 if args.variant == "sse" else 1
    cpu_flags = native_oracle() if int(os.environ.get("RANK", "0")) == 0 else "n/a"
    W_ = max(args.warmup, 3)
    K_ = args.steps
    S = args.seqs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    host_cores = os.cpu_count() or 1
    config = dict(workload=f"CoarseTracker pose tracking, 1232x368 (1241x376 cropped), 5-level pyramid, 2000 active points splatted per keyframe (makeCoarseDepthL0 dilates them to ~9.9 k template points at level 0), variant={args.variant}; "
                           f"step = one new STEREO frame for each of {S} independent sequences sharing the GPU: makeImages of the left and the right image (8-bit sources, batched) + "
                           "trackNewestCoarse against each sequence's own reference keyframe ("
                           + ("one CTA per sequence, persistent grid of two CTAs per SM pulling sequences from a work counter" if args.cluster == 1 else
                              f"one {args.cluster if args.cluster > 0 else 2}-CTA cluster per sequence, two CTAs per SM") + ", one launch)",
                  points=N_POINTS, levels=5, variant=args.variant, sequences_per_gpu=S,
                  cache=f"inputs larger than L2: {2 * S} new pyramids per step x {SETS} rotating slot sets (~{2 * S * SETS * 12} MB of pyramids + sources), {S} templates",
                  parallelism=(f"{S} independent sequences per GPU; GPUs are replicas (no collective)" if args.gpus > 1 else f"{S} independent sequences on one GPU"))

    # ------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = build_workload(seqs=min(S, host_cores))
        per_thread = max(4 * POSES, 5 * K_)   # bounded sample: 5 tracked frames per requested step on EVERY host thread (seconds of CPU work)
        ev, fr, sec = cpu_track_loop(wl, variant, 1e9, per_thread, host_cores, warm_frames=max(W_, 3))
        val = ev / sec
        line = dict(metric="photometric residual+Jacobian evals/s", value=val, unit="evals/s", n_gpus=args.gpus, steps=K_, warmup=W_,
                    ms_per_step=1e3 * sec / max(K_, 1), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=config, impl="reference",
                    tracked_frames_per_s=fr / sec,
                    cpu_baseline=dict(value=val, unit="evals/s", cores=host_cores, kind="port", build=cpu_flags,
                                      sample=f"{fr} tracked stereo frames (makeImages left + right, trackNewestCoarse) over {host_cores} threads, one sequence per thread (oracle port of CoarseTracker; "
                                             "the reference cannot be compiled here: Eigen/g2o/Boost/OpenCV absent)"),
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
    # one disjoint slice of the host cores per rank: the step's host side (problem records, launch, collect) must not migrate
    # between the ranks' cores while 8 processes share the box
    pinned_cores = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            pinned_cores = len(mine)
        except Exception:
            pinned_cores = None
    pkg = load_pkg()
    st = pkg.default_settings()
    st.cluster_size = args.cluster if args.cluster > 0 else 2   # throughput configuration: a 2-CTA cluster per sequence (148 sequences in flight keep their texel working set in L2) ...
    st.block_threads = args.threads
    st.track_cache = 0 if args.no_cache else (args.track_cache if args.track_cache >= 0 else st.track_cache)
    st.gather_batch = args.gather                                # ... compiled for two resident CTAs per SM
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, device=dev, settings=st)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    wl = build_workload(seed_shift=0.37 * rank, seqs=S)  # each rank tracks its own set of sequences
    for s_, seq in enumerate(wl):
        ctx.tracker_select_ref(s_)
        fref = ctx.frame_create()
        ctx.make_images(fref, seq["ref_img"])
        ctx.tracker_set_ref(fref, seq["pts"])
    slots = [[ctx.frame_create() for _ in range(S)] for _ in range(SETS)]
    slots_r = [[ctx.frame_create() for _ in range(S)] for _ in range(SETS)]   # the right images' pyramids
    slots_lr = [slots[k] + slots_r[k] for k in range(SETS)]
    npx = synth.W * synth.H
    # 8-bit sources (the synthetic renderer quantises to integers, so uint8 is exact): device copies for the resident-input
    # number, pinned host copies for the end-to-end number
    # host side: one contiguous pinned ring-buffer slab [S, H, W] per step pattern j (as a camera ingest ring would hold them),
    # so the library can move a step's sources with a single copy
    host8 = []
    for j in range(POSES):
        slab = torch.empty((2 * S, synth.H, synth.W), dtype=torch.uint8).pin_memory()   # [0, S): left images, [S, 2S): right images
        for s_ in range(S):
            slab[s_].copy_(torch.from_numpy(wl[s_]["new_imgs"][j].astype(np.uint8)))
            slab[S + s_].copy_(torch.from_numpy(wl[s_]["new_imgs_right"][j].astype(np.uint8)))
        host8.append(slab)
    dev8 = [host8[j].to(f"cuda:{dev}") for j in range(POSES)]
    T_init = [np.stack([wl[s_]["T_init"][j].reshape(12) for s_ in range(S)]) for j in range(POSES)]
    aff0 = np.zeros((S, 2))
    mr = np.full((S, 5), np.nan)
    coarsest = ctx.levels - 1
    ref_slots = list(range(S))

    dev_ptrs = [[dev8[j][s_].data_ptr() for s_ in range(2 * S)] for j in range(POSES)]
    host_ptrs = [[host8[j][s_].data_ptr() for s_ in range(2 * S)] for j in range(POSES)]

    def images_device(i):
        ctx.make_images_batch_device(slots_lr[i % SETS], dev_ptrs[i % POSES], u8=True)

    def track_device(i):
        ctx.track_enqueue_multi(ref_slots, slots[i % SETS], T_init[i % POSES], aff0, coarsest, mr, variant)

    def enqueue_device(i):
        images_device(i); track_device(i)

    def upload(i):
        ctx.upload_images_async(slots_lr[i % SETS], host_ptrs[i % POSES], u8=True)

    def images_host(i):
        ctx.make_images_uploaded(slots_lr[i % SETS])

    def track_host(i):
        ctx.track_enqueue_multi(ref_slots, slots[i % SETS], T_init[i % POSES], aff0, coarsest, mr, variant)

    def enqueue_host(i):
        images_host(i); track_host(i)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness guard inside the bench: the tracked poses must be the true motions (no skipped work)
    upload(0); enqueue_host(0)
    r = ctx.track_collect(S)
    # (the fork's live g2o tracker, restated as written, stops after two iterations per level and does not reach the true
    #  motion for every start pose — the CPU restatement behaves the same, tests/test_gpu_tracker.py; there the guard is the
    #  fraction of converged sequences)
    errs = np.array([float(np.abs(r["T"][s_][:, 3] - wl[s_]["T_true"][0][:, 3]).max()) for s_ in range(S)])
    if variant != 0 and (errs < 2e-2).mean() < 0.5:
        print(json.dumps(dict(error=f"g2o-variant tracking converged for only {(errs < 2e-2).mean():.2f} of the sequences")))
        return 1
    for s_ in range(S):
        err_t = errs[s_]
        if variant != 0:
            break
        if not (r["ok"][s_] and err_t < 2e-2):
            print(json.dumps(dict(error=f"tracking did not converge in bench (sequence {s_}, err_t={err_t})")))
            return 1

    # ---- device-resident timing -----------------------------------------------------------------
    for i in range(W_):
        enqueue_device(i)
        ctx.track_collect(S)
    launches0 = ctx.launch_count()
    ctx.profile_enable(True)
    barrier()
    sampler = ClockSampler(dev) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evals = 0
    e0.record(stream)
    images_device(W_)
    for i in range(K_):
        track_device(W_ + i)
        if i + 1 < K_:
            images_device(W_ + i + 1)            # the next frames' pyramids do not depend on this step's poses: queued behind the tracker
        evals += ctx.track_collect(S)["evals"]   # the LM results of step i gate the tracking of step i+1 of every sequence
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.launch_count() - launches0

    probe = None
    if args.probe:   # the tracker on pyramids that stay resident: separates the kernel from whatever the alternating writes cost it
        ctx.profile_enable(True)
        for i in range(K_):
            track_device(W_ + K_ - 1)
            ctx.track_collect(S)
        pp = ctx.profile_read(); ctx.profile_enable(False)
        probe = dict(track_only_ms_per_launch=pp["track_ms"] / max(pp["track_launches"], 1), launches=pp["track_launches"])

    # ---- end-to-end timing: pinned host sources; the upload of step i+1 runs on the copy stream under the kernels of step i
    for i in range(3):
        upload(i); enqueue_host(i); ctx.track_collect(S)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    evals_e2e = 0
    t0.record(stream)
    upload(W_)
    images_host(W_)
    if K_ > 1:
        upload(W_ + 1)
    for i in range(K_):
        track_host(W_ + i)
        if i + 1 < K_:
            images_host(W_ + i + 1)                  # waits for upload i+1 (copy stream), runs behind the tracker of step i
        evals_e2e += ctx.track_collect(S)["evals"]   # D2H of every sequence's result (pose, aff, residuals)
        if i + 2 < K_:
            upload(W_ + i + 2)                       # into the slot set step i just released
    t1.record(stream)
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    # ---- plain H2D bandwidth of this box (one pinned copy of a whole step's sources), for context
    big = host8[0]
    dst = torch.empty_like(big, device=f"cuda:{dev}")
    dst.copy_(big, non_blocking=True); torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for _ in range(5):
        dst.copy_(big, non_blocking=True)
    b1.record(stream)
    torch.cuda.synchronize()
    h2d_gbs = 5 * big.numel() / (b0.elapsed_time(b1) * 1e-3) / 1e9
    # ... and with EVERY rank copying at the same time: what the host (memory + PCIe root) delivers to N GPUs at once — the
    # ceiling of the end-to-end number at N > 1
    h2d_gbs_all = h2d_gbs
    if dist is not None:
        barrier()
        b0.record(stream)
        for _ in range(5):
            dst.copy_(big, non_blocking=True)
        b1.record(stream)
        torch.cuda.synchronize()
        tq = torch.tensor([5 * big.numel() / (b0.elapsed_time(b1) * 1e-3) / 1e9], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(tq, op=dist.ReduceOp.SUM)
        h2d_gbs_all = float(tq[0])
    del dst

    # ---- single-sequence latency (one cluster on the GPU), for information
    lat_n = min(K_, 50)
    one_lr = [slots[0][0], slots_r[0][0]]

    def one_frame(i):
        ctx.make_images_batch_device(one_lr, [dev8[i % POSES][0].data_ptr(), dev8[i % POSES][S].data_ptr()], u8=True)
        ctx.track_enqueue_multi([0], slots[0][:1], T_init[i % POSES][:1], aff0[:1], coarsest, mr[:1], variant); ctx.track_collect(1)

    for i in range(3):
        one_frame(i)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(stream)
    for i in range(lat_n):
        one_frame(i)
    l1.record(stream)
    torch.cuda.synchronize()
    ms_single = l0.elapsed_time(l1) / lat_n
    clocks = sampler.stop() if sampler else None

    # ---- legs: the rest of the path, driver-visible (bench_legs.py) -----------------------------------------
    def traffic_of(vname):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"track_{vname}_batch{S}_dram_bytes_per_launch")
        except Exception:
            return None

    want = set(["g2o", "ba3", "ba4", "trace", "sequence", "sharded"] if args.legs == "all" else [x for x in args.legs.split(",") if x and x != "none"])
    legs = {}
    peak_, _ = measured_peak()
    if "g2o" in want and variant == 0 and world == 1:   # the fork's LIVE tracker (EdgeSE3PosePhotoDSO + restated g2o LM) on the same workload
        Kg = max(3, min(K_, 20))
        for i in range(3):
            ctx.track_enqueue_multi(ref_slots, slots[i % SETS], T_init[i % POSES], aff0, coarsest, mr, 1); ctx.track_collect(S)
        ctx.profile_enable(True); torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_g = 0
        g0.record(stream)
        images_device(0)
        for i in range(Kg):
            ctx.track_enqueue_multi(ref_slots, slots[i % SETS], T_init[i % POSES], aff0, coarsest, mr, 1)
            if i + 1 < Kg:
                images_device(i + 1)
            rr = ctx.track_collect(S)
            ev_g += rr["evals"]
        g1.record(stream); torch.cuda.synchronize()
        pg = ctx.profile_read(); ctx.profile_enable(False)
        ms_g = g0.elapsed_time(g1)
        conv = float((np.array([np.abs(rr["T"][s_][:, 3] - wl[s_]["T_true"][(Kg - 1) % POSES][:, 3]).max() for s_ in range(S)]) < 2e-2).mean())
        tk = pg["track_ms"] / max(pg["track_launches"], 1)
        ach = ev_g / Kg * BYTES_PER_EVAL / (tk * 1e-3) / 1e9
        cg_ev, cg_fr, cg_sec = cpu_track_loop(wl, 1, min(args.cpu_seconds, 5.0), 100000, host_cores)
        legs["track_g2o"] = dict(
            workload=f"the same {S} sequences, variant=g2o (the fork's live code path: EdgeSE3PosePhotoDSO edges + restated g2o LM, 2 iterations per level)",
            value=ev_g / (ms_g * 1e-3), unit="evals/s", steps=Kg, ms_per_step=ms_g / Kg, tracked_frames_per_s=S * Kg / (ms_g * 1e-3), converged_fraction=conv,
            note="compare tracked_frames_per_s with the CPU arm, not evals/s: g2o's LM evaluates every edge up to 7 times per level (the CPU port counts "
                 "those), the kernel executes 3 of them when every damping trial is accepted — the others re-evaluate the same edges at the same estimate "
                 "(bit-identical errors, Jacobians and sums; tracker_g2o.cuh) — so `value` counts fewer evaluations per frame than the CPU arm does",
            roofline=dict(bound="hbm", achieved=ach, peak=peak_, unit="GB/s", frac=ach / peak_, kernel="track_g2o_kernel", avg_launch_ms=tk, traffic=traffic_of("g2o"),
                          algorithmic_bytes_per_launch=ev_g / Kg * BYTES_PER_EVAL),
            cpu_baseline=dict(value=cg_ev / cg_sec, unit="evals/s", cores=host_cores, kind="port", tracked_frames_per_s=cg_fr / cg_sec,
                              sample=f"{cg_fr} tracked stereo frames in {cg_sec:.1f} s over {host_cores} threads (oracle port, g2o variant)"))
    import bench_legs as BL

    def guarded(name, fn):
        """a leg that fails must not take the headline line with it: the error is recorded in its place"""
        try:
            return fn()
        except Exception as ex:   # noqa: BLE001
            import traceback
            print(f"[bench] leg {name} failed:\n{traceback.format_exc()}", file=sys.stderr)
            return dict(error=f"{type(ex).__name__}: {ex}")

    if world == 1 and (want & {"ba3", "ba4", "trace", "sequence"}):
        scene = synth.make_scene()
        lk = max(10, min(K_, 50))
        if "ba3" in want:
            legs["ba_config3"] = guarded("ba3", lambda: BL.leg_ba(pkg, torch, dev, scene, "config3", lk, peak_, want_g2o=True))
            legs["lba_g2o"] = legs["ba_config3"].pop("lba_g2o", None)
        if "ba4" in want:
            legs["ba_config4"] = guarded("ba4", lambda: BL.leg_ba(pkg, torch, dev, scene, "config4", lk, peak_, cpu_seconds=3.0))
        if "trace" in want:
            tr = guarded("trace", lambda: BL.leg_trace(pkg, torch, dev, scene, peak_))
            legs.update({"trace_on": tr} if "error" in tr else tr)
        if "sequence" in want:
            legs["sequence"] = guarded("sequence", lambda: BL.leg_sequence(pkg, torch, dev, scene))
    if world > 1 and "sharded" in want:
        legs["sharded_ba"] = guarded("sharded", lambda: BL.leg_sharded_ba(pkg, torch, dist, dev, rank, world, synth.make_scene(), max(10, min(K_, 50)), peak_))

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
            traffic = json.load(open(tp)).get(f"track_{args.variant}_batch{S}_dram_bytes_per_launch")
        except Exception:
            traffic = None
    # what HBM delivers for this access pattern with no arithmetic at all (tools/gather_microbench.cu), measured on this GPU now
    pattern = None
    gm = os.path.join(ROOT, "tools", "gather_microbench")
    if rank == 0 and os.path.exists(gm):
        try:
            import subprocess
            pattern = json.loads(subprocess.run([gm, "--json"], capture_output=True, text=True, timeout=60,
                                                env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(dev))).stdout.strip().splitlines()[-1])
            pattern["kernel_over_microbench"] = achieved / pattern["gather_gbs"]   # a data point, not a roofline: the kernel re-reads on chip and ends up above it
            pattern["note"] = "tools/gather_microbench.cu: the same 2x2-texel gather pattern with no arithmetic; the roofline anchor is MEASURED_PEAKS.json (frac above)"
        except Exception:
            pattern = None
    img_bytes = npx * 1 + 16 * 603911  # per image: u8 read + texels written
    # CPU baseline (rank 0, bounded sample of the same workload on all host cores)
    if world == 1:
        cev, cfr, csec = cpu_track_loop(wl, variant, args.cpu_seconds, 100000, host_cores)
    line = dict(
        metric="photometric residual+Jacobian evals/s", value=value, unit="evals/s", n_gpus=args.gpus, steps=K_, warmup=W_,
        ms_per_step=ms_dev / K_, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=config,
        tracked_frames_per_s=world * S * K_ / (ms_dev * 1e-3),
        evals_per_step=evals / (world * K_),
        e2e=dict(value=e2e_val, unit="evals/s", h2d_bytes_per_step=2 * S * npx, d2h_bytes_per_step=S * (8 * (12 + 2 + 5 + 3) + 4 * 6 + 8),
                 ms_per_step=ms_e2e / K_, tracked_frames_per_s=world * S * K_ / (ms_e2e * 1e-3),
                 h2d_gbs_plain_copy=h2d_gbs, h2d_gbs_in_step=2 * S * npx / (ms_e2e / K_ * 1e-3) / 1e9,
                 h2d_gbs_all_ranks_plain_copy=h2d_gbs_all, h2d_gbs_all_ranks_in_step=world * 2 * S * npx / (ms_e2e / K_ * 1e-3) / 1e9,
                 limiter="host link: the step moves 2 x S 8-bit images per GPU; compare h2d_gbs_all_ranks_in_step with h2d_gbs_all_ranks_plain_copy",
                 host_threads_pinned=pinned_cores),
        single_sequence=dict(ms_per_frame=ms_single, tracked_frames_per_s=1e3 / ms_single,
                             note="latency of one sequence alone on the GPU in this (throughput) configuration; the latency configuration (8-CTA cluster, gather batch 2) tracks a frame in ~0.2 ms, profiles/r1_bench_sse_first.json"),
        gpu_launches=int(launches),
        clocks=clocks,
        roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                      traffic_source="ncu --set full capture of this kernel in this command, profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch); not measurable inside an unprofiled run",
                      kernel="track_kernel" if variant == 0 else "track_g2o_kernel", peak_source=peak_src,
                      algorithmic_bytes_per_launch=evals_per_launch * BYTES_PER_EVAL, avg_launch_ms=track_ms_per_launch,
                      share_of_step=prof["track_ms"] / ms_dev, gather_microbench=pattern,
                      make_images=dict(achieved=img_bytes * 2 * S * K_ / max(prof["images_ms"], 1e-9) / 1e6, frac=img_bytes * 2 * S * K_ / max(prof["images_ms"], 1e-9) / 1e6 / peak,
                                       images_per_step=2 * S,
                                       unit="GB/s", avg_ms=prof["images_ms"] / max(prof["images_launches"], 1),
                                       algorithmic_bytes=img_bytes)),
        probe=probe,
        legs=legs,
        cpu_baseline=(dict(value=cev / csec, unit="evals/s", cores=host_cores, kind="port", build=cpu_flags,
                           sample=f"{cfr} tracked stereo frames (makeImages left + right, trackNewestCoarse) in {csec:.1f} s over {host_cores} threads, one sequence per thread (oracle port; "
                                  "trackNewestCoarse is single-threaded per sequence in the reference)",
                           tracked_frames_per_s=cfr / csec) if world == 1 else None),
    )
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

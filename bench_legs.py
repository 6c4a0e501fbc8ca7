"""bench_legs.py — the rest of the photometric path measured driver-visibly (extra keys on bench.py's JSON line).

Every leg reports: the device number (CUDA events on the stream the kernels run on, inputs resident), a `roofline`
{achieved, peak, frac} from SURVEY.md §8d's algorithmic bytes per unit (58.5 B per BA eval, 128 B per epipolar search step,
64 B per tracking eval), `e2e` through the C ABI with HOST buffers (wall clock around the public call, uploads and read-backs
inside), and `cpu_baseline` = the oracle port on the host (bounded sample, thread count stated). The oracle is only the CPU arm
and the parity check here, never on the product path.

  ba_config3 / ba_config4 : one LM iteration of the windowed BA (linearizeAll + top/SC accumulation + stitch + solve +
                            back-substitution) and FullSystem::optimize (6 iterations) — 7 KF / 2002 points / 1232x368 and
                            10 KF / 20 000 points / 1920x1088
  lba_g2o                 : the fork's live LBA driver (E2 edges + restated g2o LM, 3 iterations) at config 3
  trace_on / trace_stereo : epipolar search of 1500 immature points per host frame x 6 hosts into the newest key frame; static
                            stereo of the same points into the right image
  make_maps               : PixelSelector::makeMaps on a 1232x368 frame
  sharded_ba (N > 1)      : config 4 LM iteration with the points sharded over the ranks and ONE allreduce of the reduced system
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

BA_BYTES_PER_EVAL = 58.5       # 4 x 12 B gather + colour + weight + (point + ids) / 8
TRACE_BYTES_PER_STEP = 128.0   # 8 pattern pixels x 4 taps x 4 B
DENSE_K = (1100.0, 1100.0, 959.5, 543.5)


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def make_ba_case(scene, name):
    import ba_synth, synth
    if name == "config3":
        c = dict(n=7, P=2002, w=synth.W, h=synth.H, K=synth.K4, spacing=0.35, seed=7)
    else:
        c = dict(n=10, P=20000, w=1920, h=1088, K=DENSE_K, spacing=0.3, seed=12)
    win = ba_synth.make_window(scene, n=c["n"], P=c["P"], seed=c["seed"], spacing=c["spacing"], w=c["w"], h=c["h"], K=c["K"])
    return c, win


def device_window(pkg, ctx, win, fids, begin=0, end=None, rank=0, world=1):
    """Window of points [begin, end) of allPoints (colour / weights from the device D1 operator)."""
    pts = win["points"][begin:end]
    W = pkg.Window(ctx)
    for k, f in enumerate(win["frames"]):
        idx = W.add_frame(fids[k], f["T_w2c"], f["a"], f["b"], f["frameID"])
        W.set_state(idx, f["state"]); W.set_energy_th(idx, f["energyTH"])
    col = np.zeros((len(pts), 8), np.float32); wts = np.zeros((len(pts), 8), np.float32)
    host = np.array([p["host"] for p in pts], np.int32)
    uv = np.array([[p["u"], p["v"]] for p in pts], np.float32).reshape(-1, 2)
    for h in range(win["n"]):
        sel = np.nonzero(host == h)[0]
        if sel.size == 0:
            continue
        rec, _ = ctx.immature_init(fids[h], uv[sel])
        col[sel] = rec["color"]; wts[sel] = rec["weights"]
    args = (host, uv[:, 0].copy(), uv[:, 1].copy(), np.array([p["idepth"] for p in pts], np.float32), np.array([p["idepth_zero"] for p in pts], np.float32),
            col, wts, np.array([p["has_prior"] for p in pts], np.uint8))
    W.set_points(*args)
    rp = np.repeat(np.arange(len(pts), dtype=np.int32), [len(p["targets"]) for p in pts])
    rt = np.array([t for p in pts for t in p["targets"]], np.int32)
    W.set_residuals(rp, rt)
    W.set_shard(rank, world)
    W.prepare()
    return W, args, (rp, rt)


def leg_ba(pkg, torch, dev, scene, name, iters, peak, cpu_seconds=6.0, want_g2o=False):
    import oracle_py as O, oracle_ba_py as OB, ba_synth, synth
    c, win = make_ba_case(scene, name)
    n = c["n"]
    ctx = pkg.Context(c["w"], c["h"], c["K"], synth.BASELINE, device=dev)
    stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
    fids = []
    for f in win["frames"]:
        fid = ctx.frame_create(); ctx.make_images(fid, f["image"]); fids.append(fid)
    W, pargs, (rp, rt) = device_window(pkg, ctx, win, fids)
    R = W.counts()["res"]
    l0 = ctx.launch_count()

    def lm_iteration():
        W.linearize_all_async(True); W.assemble(); W.solve_assembled(2, want=False)

    for _ in range(5):
        lm_iteration()
    torch.cuda.synchronize()
    launches_per_it = (ctx.launch_count() - l0) / 5
    e0, e1 = _events(torch)
    e0.record(stream)
    for _ in range(iters):
        lm_iteration()
    e1.record(stream); torch.cuda.synchronize()
    ms_it = e0.elapsed_time(e1) / iters
    # linearize alone (the gather kernel: 58.5 B per eval is ITS algorithmic traffic)
    e0.record(stream)
    for _ in range(iters):
        W.linearize_all_async(True)
    e1.record(stream); torch.cuda.synchronize()
    ms_lin = e0.elapsed_time(e1) / iters
    evals = 8 * R
    bytes_it = evals * BA_BYTES_PER_EVAL

    # FullSystem::optimize through the ABI with host buffers: window upload + optimize(6) + state read-back, wall clock
    def e2e_once():
        t0 = time.perf_counter()
        W2 = pkg.Window(ctx)
        for k, f in enumerate(win["frames"]):
            idx = W2.add_frame(fids[k], f["T_w2c"], f["a"], f["b"], f["frameID"])
            W2.set_state(idx, f["state"]); W2.set_energy_th(idx, f["energyTH"])
        W2.set_points(*pargs); W2.set_residuals(rp, rt); W2.prepare()
        t1 = time.perf_counter()
        rmse, its = W2.optimize(6)
        t2 = time.perf_counter()
        st = W2.get_state()
        t3 = time.perf_counter()
        return (t1 - t0, t2 - t1, t3 - t2, rmse, its, st)

    e2e_once()
    runs = [e2e_once() for _ in range(3)]
    up, opt, down = (float(np.median([r[k] for r in runs])) for k in range(3))
    rmse_dev, its_dev = runs[-1][3], runs[-1][4]
    h2d = sum(a.nbytes for a in pargs) + rp.nbytes + rt.nbytes + n * (12 + 10 + 3) * 8
    d2h = n * (10 + 12) * 8 + len(win["points"]) * 4 + 32
    lin_calls = its_dev + 2   # initial linearizeAll + one per iteration + the final one (FullSystemOptimize.cpp:889-1010)

    # CPU port, one thread (the reference runs these operators on 6 workers; the port is single-threaded — stated as cores=1)
    orc = O.Oracle(c["w"], c["h"], c["K"], synth.BASELINE)
    ba, _, cw = ba_synth.fill_oracle(win, orc, OB.OracleBA, OB.immature_init)
    ba.linearize_all(True); x, _, _ = ba.solve(2); ba.resubstitute(x)
    reps, t0 = 0, time.perf_counter()
    while True:
        ba.linearize_all(True); x, _, _ = ba.solve(2); ba.resubstitute(x)
        reps += 1
        if time.perf_counter() - t0 > cpu_seconds or reps >= 50:
            break
    ms_cpu = 1e3 * (time.perf_counter() - t0) / reps
    t0 = time.perf_counter(); rmse_cpu, its_cpu = ba.optimize(6); opt_cpu = time.perf_counter() - t0
    out = dict(
        workload=f"windowed BA, {n} key frames, {len(win['points'])} points, {R} residuals, {c['w']}x{c['h']} (SURVEY config {'3' if name == 'config3' else '4'})",
        metric="BA residual+Jacobian evals/s (one LM iteration: linearizeAll + accumulate + stitch + solve + resubstitute)",
        value=evals / (ms_it * 1e-3), unit="evals/s", ms_per_lm_iteration=ms_it, ms_linearize=ms_lin, kernel_launches_per_iteration=launches_per_it,
        roofline=dict(bound="hbm", achieved=bytes_it / (ms_it * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=bytes_it / (ms_it * 1e-3) / 1e9 / peak,
                      algorithmic_bytes_per_iteration=bytes_it, kernel="whole LM iteration (launch chain)",
                      linearize_only=dict(achieved=bytes_it / (ms_lin * 1e-3) / 1e9, frac=bytes_it / (ms_lin * 1e-3) / 1e9 / peak, kernel="ba_linearize_kernel")),
        optimize=dict(iterations=int(its_dev), rmse=rmse_dev, ms_wall=1e3 * opt, ms_wall_per_iteration=1e3 * opt / max(its_dev, 1),
                      cpu_port_ms_wall=1e3 * opt_cpu, cpu_iterations=int(its_cpu), cpu_rmse=rmse_cpu),
        e2e=dict(value=evals * lin_calls / (up + opt + down), unit="evals/s", ms_upload=1e3 * up, ms_optimize=1e3 * opt, ms_readback=1e3 * down,
                 h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                 note="host SoA window -> sdso_ba_set_points/_residuals/_prepare -> sdso_ba_optimize(6) -> sdso_ba_get_state, wall clock"),
        cpu_baseline=dict(value=evals / (ms_cpu * 1e-3), unit="evals/s", cores=1, kind="port", ms_per_lm_iteration=ms_cpu,
                          sample=f"{reps} LM iterations of the same window on one host thread (oracle port)"))
    if want_g2o:
        st0 = ba.get_state()
        T_wh = np.stack([np.hstack([T[:, :3].T, (-T[:, :3].T @ T[:, 3])[:, None]]) for T in W.get_state()["T_w2c"]])
        rng = np.random.default_rng(0)
        Tp = np.stack([synth.perturb_T(T, rng, 3e-3, 3e-4) for T in T_wh])
        idp = np.array([float(p["idepth"]) for p in win["points"] for _ in p["targets"]])
        Kd = np.array(c["K"], float)
        W3, _, _ = device_window(pkg, ctx, win, fids)
        orc3 = O.Oracle(c["w"], c["h"], c["K"], synth.BASELINE)
        ba3, _, _ = ba_synth.fill_oracle(win, orc3, OB.OracleBA, OB.immature_init)
        W3.lba_g2o(Kd, Tp, np.zeros((n, 2)), idp, 3)
        ts = []
        for _ in range(7):
            l1 = ctx.launch_count()
            t0 = time.perf_counter(); g = W3.lba_g2o(Kd, Tp, np.zeros((n, 2)), idp, 3); ts.append(time.perf_counter() - t0)
            lba_launches = ctx.launch_count() - l1
        t0 = time.perf_counter(); o = ba3.lba_g2o(Kd, Tp, np.zeros((n, 2)), idp, 3); t_cpu = time.perf_counter() - t0
        ev = 8 * R * (g["trials"] + g["iterations"] + 1)
        t_dev = float(np.median(ts))   # (a host-driven loop of ~200 launches and ~50 synchronisations: the spread between calls is reported too)
        out["lba_g2o"] = dict(
            workload="FullSystem::optimize, g2o body (E2 edges + restated g2o LM + Schur over per-residual idepth vertices), 3 LM iterations, same window",
            ms_wall=1e3 * t_dev, ms_wall_min=1e3 * float(np.min(ts)), ms_wall_max=1e3 * float(np.max(ts)), iterations=int(g["iterations"]), trials=int(g["trials"]), chi2=float(g["chi2"]), kernel_launches=int(lba_launches),
            value=ev / t_dev, unit="evals/s",
            roofline=dict(bound="hbm", achieved=ev * BA_BYTES_PER_EVAL / t_dev / 1e9, peak=peak, unit="GB/s", frac=ev * BA_BYTES_PER_EVAL / t_dev / 1e9 / peak,
                          note="wall clock of the whole driver (host round trips included): latency-bound"),
            e2e=dict(value=ev / t_dev, unit="evals/s", h2d_bytes_per_step=int(idp.nbytes + Tp.nbytes + 16 * n + 32), d2h_bytes_per_step=int(idp.nbytes + Tp.nbytes + 16 * n + R * 20),
                     note="the call takes and returns host arrays; value is the same wall clock"),
            cpu_baseline=dict(value=8 * R * (o["trials"] + o["iterations"] + 1) / t_cpu, unit="evals/s", cores=1, kind="port", ms_wall=1e3 * t_cpu,
                              iterations=int(o["iterations"]), chi2=float(o["chi2"]), sample="one run of the restated driver on one host thread"),
            parity=dict(iterations_equal=bool(g["iterations"] == o["iterations"]), chi2_rel=abs(g["chi2"] - o["chi2"]) / abs(o["chi2"])))
    ctx.close()
    return out


def leg_trace(pkg, torch, dev, scene, peak, per_host=1500, n_kf=7, reps=10):
    import oracle_py as O, oracle_trace_py as OT, synth, trace_synth as TS
    poses = [synth.camera_pose(0.5 * k) for k in range(n_kf)]
    imgs = [synth.render(scene, p) for p in poses]
    right = synth.render(scene, synth.right_of(poses[-1]))
    ctx = pkg.Context(synth.W, synth.H, synth.K4, synth.BASELINE, device=dev)
    stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
    orc = O.Oracle(synth.W, synth.H, synth.K4, synth.BASELINE)
    gf, of = [], []
    for im, _ in imgs + [right]:
        g, o = ctx.frame_create(), orc.frame_new()
        ctx.make_images(g, im); orc.make_images(o, im)
        gf.append(g); of.append(o)
    newest = n_kf - 1
    rng = np.random.default_rng(12)
    cases = []
    for h in range(newest):
        uv = TS.candidate_pixels(imgs[h][0], per_host, rng, margin=16, min_grad=6.0)
        pts, ok = ctx.immature_init(gf[h], uv)
        pts = pts[ok]
        tid = 1.0 / imgs[h][1][pts["v"].astype(int), pts["u"].astype(int)]
        pts["idepth_min"] = (tid * 0.5).astype(np.float32); pts["idepth_max"] = (tid * 2.0).astype(np.float32)
        KRKi, Kt = TS.krki_kt(poses[h], poses[newest], synth.K4)
        cases.append((h, pts, KRKi, Kt))
    npts = sum(c[1].size for c in cases)
    allp = np.ascontiguousarray(np.concatenate([c[1] for c in cases]))
    host_of = np.concatenate([np.full(c[1].size, i, np.int32) for i, c in enumerate(cases)])
    KR, KT = np.stack([c[2] for c in cases]), np.stack([c[3] for c in cases])
    AF = np.tile(np.array([[1.0, 0.0]], np.float32), (len(cases), 1))

    def run_host():   # host records: upload, ONE launch for all hosts, read back
        p = allp.copy()
        ctx.trace_on_hosts(gf[newest], KR, KT, AF, host_of, p, want_status=False)
        return p

    def run_per_host():   # the reference's shape: one call per host key frame
        for h, pts, KRKi, Kt in cases:
            ctx.trace_on(gf[newest], KRKi, Kt, (1.0, 0.0), pts.copy())

    pg_all = run_host()
    steps = int(np.maximum(pg_all["numSteps"], 0).sum())
    t0 = time.perf_counter()
    for _ in range(reps):
        run_host()
    t_dev = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        run_per_host()
    t_per_host = (time.perf_counter() - t0) / reps
    # device-resident records: the kernel alone between CUDA events (every repetition traces the same freshly uploaded records)
    e0, e1 = _events(torch)
    ms_k = []
    for _ in range(reps):
        ctx.immature_upload(allp)
        l0 = ctx.launch_count()
        e0.record(stream)
        ctx.trace_on_hosts(gf[newest], KR, KT, AF, host_of, None, want_status=False)
        e1.record(stream); torch.cuda.synchronize()
        ms_k.append(e0.elapsed_time(e1))
        launches = ctx.launch_count() - l0
    ms_ev = float(np.median(ms_k))
    resident_equal = ctx.immature_download(allp.size).tobytes() == pg_all.tobytes()
    # oracle on one thread + parity
    po_all = []
    t0 = time.perf_counter()
    for h, pts, KRKi, Kt in cases:
        po = pts.copy()
        OT.trace_on(orc, of[newest], KRKi, Kt, (1.0, 0.0), po)
        po_all.append(po)
    t_cpu = time.perf_counter() - t0
    po_all = np.concatenate(po_all)
    same = bool(np.array_equal(po_all["lastTraceStatus"], pg_all["lastTraceStatus"]) and np.array_equal(po_all["bestIdx"], pg_all["bestIdx"])
                and np.array_equal(po_all["numSteps"], pg_all["numSteps"]))
    rec_bytes = pkg.IMMATURE_DTYPE.itemsize
    out = dict(
        trace_on=dict(
            workload=f"ImmaturePoint::traceOn, {npts} immature points of {newest} host key frames into the newest key frame (traceNewCoarse), 1232x368, prior interval "
                     "[0.5, 2] x true idepth; one warp per point, all hosts in ONE launch",
            metric="epipolar search steps/s", value=steps / (ms_ev * 1e-3), unit="steps/s", steps=int(steps), points=int(npts), ms=ms_ev, kernel_launches=int(launches),
            roofline=dict(bound="hbm", achieved=steps * TRACE_BYTES_PER_STEP / (ms_ev * 1e-3) / 1e9, peak=peak, unit="GB/s",
                          frac=steps * TRACE_BYTES_PER_STEP / (ms_ev * 1e-3) / 1e9 / peak, kernel="trace_kernel<false>",
                          note="device-resident records, CUDA events around the one launch (+ the 1 KB transform upload); latency-bound: "
                               "one warp's dependent chain (set-up, <= 2 search rounds, <= 3 refinement round trips) with < 1 wave of warps"),
            e2e=dict(value=steps / t_dev, unit="steps/s", ms_wall=1e3 * t_dev, ms_wall_one_call_per_host=1e3 * t_per_host, h2d_bytes_per_step=int(npts * rec_bytes + npts * 4),
                     d2h_bytes_per_step=int(npts * rec_bytes), note="sdso_trace_on_hosts with host records (upload + launch + read-back), wall clock"),
            cpu_baseline=dict(value=steps / t_cpu, unit="steps/s", cores=1, kind="port", ms_wall=1e3 * t_cpu, sample="the same points once on one host thread (oracle port)"),
            parity=dict(status_bestIdx_numSteps_equal=same, resident_equals_host_records=bool(resident_equal))))
    # static stereo of the newest frame's own candidates into the right image
    uv = TS.candidate_pixels(imgs[newest][0], per_host * 2, rng, margin=16, min_grad=6.0)
    pts, ok = ctx.immature_init(gf[newest], uv)
    pts = pts[ok]
    K33 = TS.K33(synth.K4)
    pts["idepth_min_stereo"] = 0.0; pts["idepth_max_stereo"] = np.nan

    def run_st():
        p = pts.copy()
        ctx.trace_stereo(gf[-1], K33, True, p)
        return int(np.maximum(p["numSteps"], 0).sum())

    steps = run_st()
    e0.record(stream)
    for _ in range(reps):
        run_st()
    e1.record(stream); torch.cuda.synchronize()
    ms_ev = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        run_st()
    t_dev = (time.perf_counter() - t0) / reps
    po = pts.copy()
    t0 = time.perf_counter(); so = OT.trace_stereo(orc, of[-1], K33, True, po); t_cpu = time.perf_counter() - t0
    pg = pts.copy(); sg = ctx.trace_stereo(gf[-1], K33, True, pg)
    out["trace_stereo"] = dict(
        workload=f"ImmaturePoint::traceStereo (left -> right, unbounded prior), {pts.size} candidates of one key frame, 1232x368",
        metric="epipolar search steps/s", value=steps / (ms_ev * 1e-3), unit="steps/s", steps=int(steps), points=int(pts.size), ms=ms_ev,
        roofline=dict(bound="hbm", achieved=steps * TRACE_BYTES_PER_STEP / (ms_ev * 1e-3) / 1e9, peak=peak, unit="GB/s",
                      frac=steps * TRACE_BYTES_PER_STEP / (ms_ev * 1e-3) / 1e9 / peak, kernel="trace_kernel<true>"),
        e2e=dict(value=steps / t_dev, unit="steps/s", ms_wall=1e3 * t_dev, h2d_bytes_per_step=int(pts.size * rec_bytes), d2h_bytes_per_step=int(pts.size * (rec_bytes + 4))),
        cpu_baseline=dict(value=steps / t_cpu, unit="steps/s", cores=1, kind="port", ms_wall=1e3 * t_cpu, sample="the same points once on one host thread (oracle port)"),
        parity=dict(status_and_bestIdx_equal=bool(np.array_equal(so, sg) and np.array_equal(po["bestIdx"], pg["bestIdx"]))))
    # pixel selector on the newest frame
    import oracle_select_py as S
    sel = S.Selector(orc)
    ctx.selector_reset(); ctx.make_maps(gf[newest], 2000.0, want_map=False)
    e0.record(stream)
    for _ in range(reps):
        ctx.selector_reset(); ctx.selector_make_hists(gf[newest])
        _, ng = ctx.make_maps(gf[newest], 2000.0, want_map=False)
    e1.record(stream); torch.cuda.synchronize()
    ms_sel = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.selector_reset(); ctx.selector_make_hists(gf[newest])
        m, ng = ctx.make_maps(gf[newest], 2000.0, want_map=True)
    t_sel = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(3):
        sel.potential(3); sel.forget_hist()
        mo, no = sel.make_maps(of[newest], 2000.0)
    t_cpu = (time.perf_counter() - t0) / 3
    px = synth.W * synth.H
    out["make_maps"] = dict(
        workload="PixelSelector::makeMaps (histograms + select + re-select + sub-sampling to density 2000), 1232x368",
        metric="selector pixels/s", value=px / (ms_sel * 1e-3), unit="pixels/s", ms=ms_sel, selected=int(ng),
        roofline=dict(bound="hbm", achieved=px * 32 / (ms_sel * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=px * 32 / (ms_sel * 1e-3) / 1e9 / peak,
                      note="16 B texel per pixel x 2 passes; the operator is a chain of latency-bound launches"),
        e2e=dict(value=px / t_sel, unit="pixels/s", ms_wall=1e3 * t_sel, h2d_bytes_per_step=0, d2h_bytes_per_step=int(px * 4), note="selection map read back to the host"),
        cpu_baseline=dict(value=px / t_cpu, unit="pixels/s", cores=1, kind="port", ms_wall=1e3 * t_cpu, sample="3 calls on one host thread (oracle port)"),
        parity=dict(map_equal=bool(np.array_equal(m, mo)) and int(ng) == int(no)))
    ctx.close()
    return out


def leg_sequence(pkg, torch, dev, scene, n_frames=60):
    """SURVEY config 1 as the harness runs it (tests/pipeline.py): one stereo sequence, tracking + mapping chained — both pyramids,
    trackNewestCoarse, traceOn of every key frame's immature points, and at every 5th frame the key-frame chain (candidate filter,
    activation, windowed optimisation, marginalisation, pixel selection, static stereo). Wall clock per frame through the C ABI
    with host images, next to the same harness on the oracle port (one host thread). The harness itself is Python (per-point
    dicts, numpy); its share is reported by timing the operator calls separately."""
    import synth, pipeline as PL
    shape = dict(w=synth.W, h=synth.H, K=synth.K4)
    poses = [synth.camera_pose(0.2 * k) for k in range(n_frames)]
    left = [synth.render_torch(scene, p, device=f"cuda:{dev}")[0] for p in poses]
    right = [synth.render_torch(scene, synth.right_of(p), device=f"cuda:{dev}")[0] for p in poses]

    def run(backend):
        P = PL.StereoPipeline(backend)
        per_frame, per_frame_ops = [], []
        for k, (l, r) in enumerate(zip(left, right)):
            t0 = time.perf_counter(); o0 = backend.t_ops; b0 = dict(backend.t_by_op)
            P.step(l, r)
            per_frame.append(time.perf_counter() - t0); per_frame_ops.append(backend.t_ops - o0)
            if os.environ.get("SDSO_BENCH_VERBOSE") and k % 5 == 0:
                d = {n: round(1e3 * (v - b0.get(n, 0.0)), 2) for n, v in backend.t_by_op.items() if v - b0.get(n, 0.0) > 2e-4}
                print(f"[sequence] frame {k}: {d}", file=sys.stderr)
        return P, np.array(per_frame), np.array(per_frame_ops)

    Bg = PL.Backend(shape, pkg)
    run(Bg)                      # warm-up (allocations, graph captures)
    Bg.close()
    Bg = PL.Backend(shape, pkg)
    l0 = Bg.api.launch_count()
    Pg, tg, og = run(Bg)
    launches = Bg.api.launch_count() - l0
    by_op = {k: round(1e3 * v, 3) for k, v in sorted(Bg.t_by_op.items(), key=lambda kv: -kv[1])}
    Bg.close()
    Po, to, oo = run(PL.Backend(shape))
    kf = np.arange(n_frames) % 5 == 0
    dt = np.array([np.abs(Pg.traj[k][:3, 3] - Po.traj[k][:3, 3]).max() for k in range(n_frames)])
    return dict(
        workload=f"one synthetic KITTI-shape stereo sequence, {n_frames} frames, 1232x368, key frame every 5th frame, window of 7, 1500 immature / 2000 active points: "
                 "makeImages (left + right) + trackNewestCoarse + traceOn per frame; selector, static stereo, distance map, candidate loop, activation, "
                 "windowed optimisation (6 iterations), marginalisation per key frame (tests/pipeline.py)",
        metric="tracked stereo frames/s (one sequence, end to end through the C ABI, host images in, poses out)",
        note="frame 0 (initialisation: first allocations of the context, stereo initialisation of the map) is reported separately and left out of the rate on both arms; "
             "value = frames / time spent INSIDE the operator calls (ABI entry points with host buffers: uploads, launches, read-backs, ctypes marshalling) — what a "
             "C++ caller pays; the Python harness around them (per-point dicts, numpy bookkeeping that stands in for FullSystem) is reported separately as wall time",
        value=(n_frames - 1) / og[1:].sum(), unit="frames/s", ms_per_tracked_frame=1e3 * float(np.median(og[~kf])), ms_per_key_frame=1e3 * float(np.median(og[kf][1:])),
        first_frame_ms=1e3 * float(og[0]), operator_ms_per_frame=[round(1e3 * float(x), 3) for x in og], operator_ms_total_by_name=by_op,
        wall_including_python_harness=dict(frames_per_s=(n_frames - 1) / tg[1:].sum(), ms_per_tracked_frame=1e3 * float(np.median(tg[~kf])), ms_per_key_frame=1e3 * float(np.median(tg[kf][1:]))),
        kernel_launches=int(launches),
        e2e=dict(value=(n_frames - 1) / og[1:].sum(), unit="frames/s", h2d_bytes_per_step=int(2 * synth.W * synth.H * 4), d2h_bytes_per_step=int(12 * 8 + 16 + 40),
                 note="host float images are uploaded inside sdso_make_images; immature records cross once per frame (one launch for all hosts)"),
        cpu_baseline=dict(value=(n_frames - 1) / oo[1:].sum(), unit="frames/s", cores=1, kind="port", ms_per_tracked_frame=1e3 * float(np.median(oo[~kf])),
                          ms_per_key_frame=1e3 * float(np.median(oo[kf][1:])), wall_frames_per_s=(n_frames - 1) / to[1:].sum(), first_frame_ms=1e3 * float(oo[0]),
                          sample=f"the same {n_frames} frames through the same harness on the oracle port, one host thread, operator time"),
        parity=dict(max_translation_difference_m=float(dt.max()), note="free-running chains: see tests/test_pipeline.py for the envelope (the oracle's own spread)"))


def leg_sharded_ba(pkg, torch, dist, dev, rank, world, scene, iters, peak):
    """Config 4 LM iteration with the points sharded over `world` ranks and ONE allreduce of the damped reduced system, next to
    the unsharded iteration on the same GPU; increments compared with the single-GPU solve on every rank."""
    import synth
    c, win = make_ba_case(scene, "config4")
    ctx = pkg.Context(c["w"], c["h"], c["K"], synth.BASELINE, device=dev)
    stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [pkg.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.nccl_init(rank, world, uid[0])
        # the hand-written exchange over NVLink peer memory (csrc/collective.cu): every rank exports its block, all map all
        # (if any rank cannot export / map a block — no peer access, IPC not permitted — every rank stays on NCCL)
        dsys = 4 + 8 * c["n"]
        peer_ok, peer_why = 1, ""
        try:
            handle = ctx.peer_alloc(world, dsys * dsys + dsys + 1)
        except Exception as ex:   # noqa: BLE001
            handle, peer_ok, peer_why = b"\0" * 64, 0, str(ex)
        handles = [None] * world
        dist.all_gather_object(handles, (handle, peer_ok))
        if all(h[1] for h in handles):
            try:
                ctx.peer_connect(rank, world, [h[0] for h in handles])
            except Exception as ex:   # noqa: BLE001
                peer_ok, peer_why = 0, str(ex)
        else:
            peer_ok = 0
        flag = torch.tensor([peer_ok], device=f"cuda:{dev}", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        have_peer = bool(int(flag[0]))
        if not have_peer and peer_ok:
            ctx.peer_select(0)
        dist.barrier()
    fids = []
    for f in win["frames"]:
        fid = ctx.frame_create(); ctx.make_images(fid, f["image"]); fids.append(fid)
    P = len(win["points"])
    b, e = pkg.shard_range(P, rank, world)

    if world <= 1:
        have_peer, peer_why = False, ""

    def timed(W, sharded, allreduce=True):
        def it():
            W.linearize_all_async(True); W.assemble()
            if sharded and world > 1 and allreduce:
                W.allreduce()
            W.solve_assembled(2, want=False)
        for _ in range(5):
            it()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = _events(torch)
        e0.record(stream)
        for _ in range(iters):
            it()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{dev}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        return ms

    Ws, _, _ = device_window(pkg, ctx, win, fids, b, e, rank, world)
    Ws.linearize_all_async(True); Ws.assemble()
    E = Ws.allreduce(want_energy=True) if world > 1 else None
    xs, Hs, bs = Ws.solve_assembled(2)
    steps_s = Ws.get_points()["step"].copy()
    ms_sharded = timed(Ws, True)
    ms_no_ar = timed(Ws, True, allreduce=False) if world > 1 else ms_sharded

    def exchange_alone():
        for _ in range(5):
            Ws.allreduce()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = _events(torch)
        e0.record(stream)
        for _ in range(50):
            Ws.allreduce()
        e1.record(stream); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50 * 1e3], device=f"cuda:{dev}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    ar_us = ar_nccl_us = ms_sharded_nccl = None
    timed_out = 0
    if world > 1:
        ar_us = exchange_alone()
        if have_peer:
            timed_out = ctx.peer_status()
            dist.barrier()
            ctx.peer_select(0)                   # the same iteration with NCCL's allreduce in place of the peer-memory kernel
            ms_sharded_nccl = timed(Ws, True)
            ar_nccl_us = exchange_alone()
            ctx.peer_select(1)
        else:
            ms_sharded_nccl, ar_nccl_us = ms_sharded, ar_us
    Wf, _, _ = device_window(pkg, ctx, win, fids)
    Ef = Wf.linearize_all(True)
    xf, Hf, bf = Wf.solve(2)
    steps_f = Wf.get_points()["step"]
    ms_full = timed(Wf, False)
    R = Wf.counts()["res"]
    relx = float(np.abs(xs - xf).max() / np.abs(xf).max())
    rels = float(np.abs(steps_s - steps_f[b:e]).max() / np.abs(steps_f).max())
    ok = torch.tensor([relx, rels], device=f"cuda:{dev}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MAX)
    ctx.close()
    evals = 8 * R
    return dict(
        workload=f"config 4 LM iteration ({c['n']} KF, {P} points, {R} residuals, {c['w']}x{c['h']}), points sharded over {world} GPU(s), one exchange of the "
                 f"(4+8n)^2+(4+8n)+1 = {(4 + 8 * c['n']) ** 2 + 4 + 8 * c['n'] + 1} doubles per iteration",
        n_gpus=world, ms_per_lm_iteration_sharded=ms_sharded, ms_per_lm_iteration_1gpu=ms_full, speedup_vs_1gpu=ms_full / ms_sharded,
        ms_per_lm_iteration_sharded_without_allreduce=ms_no_ar, allreduce_us=ar_us,
        exchange=("one kernel per rank over NVLink peer memory (CUDA IPC): flag-carrying words pushed into every peer's block, polled locally, rank-ordered sum; "
                  "no fence / atomic / barrier (csrc/collective.cu)") if have_peer else f"NCCL allreduce (peer-memory set-up not available: {peer_why})",
        exchange_timed_out=timed_out, with_nccl_allreduce=dict(ms_per_lm_iteration_sharded=ms_sharded_nccl, allreduce_us=ar_nccl_us,
                                                                speedup_vs_1gpu=(ms_full / ms_sharded_nccl) if ms_sharded_nccl else None),
        value=evals / (ms_sharded * 1e-3), unit="evals/s",
        roofline=dict(bound="hbm", achieved=evals * BA_BYTES_PER_EVAL / (ms_sharded * 1e-3) / 1e9, peak=peak * world, unit="GB/s",
                      frac=evals * BA_BYTES_PER_EVAL / (ms_sharded * 1e-3) / 1e9 / (peak * world)),
        parity=dict(rel_dx_vs_1gpu=float(ok[0]), rel_dstep_vs_1gpu=float(ok[1]), energy_sharded=E, energy_1gpu=Ef))

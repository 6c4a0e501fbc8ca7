"""Point-sharded windowed BA over N GPUs, run under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_ba.py [--frames 7 --points 2000]
Checks the sharded solve against the single-GPU solve of the same window (rank 0) and times one LM iteration
(linearizeAll + accumulate + stitch + allreduce + solve + resubstitute) sharded vs unsharded with CUDA events (max over ranks)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import conftest, synth, ba_synth


def build(pkg, ctx, win, fids, rank, world, shard):
    P = len(win["points"])
    b, e = pkg.shard_range(P, rank, world) if shard else (0, P)
    pts = win["points"][b:e]
    W = pkg.Window(ctx)
    for k, f in enumerate(win["frames"]):
        idx = W.add_frame(fids[k], f["T_w2c"], f["a"], f["b"], f["frameID"])
        W.set_state(idx, f["state"]); W.set_energy_th(idx, f["energyTH"])
    col = np.zeros((len(pts), 8), np.float32); wts = np.zeros((len(pts), 8), np.float32)
    for h in range(win["n"]):  # D1 on the device, per host frame
        sel = [i for i, p in enumerate(pts) if p["host"] == h]
        if not sel: continue
        rec, ok = ctx.immature_init(fids[h], np.array([[pts[i]["u"], pts[i]["v"]] for i in sel], np.float32))
        col[sel] = rec["color"]; wts[sel] = rec["weights"]
    W.set_points([p["host"] for p in pts], [p["u"] for p in pts], [p["v"] for p in pts], [p["idepth"] for p in pts],
                 [p["idepth_zero"] for p in pts], col, wts, [p["has_prior"] for p in pts])
    rp, rt = [], []
    for pi, p in enumerate(pts):
        for t in p["targets"]:
            rp.append(pi); rt.append(t)
    W.set_residuals(rp, rt)
    W.set_shard(rank if shard else 0, world if shard else 1)
    W.prepare()
    return W, len(rp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", dest="n", type=int, default=7); ap.add_argument("--points", type=int, default=2002)
    ap.add_argument("--w", type=int, default=synth.W); ap.add_argument("--h", type=int, default=synth.H)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    pkg = conftest.load_pkg()
    K = synth.K4 if a.w == synth.W else (360.0 * a.w / 640, 360.0 * a.w / 640, a.w / 2 - 0.5, a.h / 2 - 0.5)
    ctx = pkg.Context(a.w, a.h, K, synth.BASELINE, device=lr)
    stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [pkg.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.nccl_init(rank, world, uid[0])
    scene = synth.make_scene()
    win = ba_synth.make_window(scene, n=a.n, P=a.points, seed=7, spacing=0.35, w=a.w, h=a.h, K=K)
    fids = []
    for f in win["frames"]:
        fid = ctx.frame_create(); ctx.make_images(fid, f["image"]); fids.append(fid)

    def lm_iteration(W, sharded):
        W.linearize_all_async(True)
        W.assemble()
        if sharded and world > 1: W.allreduce()
        W.solve_assembled(2, want=False)

    def timed(W, sharded):
        for _ in range(5): lm_iteration(W, sharded)
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.iters): lm_iteration(W, sharded)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{lr}", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t[0])
        return ms

    # sharded
    Ws, Rloc = build(pkg, ctx, win, fids, rank, world, True)
    Ws.linearize_all_async(True); Ws.assemble()
    E = Ws.allreduce(want_energy=True) if world > 1 else None
    xs, Hs, bs = Ws.solve_assembled(2)
    ms_sharded = timed(Ws, True)
    steps_s = Ws.get_points()["step"].copy()
    # unsharded reference on every rank (same device), compared on rank 0
    Wf, Rfull = build(pkg, ctx, win, fids, 0, 1, False)
    Ef = Wf.linearize_all(True)
    xf, Hf, bf = Wf.solve(2)
    ms_full = timed(Wf, False)
    steps_f = Wf.get_points()["step"]
    b, e = pkg.shard_range(len(win["points"]), rank, world)
    ok_steps = bool(np.allclose(steps_s, steps_f[b:e], rtol=2e-3, atol=2e-4 * np.abs(steps_f).max()))
    glob = np.abs(xf).max()
    relx = float(np.abs(xs - xf).max() / glob)
    relH = float(np.abs(Hs - Hf).max() / np.abs(Hf).max())
    oks = torch.tensor([int(ok_steps and relx < 3e-4)], device=f"cuda:{lr}")
    if world > 1: dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, n=a.n, points=len(win["points"]), residuals=Rfull, local_residuals=Rloc, size=[a.w, a.h],
                              rel_dx=relx, rel_dH=relH, energy_sharded=E, energy_full=Ef, parity_ok=bool(int(oks[0])),
                              ms_per_lm_iteration_sharded=ms_sharded, ms_per_lm_iteration_1gpu=ms_full,
                              evals_per_s_sharded=8 * Rfull / (ms_sharded * 1e-3), evals_per_s_1gpu=8 * Rfull / (ms_full * 1e-3))))
    ctx.close()
    if world > 1: dist.destroy_process_group()


if __name__ == "__main__":
    main()

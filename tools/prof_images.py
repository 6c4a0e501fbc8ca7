"""makeImages throughput: 64 frames of 1232x368 from resident 8-bit sources, batches of 32 per launch.
Usage: python tools/prof_images.py [--reps 20]   (also the ncu target for pyr_fused_kernel)"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G
ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=20); ap.add_argument("--frames", type=int, default=128)
a = ap.parse_args()
pkg = G._load_pkg()
w, h = 1232, 368
ctx = pkg.Context(w, h, (718.856, 718.856, 607.19, 185.2), 0.54)
rng = np.random.default_rng(0)
N = a.frames
src = torch.from_numpy(rng.integers(0, 255, (N, h, w), dtype=np.uint8)).cuda()
fids = [ctx.frame_create() for _ in range(N)]
ptrs = [src[i].data_ptr() for i in range(N)]
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
def run():
    ctx.make_images_batch_device(fids, ptrs, u8=True)
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
levels = ctx.levels
px = sum((w >> l) * (h >> l) for l in range(levels))
bytes_img = w * h + 16 * px
print(json.dumps({"frames": N, "ms_per_pass": ms, "us_per_32": ms / (N / 32) * 1e3, "GBps": bytes_img * N / ms / 1e6, "bytes_per_image": bytes_img}))

"""The sharded-BA leg of bench.py on its own (torchrun): config 4 LM iteration with the points sharded over the ranks, the
peer-memory exchange next to NCCL's allreduce. usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_sharded.py"""
import importlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
pkg = importlib.import_module("stereo-dso-g2o_b200")
import bench_legs as BL, synth
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6534.8
out = BL.leg_sharded_ba(pkg, torch, dist, local, rank, world, synth.make_scene(), 50, peak)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()

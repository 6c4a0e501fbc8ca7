"""N>1 path on the CPU (gloo, world_size 2): the host-side sharding logic of SURVEY.md 8e. Each rank owns a contiguous
block of the allPoints order (sdso_shard_range of the product library), builds the partial top / Schur systems of its
block with the oracle, the partials are summed with one all_reduce, and the sum must equal the unsharded system."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest, synth, ba_synth
    import oracle_py as O, oracle_ba_py as OB
    pkg = conftest.load_pkg()
    w, h, K = 640, 192, (360.0, 360.0, 319.5, 95.5)
    scene = synth.make_scene()
    win = ba_synth.make_window(scene, n=3, P=150, seed=9, spacing=0.6, w=w, h=h, K=K)
    P = len(win["points"])
    b, e = pkg.shard_range(P, rank, world)
    sub = ba_synth.shard_window(win, b, e)
    orc = O.Oracle(w, h, K, synth.BASELINE)
    ba, _, _ = ba_synth.fill_oracle(sub, orc, OB.OracleBA, OB.immature_init)
    E = ba.linearize_all(True)
    HA, bA, _ = ba.accumulate_top(0, False)
    HL, bL, _ = ba.accumulate_top(1, rank == 0)     # priors enter once
    Hsc, bsc = ba.accumulate_sc(True)
    d = HA.shape[0]
    lam = 1e-5
    HF = HA + HL
    HF[np.diag_indices(d)] *= (1 + lam)
    HF -= Hsc / (1 + lam)
    bF = bA + bL - bsc
    buf = torch.from_numpy(np.concatenate([HF.ravel(), bF, [E]]))
    dist.all_reduce(buf)                            # the ONE exchange step
    if rank == 0:
        orc2 = O.Oracle(w, h, K, synth.BASELINE)
        full, _, _ = ba_synth.fill_oracle(win, orc2, OB.OracleBA, OB.immature_init)
        Ef = full.linearize_all(True)
        x, Hf, bf = full.solve(0)
        np.savez(os.path.join(out_dir, "r.npz"), H=buf[:d * d].numpy().reshape(d, d), b=buf[d * d:d * d + d].numpy(), E=float(buf[-1]),
                 Hf=Hf, bf=bf, Ef=Ef, ranges=np.array([pkg.shard_range(P, r, world) for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_all_points():
    sys.path.insert(0, HERE)
    import conftest
    pkg = conftest.load_pkg()
    for P in (0, 1, 7, 2000, 20001):
        for world in (1, 2, 4, 8):
            r = [pkg.shard_range(P, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == P
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_partial_systems_sum_to_the_full_system(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = np.load(os.path.join(str(tmp_path), "r.npz"))
    scale = np.abs(r["Hf"]).max()
    # float accumulators are summed per shard and then in double across shards: agreement to float rounding of the blocks
    d = r["Hf"].shape[0]
    n = (d - 4) // 8
    edges = [0, 4] + [4 + 8 * (i + 1) for i in range(n)]
    for i in range(len(edges) - 1):
        for j in range(len(edges) - 1):
            a = r["H"][edges[i]:edges[i + 1], edges[j]:edges[j + 1]]
            b = r["Hf"][edges[i]:edges[i + 1], edges[j]:edges[j + 1]]
            s = max(np.abs(b).max(), 1e-7 * scale)
            assert np.abs(a - b).max() <= 1e-4 * s, (i, j)
    assert np.allclose(r["b"], r["bf"], rtol=1e-4, atol=1e-4 * np.abs(r["bf"]).max())
    assert np.isclose(r["E"], r["Ef"], rtol=1e-6)

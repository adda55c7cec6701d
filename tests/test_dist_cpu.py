"""CPU, gloo, world_size 2: the host-side logic of the multi-GPU path (g2pc/dist.py) — accumulator merge rule, global
points-per-Gaussian, global histogram, shard partitions — against the single-process result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "3dgs-to-pc_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from g2pc import dist as gd
    try:
        g = torch.Generator().manual_seed(7)
        n, ncam = 5000, 9
        # synthetic per-camera per-Gaussian contributions and colours (same on all ranks)
        contrib = torch.rand(ncam, n, generator=g)
        contrib[contrib < 0.6] = 0.0
        contrib[3] = contrib[1]  # exact ties between cameras 1 and 3: the earlier camera must win
        cols = torch.rand(ncam, n, 3, generator=g)
        # single-process reference: strict > in camera order
        ref_max = torch.zeros(n)
        ref_col = torch.zeros(n, 3)
        for c in range(ncam):
            upd = contrib[c] > ref_max
            ref_max[upd] = contrib[c][upd]
            ref_col[upd] = cols[c][upd]
        # sharded: this rank's cameras
        my_max = torch.zeros(n)
        my_col = torch.zeros(n, 3)
        first = torch.full((n,), torch.iinfo(torch.int32).max, dtype=torch.int32)
        for c in gd.camera_shard(ncam):
            upd = contrib[c] > my_max
            my_max[upd] = contrib[c][upd]
            my_col[upd] = cols[c][upd]
            first[upd] = c
        gd.merge_colour_accumulators(my_max, my_col, first)
        ok_merge = bool(torch.equal(my_max, ref_max) and torch.equal(my_col, ref_col))

        # global points-per-Gaussian on index shards == single-process distribute_points
        mags = torch.rand(n, generator=g, dtype=torch.float64) ** 4
        from oracle import sampling as osamp
        ref_ppg = osamp.distribute_points(mags.clone(), 12000)
        b, e = gd.gaussian_shard(n)
        loc = gd.global_points_per_gaussian(mags[b:e].clone(), 12000)
        ok_ppg = bool(torch.equal(loc, ref_ppg[b:e]))
        hist = gd.global_histogram(loc.to(torch.int32))
        ok_hist = bool(torch.equal(hist, torch.bincount(ref_ppg.to(torch.int64))))
        # partitions cover everything exactly once
        shards = [gd.gaussian_shard(n, r, world) for r in range(world)]
        ok_part = shards[0][0] == 0 and shards[-1][1] == n and all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))
        cams = sorted(sum([gd.camera_shard(ncam, r, world) for r in range(world)], []))
        ok_part = ok_part and cams == list(range(ncam))
        q.put((rank, ok_merge, ok_ppg, ok_hist, ok_part))
    finally:
        dist.destroy_process_group()


def test_dist_host_logic_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for (rank, ok_merge, ok_ppg, ok_hist, ok_part) in res:
        assert ok_merge, f"rank {rank}: merged accumulators differ from the single-process update rule"
        assert ok_ppg, f"rank {rank}: sharded points-per-Gaussian differ"
        assert ok_hist, f"rank {rank}: global histogram differs"
        assert ok_part, f"rank {rank}: shard partition"


def test_local_bin_counts():
    from g2pc import dist as gd, sampler
    hist = np.array([5, 10, 0, 7, 3, 1, 0, 2])
    bins = sampler.plan_bins(hist, True)
    local = np.array([1, 4, 0, 0, 3, 0, 0, 1])
    lb = gd.local_bin_counts(bins, local)
    assert [b[:3] for b in lb] == [b[:3] for b in bins]
    assert sum(b[3] for b in lb) == int(local[1:].sum())  # value 0 -> n = 0 bin is dropped

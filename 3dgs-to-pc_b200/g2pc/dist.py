"""Multi-GPU execution of the hot path: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The reference has no distributed code.  How the path shards:

  colour stage   front-to-back blending at a pixel depends on ALL nearer Gaussians, so index shards cannot be blended
                 independently.  Cameras are sharded instead: every rank holds the whole Gaussian array (3 M x ~250 B
                 < 1 GB) and renders cameras rank, rank+W, ...  The per-Gaussian accumulators are then merged with the
                 reference's own update rule (strict >, earlier camera wins ties):
                     all_reduce(MAX)  on the max contribution
                     all_reduce(MIN)  on the index of the first camera that reached it   (tie-break = reference order)
                     all_reduce(SUM)  on the colour, zeroed everywhere except on the winning rank
                 ~3 M x 20 B per run: well under a millisecond over NVLink 5, no custom transport needed.
  sampling       independent per Gaussian: contiguous index ranges, Philox keyed by the GLOBAL Gaussian id, so the
                 emitted points do not depend on the number of ranks.  Exchange: one f64 all_reduce(SUM) of the
                 magnitude sum, one all_reduce(SUM) of the points-per-Gaussian histogram (bins are global), and — only
                 if the caller wants the full cloud on every rank — an all_gather of the per-rank point counts.

Everything here is host-side plumbing over torch.distributed; the math stays in the kernels.  The functions take
a `group`-less default process group and work with the gloo backend on CPU tensors too (tests/test_dist_cpu.py
exercises the merge and partition logic with world_size 2).
"""
import math

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def camera_shard(num_cameras, rank=None, world_size=None):
    """Indices of the cameras rank renders (round-robin keeps the per-rank cost even along a camera trajectory)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, num_cameras, world_size))


def gaussian_shard(n, rank=None, world_size=None):
    """[begin, end) of the contiguous Gaussian index range owned by `rank`."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    per = (n + world_size - 1) // world_size
    return min(n, rank * per), min(n, (rank + 1) * per)


def merge_colour_accumulators(max_contrib, colours, first_camera):
    """In-place merge of the per-Gaussian colour accumulators over all ranks.

    max_contrib (N,) f32: best contribution seen by this rank's cameras; colours (N,3) f32: blended colour of that
    pixel; first_camera (N,) int32: global index of the camera that FIRST produced this rank's maximum (large sentinel
    if none).  After the call all ranks hold the values a single process rendering the cameras in index order would
    hold (gauss_render.py:387-395 update rule: strict >, so the earliest camera wins a tie)."""
    if world()[1] == 1:
        return
    mine = max_contrib.clone()
    dist.all_reduce(max_contrib, op=dist.ReduceOp.MAX)
    # among the ranks that hold the global maximum, the one whose camera came first in the reference's loop order wins
    cand = torch.where(mine == max_contrib, first_camera, torch.full_like(first_camera, torch.iinfo(torch.int32).max))
    best_cam = cand.clone()
    dist.all_reduce(best_cam, op=dist.ReduceOp.MIN)
    winner = (cand == best_cam) & (mine == max_contrib) & (max_contrib > 0)
    colours.mul_(winner.unsqueeze(1).to(colours.dtype))
    dist.all_reduce(colours, op=dist.ReduceOp.SUM)
    first_camera.copy_(best_cam)


def global_points_per_gaussian(local_magnitudes, num_points):
    """distribute_points (gauss_to_pc.py:73-90) when the magnitudes are sharded: the ratio uses the global sum; the
    zero -> one fix-up walks the zeros in global index order (rank-major)."""
    r, w = world()
    total = local_magnitudes.sum().to(torch.float64).reshape(1)
    if w > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    ppg = torch.round(local_magnitudes * (num_points / total))
    is_zero = ppg == 0
    stats = torch.stack([ppg.sum().to(torch.float64), is_zero.sum().to(torch.float64)]).reshape(1, 2)
    if w > 1:
        allstats = [torch.zeros_like(stats) for _ in range(w)]
        dist.all_gather(allstats, stats)
        allstats = torch.cat(allstats, 0)
    else:
        allstats = stats
    allstats = allstats.cpu()
    deficit = num_points - float(allstats[:, 0].sum())
    zeros_before = float(allstats[:r, 1].sum())
    zeros_total = float(allstats[:, 1].sum())
    take = int(min(deficit, zeros_total))
    if take < 0:
        take = int(zeros_total) + take
    # this rank promotes the zeros whose global rank (1-based) is <= take
    local_quota = int(max(0, min(take - zeros_before, float(allstats[r, 1]))))
    rank_among_zeros = torch.cumsum(is_zero.to(torch.int64), 0)
    ppg[is_zero & (rank_among_zeros <= local_quota)] = 1
    return ppg


def global_histogram(local_ppg_int, device=None):
    """bincount of the points-per-Gaussian over all ranks (bins are planned on the global histogram)."""
    r, w = world()
    mx = local_ppg_int.max().reshape(1).to(torch.int64) if local_ppg_int.numel() else torch.zeros(1, dtype=torch.int64, device=local_ppg_int.device)
    if w > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    size = int(mx.item()) + 1
    hist = torch.bincount(local_ppg_int.to(torch.int64), minlength=size)
    if w > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def local_bin_counts(bins, local_hist):
    """Members of every global bin that live on this rank (same (start, end, n) bins, local counts; bins that are
    empty locally keep count 0 and are skipped by the planner)."""
    csum = np.concatenate([[0], np.cumsum(local_hist)])
    out = []
    for (start, end, n, _) in bins:
        lo = min(max(int(math.ceil(start)), 0), local_hist.shape[0])
        hi = min(max(int(math.ceil(end)), 0), local_hist.shape[0])
        out.append((start, end, n, int(csum[hi] - csum[lo]) if hi > lo else 0))
    return out


LAST_PHASES = {}


class _Phases:
    """Per-rank wall-clock phases of one sharded step (device-synchronised at every boundary, so only used in a separate
    profiling step: G2PC_PHASE_TIMING=1 or bench.py's timeline step).  There is no nsys in this image; this is the
    per-rank timeline that names what is left outside the kernels."""

    def __init__(self, enabled, device):
        import time
        self.enabled, self.device, self.t, self.out, self.time = enabled, device, None, {}, time
        if enabled:
            torch.cuda.synchronize(device)
            self.t = time.perf_counter()

    def mark(self, name):
        if not self.enabled:
            return
        torch.cuda.synchronize(self.device)
        now = self.time.perf_counter()
        self.out[name] = self.out.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


def upload_sharded(host_scene, device):
    """e2e path of an N-rank run: every rank copies only ITS row range of each (pinned) host array over PCIe and the ranks
    exchange the shards with one in-place all_gather per array over NVLink — instead of N full host->device uploads
    (1.44 GB each at C3).  Returns the full device tensors on every rank."""
    rank, W = world()
    out = {}
    for k, v in host_scene.items():
        n = v.shape[0]
        if W == 1:
            out[k] = v.to(device, non_blocking=True)
            continue
        per = (n + W - 1) // W
        full = torch.empty((per * W,) + tuple(v.shape[1:]), dtype=v.dtype, device=device)
        b, e = min(n, rank * per), min(n, (rank + 1) * per)
        if e > b:
            full[b:e].copy_(v[b:e], non_blocking=True)
        dist.all_gather_into_tensor(full, full[rank * per:(rank + 1) * per])  # in place: shard r sits at r * per
        out[k] = full[:n]
    return out


def convert_gaussians_to_pc_sharded(scene, transforms, intrinsics, settings, render_shs=False, gather=False,
                                    phase_timing=None):
    """The device-resident pipeline of gauss_to_pc.convert_gaussians_to_pc on W ranks.

    scene: dict of device tensors (xyz, scales, rots, colours, opacities, shs) holding the WHOLE scene on every rank.
    Returns PointCloudData with this rank's slice of the point cloud (Gaussian index shard)."""
    import gauss_to_pc as g2p
    from gauss_handler import Gaussians
    from gauss_render import get_renderer
    from camera_handler import get_camera
    from . import config, sampler

    rank, W = world()
    s = settings
    if s.cull_large_percentage > 0.0 or s.generate_mesh:
        raise NotImplementedError("the sharded pipeline covers the visibility / opacity / bounding-box / surface-distance "
                                  "culls; size-percentile culls and meshing run through convert_gaussians_to_pc")
    import os
    global LAST_PHASES
    ph = _Phases(bool(int(os.environ.get("G2PC_PHASE_TIMING", "0"))) if phase_timing is None else phase_timing,
                 scene["xyz"].device)
    n_all = scene["xyz"].shape[0]
    gaussians = Gaussians(scene["xyz"], scene["scales"], scene["rots"], scene["colours"], scene["opacities"],
                          shs=scene.get("shs"))
    if s.calculate_normals:
        gaussians.calculate_normals()
    ph.mark("covariances+normals")

    contributions = None
    max_contrib, surface_mask = None, None
    if s.render_colours:
        renderer = get_renderer(s.renderer_type, gaussians.xyz, torch.unsqueeze(torch.clone(gaussians.opacities), 1),
                                gaussians.colours, gaussians.covariances, shs=gaussians.shs if render_shs else None,
                                visible_gaussian_threshold=s.visibility_threshold,
                                surface_distance_std=s.surface_distance_std,
                                calculate_surface_distance=s.surface_distance_std is not None)
        names = list(transforms.keys())
        # the accumulate kernel records the index of the camera that raised each maximum (needed by the merge)
        first_cam = torch.full((n_all,), torch.iinfo(torch.int32).max, dtype=torch.int32, device=scene["xyz"].device)
        renderer.first_frame = first_cam
        renderer.async_mode = True
        for ci in camera_shard(len(names)):
            name = names[ci]
            tr = transforms[name]
            tr = torch.as_tensor(tr, dtype=torch.float32) if not torch.is_tensor(tr) else tr
            cam = get_camera(s.renderer_type, tr, intrinsics[name], colour_resolution=s.colour_resolution,
                             sh_degree=s.max_sh_degree, white_bkgd=True, mask=None)
            renderer(cam, camera_index=ci)
        renderer.flush()
        ph.mark("colour stage (this rank's cameras)")
        merge_colour_accumulators(renderer.gaussian_max_contribution, renderer.gaussian_colours, first_cam)
        if s.renderer_type == "cuda" and W > 1:
            # CUDA back-end extras (gaussian_pointcloud_rasterization/__init__.py:152-158): the total contribution is a
            # SUM over cameras (float addition is re-associated across ranks: last-bit differences vs one process), the
            # surface distance a MIN
            dist.all_reduce(renderer.gaussian_total_contribution, op=dist.ReduceOp.SUM)
            if s.surface_distance_std is not None:
                dist.all_reduce(renderer.gaussian_min_surface_distance, op=dist.ReduceOp.MIN)
        gaussians.colours = renderer.get_gaussian_colours()
        if s.surface_distance_std is not None:
            surface_mask = renderer.get_gaussians_with_low_surface_distance()
        if s.remove_unrendered_gaussians:
            max_contrib = renderer.gaussian_max_contribution
        if s.prioritise_visible_gaussians:
            contributions = renderer.get_total_gaussian_contributions()
        g2p.LAST_RENDER_STATS = {"stats": renderer._stats, "replays": renderer.replays}
        vis_thr = renderer.visible_gaussian_threshold
        del renderer
        ph.mark("accumulator merge (NCCL)")
    else:
        gaussians.colours = gaussians.colours * 255
        vis_thr = 0.0

    # ---- sampling: this rank owns the Gaussians [b, e): every cull + the index shard in ONE fused mask / compaction ------
    b, e = gaussian_shard(n_all)
    gid = gaussians.fused_cull(max_contribution=max_contrib, visibility_threshold=vis_thr, min_opacity=s.min_opacity,
                               bounding_box_min=s.bounding_box_min, bounding_box_max=s.bounding_box_max,
                               extra_mask=surface_mask, index_range=(b, e))
    if contributions is not None:
        contributions = contributions[gid]
    ph.mark("cull + compaction")
    valid = gaussians.validate_covariances()
    gid = gid[valid]
    if contributions is not None:
        contributions = contributions[valid]

    # the same magnitude kernel as the single-process pipeline (bit-identical magnitudes), then the global budget
    _, mags = gaussians.points_per_gaussian(s.num_points, contributions)
    ppg = global_points_per_gaussian(mags, s.num_points).to(torch.int32)
    hist = global_histogram(ppg).cpu().numpy()
    bins = sampler.plan_bins(hist, s.exact_num_points)
    attempts = 5 if not s.exact_num_points else 100
    pts, cols, nrm, total, status, dbg = g2p.sample_points_per_gaussian(
        gaussians.xyz, gaussians.covariances, gaussians.colours, gaussians.normals if s.calculate_normals else None, ppg,
        s.mahalanobis_distance_std, s.exact_num_points, attempts, config.SEED, 0, gids=gid, global_bins=bins)
    t = int(total.item())
    g2p._check_status(status)
    ph.mark("validate + point budget + sampling")
    LAST_PHASES = ph.out
    return g2p.PointCloudData(points=pts[:t], colours=cols[:t], normals=(nrm[:t] if nrm is not None else None))

"""Host side of S2 (sampling + Mahalanobis cull): bin planning and kernel driver.

The reference does its bin planning on the host too (gauss_to_pc.py:308-343: torch.unique / bincount / numpy
gradient heuristic), then loops over bins and attempts from Python.  Here the host only turns the
points-per-Gaussian histogram into two small tables (tiles, units); all per-Gaussian work happens in two kernels
(g2pc_sample_count, g2pc_sample_emit; csrc/s2_sample.cu).
"""
import math
import threading

import numpy as np
import torch

from . import capi, config

_call_counter = 0
_call_lock = threading.Lock()


def next_call_id():
    """Each sampling call gets its own RNG sub-stream (4th Philox counter word)."""
    global _call_counter
    with _call_lock:
        c = _call_counter
        _call_counter += 1
    return c


def reset_call_counter(value=0):
    global _call_counter
    with _call_lock:
        _call_counter = value


# ------------------------------------------------------------------------------------------------------------
def calculate_bin_sizes_from_hist(hist_nonzero):
    """Bin-width heuristic of the reference (gauss_to_pc.py:105-138) evaluated on the host from the histogram of
    points-per-Gaussian (counts of the occurring values, ascending).  Returns (start_bin, bin_size).
    Raises ValueError for fewer than two distinct values, as numpy.gradient does in the reference."""
    dist = np.asarray(hist_nonzero)
    second = np.absolute(np.gradient(np.gradient(dist)))
    bin_size = max(len(dist) // 100, 1)
    usable = len(second) - len(second) % bin_size
    per_bin = second[:usable].reshape(-1, bin_size).sum(axis=1)
    cut_off = np.max(per_bin) // 50
    peak = int(np.argmax(per_bin))
    quiet = np.nonzero(per_bin[peak:] < cut_off)[0]
    start_bin = int(quiet[0]) if quiet.shape[0] != 0 else 1
    return start_bin, bin_size


def plan_bins(hist, exact_num_points):
    """hist[v] = number of Gaussians assigned v points.  Returns the reference's bins in loop order as a list of
    (start, end, n, count) where Gaussians with start <= ppg < end all receive n points
    (n = floor(start + (end - start) / 2), gauss_to_pc.py:326-337).  Bins with n <= 0 or count == 0 are dropped
    (:339-343)."""
    hist = np.asarray(hist)
    values = np.nonzero(hist)[0]
    if values.size == 0:
        return []
    if exact_num_points:
        edges = values.astype(np.float64)
    else:
        start_bin, bin_size = calculate_bin_sizes_from_hist(hist[values])
        head = values[:start_bin].astype(np.float32)
        tail = np.unique(np.ceil(values[start_bin:].astype(np.float32) / np.float32(bin_size))) * np.float32(bin_size)
        edges = np.concatenate([head, tail.astype(np.float32)]).astype(np.float64)
    csum = np.concatenate([[0], np.cumsum(hist)])
    bins = []
    for i in range(edges.shape[0]):
        start = float(edges[i])
        end = float(edges[i + 1]) if i != edges.shape[0] - 1 else start + 1
        n = math.floor(start + (end - start) / 2)
        if n <= 0:
            continue
        lo = min(max(int(math.ceil(start)), 0), hist.shape[0])
        hi = min(max(int(math.ceil(end)), 0), hist.shape[0])
        count = int(csum[hi] - csum[lo]) if hi > lo else 0
        if count < 1:
            continue
        bins.append((start, end, n, count))
    return bins


def _lpg_for(k):
    """Threads cooperating on one Gaussian: each draws at most ~8 samples per attempt."""
    if k <= 8:
        return 1
    return min(256, 1 << int(math.ceil(math.log2(k / 8.0))))


class SamplePlan:
    """Tile and unit tables for one sampling call (see include/g2pc.h: g2pc_tile_t, g2pc_unit_t)."""

    def __init__(self, bin_k_count, attempts_stored, include_centres=True):
        """bin_k_count: list of (k, count) in output order; Gaussians of bin b occupy bin-order indices
        [J0_b, J0_b + count_b)."""
        if attempts_stored > 255:
            raise capi.G2pcError("attempts_stored must be <= 255 (8-bit attempt tag of the emit pass)")
        if any(k >= (1 << 24) for (k, _) in bin_k_count):
            raise capi.G2pcError("more than 2^24 - 1 samples per Gaussian per attempt (24-bit sample tag of the emit pass)")
        # vectorised over bins (exact_num_points makes one bin per distinct count: thousands of bins x attempts)
        A = attempts_stored
        ks = np.asarray([k for (k, _) in bin_k_count], dtype=np.int64)
        counts = np.asarray([c for (_, c) in bin_k_count], dtype=np.int64)
        nb = ks.shape[0]
        lpgs = np.asarray([_lpg_for(int(k)) for k in ks], dtype=np.int64) if nb else np.zeros(0, np.int64)
        per_tile = 256 // np.maximum(lpgs, 1)
        nts = (counts + per_tile - 1) // np.maximum(per_tile, 1)
        j0s = np.concatenate([[0], np.cumsum(counts)])[:-1] if nb else np.zeros(0, np.int64)
        tile_base = np.concatenate([[0], np.cumsum(nts)])[:-1] if nb else np.zeros(0, np.int64)
        nt_total = int(nts.sum()) if nb else 0
        # tiles: bin of every tile, position inside the bin
        tbin = np.repeat(np.arange(nb), nts)
        tpos = np.arange(nt_total) - np.repeat(tile_base, nts)
        tstart = j0s[tbin] + tpos * per_tile[tbin]
        tcount = np.minimum(per_tile[tbin], j0s[tbin] + counts[tbin] - tstart)
        tiles_arr = np.stack([tstart, tcount, ks[tbin], lpgs[tbin]], axis=1) if nt_total else np.zeros((0, 4), np.int64)
        # units in output order: per bin [centre unit][attempt 0: tiles..][attempt 1: tiles..]...
        has_k = ks > 0
        per_bin_units = (1 if include_centres else 0) + np.where(has_k, A * nts, 0)
        ubase = np.concatenate([[0], np.cumsum(per_bin_units)])[:-1] if nb else np.zeros(0, np.int64)
        nu_total = int(per_bin_units.sum()) if nb else 0
        units_arr = np.zeros((nu_total, 4), dtype=np.int64)
        src_arr = np.zeros((nu_total,), dtype=np.int64)
        if include_centres and nb:
            units_arr[ubase] = np.stack([np.full(nb, -1), j0s, counts, np.zeros(nb, np.int64)], axis=1)
            src_arr[ubase] = nt_total * A + np.arange(nb)  # lengths gathered from concat([tile_totals, centre_lens])
        if nt_total:
            tk = has_k[tbin]
            t_idx = np.nonzero(tk)[0]            # tiles of bins with samples
            tb = tbin[t_idx]
            a = np.arange(A)
            # unit index of (tile t, attempt a) = ubase[bin] + centre + a * nts[bin] + tpos
            uidx = (ubase[tb] + (1 if include_centres else 0) + tpos[t_idx])[:, None] + a[None, :] * nts[tb][:, None]
            units_arr[uidx.reshape(-1)] = np.stack([np.broadcast_to(a[None, :], uidx.shape).reshape(-1),
                                                    np.repeat(tstart[t_idx], A), np.repeat(tcount[t_idx], A),
                                                    np.repeat(ks[tb], A)], axis=1)
            src_arr[uidx.reshape(-1)] = (t_idx[:, None] * A + a[None, :]).reshape(-1)
        j0 = int(counts.sum()) if nb else 0
        centre_lens = counts.tolist() if include_centres else []
        self.n = j0
        self.attempts_stored = A
        self.tiles = tiles_arr.astype(np.int32)
        self.units = units_arr.astype(np.int32)
        # lengths are gathered from concat([tile_totals (num_tiles*A), centre_lens])
        self.unit_src = src_arr.astype(np.int64)
        self.centre_lens = np.asarray(centre_lens, dtype=np.int64)
        self.capacity = int(sum(c * (k + (1 if include_centres else 0)) for (k, c) in bin_k_count))


def run_plan(plan, xyz, cov, colours, normals, perm, num_attempts, std, seed, call_id, gid_offset=0,
             out_dtype=None, cull_mode=None, want_normals=True, gids=None):
    """Launch the two S2 kernels for `plan`.  All tensors on one CUDA device.  Returns
    (points, colours, normals|None, total_tensor, status_tensor) with outputs sized plan.capacity (valid rows:
    [0, total))."""
    lib = capi.load()
    capi.require_cuda(xyz, cov, colours, normals, perm)
    dev = xyz.device
    out_dtype = out_dtype or config.OUTPUT_DTYPE
    cull_mode = config.CULL_MODE if cull_mode is None else cull_mode
    n = plan.n
    A = plan.attempts_stored
    nt = plan.tiles.shape[0]
    nu = plan.units.shape[0]
    st = capi.stream_ptr(dev)

    assert xyz.dtype == torch.float32 and cov.dtype == torch.float32
    xyz = xyz.contiguous()
    cov = cov.contiguous()
    colours = colours.contiguous()
    if normals is not None:
        normals = normals.contiguous().to(torch.float32)
    perm = perm.to(torch.int32).contiguous()
    assert perm.shape[0] == n
    if gids is not None:
        gids = gids.to(torch.int32).contiguous()  # uint32 bit pattern
        assert gids.shape[0] == xyz.shape[0]

    tiles_d = torch.from_numpy(plan.tiles).to(dev, non_blocking=True)
    units_d = torch.from_numpy(plan.units).to(dev, non_blocking=True)
    src_d = torch.from_numpy(plan.unit_src).to(dev, non_blocking=True)
    centre_d = torch.from_numpy(plan.centre_lens).to(dev, non_blocking=True)

    records = torch.empty((max(n, 1), 16), dtype=torch.float32, device=dev)
    xl = torch.empty((A, max(n, 1)), dtype=torch.int32, device=dev)
    tile_totals = torch.zeros((max(nt, 1) * A,), dtype=torch.int32, device=dev)
    status = torch.zeros((capi.ST_WORDS,), dtype=torch.int32, device=dev)

    capi.call("g2pc_sample_count", capi.ptr(xyz), capi.ptr(cov), capi.ptr(colours), capi.dtype_code(colours), capi.ptr(normals),
        capi.ptr(perm), capi.ptr(gids), int(gid_offset), n, capi.ptr(tiles_d), nt, int(num_attempts), A, float(std),
        int(cull_mode), int(seed) & 0xFFFFFFFFFFFFFFFF, int(call_id) & 0xFFFFFFFF, capi.ptr(records), capi.ptr(xl),
        capi.ptr(tile_totals), capi.ptr(status), st)

    lens = torch.cat([tile_totals[: nt * A].to(torch.int64), centre_d])[src_d]
    # drop the empty units (later attempts of finished tiles) without a host sync: fixed-size nonzero, padding at the end
    if nu:
        nz = torch.nonzero_static(lens, size=nu, fill_value=nu).squeeze(1)
        lens = torch.cat([lens, lens.new_zeros(1)])[nz]
        units_d = torch.cat([units_d, units_d.new_zeros((1, 4))])[nz]
    unit_base = torch.zeros((nu + 1,), dtype=torch.int64, device=dev)
    if nu:
        torch.cumsum(lens, 0, out=unit_base[1:])

    cap = plan.capacity
    pts = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=dev)
    rgb = torch.empty((max(cap, 1), 3), dtype=out_dtype, device=dev)
    nrm = torch.empty((max(cap, 1), 3), dtype=out_dtype, device=dev) if (want_normals and normals is not None) else None

    # unit of every output chunk (one vectorised binary search instead of one per CTA)
    chunk = lib.g2pc_sample_emit_chunk_points()
    nchunks = (cap + chunk - 1) // chunk
    starts = torch.arange(nchunks + 1, dtype=torch.int64, device=dev) * chunk
    chunk_unit = (torch.searchsorted(unit_base, starts, right=True) - 1).clamp_(0, max(nu - 1, 0)).to(torch.int32)

    capi.call("g2pc_sample_emit", capi.ptr(records), capi.ptr(xl), n, capi.ptr(units_d), capi.ptr(unit_base),
              capi.ptr(chunk_unit), nu, int(seed) & 0xFFFFFFFFFFFFFFFF, int(call_id) & 0xFFFFFFFF, capi.ptr(pts),
              capi.ptr(rgb), capi.ptr(nrm), capi.dtype_code(rgb), cap, st)
    return pts, rgb, nrm, unit_base[nu], status, (records, xl, tile_totals, unit_base)


def dump_eps(gids, k, attempt, seed, call_id=0):
    """eps (k, n', 3) the sampler uses for Gaussians `gids` (int64 CUDA tensor) in `attempt` — for injecting the
    kernel's stream into the reference / oracle in parity tests."""
    lib = capi.load()
    capi.require_cuda(gids)
    gids = gids.to(torch.int64).contiguous()
    eps = torch.empty((k, gids.shape[0], 3), dtype=torch.float32, device=gids.device)
    capi.call("g2pc_dump_eps", capi.ptr(gids), gids.shape[0], int(k), int(attempt),
                                 int(seed) & 0xFFFFFFFFFFFFFFFF, int(call_id) & 0xFFFFFFFF, capi.ptr(eps),
                                 capi.stream_ptr(gids.device))
    return eps

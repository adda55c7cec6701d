"""Host-side geometry of the python renderer's tile quadtree (gauss_render.py:290-335).

The reference discovers its tiles with a host-driven BFS: a tile is dropped if w <= 1 or h <= 1, clamped to the image,
background-filled if no Gaussian overlaps it, split into 4 children of size (ceil(w/2), ceil(h/2)) if it is larger
than max_tile_size or holds more than max_gaussians_per_tile Gaussians, otherwise rendered as a leaf.

The geometry is separable: a 2-D node at level l is the product of an x-node and a y-node of two 1-D binary trees
over [0, W) and [0, H) (child offsets {0, ceil(size/2)}, clamp to the image before halving).  This module tabulates
those 1-D trees per level; the kernels (csrc/s3_preprocess.cu, s4_tree.cu) do per-Gaussian range queries on the
tables instead of testing every tile against every Gaussian.

Node order.  The reference's BFS queue appends the children as TL, BL, TR, BR (gauss_render.py:323-333), so the BFS
order of the nodes of one level is the lexicographic order of their child-rank paths, rank = 2*xbit + ybit.
"""
import numpy as np

MAX_LEVELS = 12  # levels 0..11 can be tabulated (2^11 = 2048 nodes per axis)
FLAG_DROPPED = 1  # pre-clamp size <= 1  -> the 2-D node is skipped (gauss_render.py:301)
FLAG_BIG = 2      # post-clamp size > max_tile_size -> forces a split (gauss_render.py:319)


def axis_levels(extent, max_tile_size, num_levels):
    """1-D tree over [0, extent).  Returns a list (one entry per level) of dicts of int arrays:
    start, end (inclusive, post-clamp), flags."""
    levels = []
    start = np.array([0], dtype=np.int64)
    pre = np.array([extent], dtype=np.int64)
    for l in range(num_levels):
        size = np.minimum(pre, extent - start)
        flags = np.where(pre <= 1, FLAG_DROPPED, 0) | np.where(size > max_tile_size, FLAG_BIG, 0)
        levels.append(dict(start=start.copy(), end=start + size - 1, size=size.copy(), flags=flags.astype(np.int64)))
        half = -(-size // 2)  # ceil(size / 2) of the clamped size
        nstart = np.empty(start.shape[0] * 2, dtype=np.int64)
        nstart[0::2] = start
        nstart[1::2] = start + half
        start = nstart
        pre = np.repeat(half, 2)
    return levels


class QuadtreeTables:
    """Per-level 1-D interval tables for one (width, height, max_tile_size)."""

    def __init__(self, width, height, max_tile_size, max_gaussians_per_tile, extra_levels=0):
        self.width, self.height = int(width), int(height)
        self.max_tile_size = int(max_tile_size)
        self.max_gaussians_per_tile = int(max_gaussians_per_tile)
        # geometric depth: first level at which no node is forced to split by its size
        probe_x = axis_levels(self.width, self.max_tile_size, MAX_LEVELS)
        probe_y = axis_levels(self.height, self.max_tile_size, MAX_LEVELS)
        depth = None
        for l in range(MAX_LEVELS):
            live_x = (probe_x[l]["flags"] & FLAG_DROPPED) == 0
            live_y = (probe_y[l]["flags"] & FLAG_DROPPED) == 0
            big = ((probe_x[l]["flags"][live_x] & FLAG_BIG).any() or (probe_y[l]["flags"][live_y] & FLAG_BIG).any())
            if not big:
                depth = l
                break
        if depth is None:
            raise ValueError("image too large for the tabulated quadtree depth")
        self.geo_depth = depth
        # the kernels' range queries need monotone starts / ends within a level; that holds until the nodes shrink to
        # ~2 pixels (accumulated 1-pixel overhangs), far below any useful tile size
        mono = MAX_LEVELS
        for l in range(MAX_LEVELS):
            bad = any((np.diff(ax[l]["start"]) < 0).any() or (np.diff(ax[l]["end"]) < 0).any() for ax in (probe_x, probe_y))
            if bad:
                mono = l
                break
        if mono <= depth:
            raise NotImplementedError("non-monotone quadtree level (unsupported image / tile size)")
        self.num_levels = min(mono, depth + 1 + extra_levels)
        self.x = probe_x[: self.num_levels]
        self.y = probe_y[: self.num_levels]
        self.off1 = np.array([(1 << l) - 1 for l in range(self.num_levels + 1)], dtype=np.int64)
        self.off2 = np.array([((1 << (2 * l)) - 1) // 3 for l in range(self.num_levels + 1)], dtype=np.int64)
        self.nodes_1d = int(self.off1[self.num_levels])
        self.nodes_2d = int(self.off2[self.num_levels])

    def flat(self):
        """(xs, xe, xf, ys, ye, yf) int32 arrays, levels concatenated (level l at offset 2^l - 1)."""
        cat = lambda ax, k: np.concatenate([lv[k] for lv in ax]).astype(np.int32)
        return (cat(self.x, "start"), cat(self.x, "end"), cat(self.x, "flags"),
                cat(self.y, "start"), cat(self.y, "end"), cat(self.y, "flags"))

    def pixel_luts(self):
        """Per level and axis, the node range of an interval as a lookup over PIXEL coordinates (uint16):
            lo[q]  = first node i with end_i   > q   (q = floor(rect_min));  2^level if none
            hi1[c] = 1 + last node i with start_i < c (c = ceil(rect_max));  0 if none
        — exactly the strict float compares of gauss_render.py:308-310 (starts / ends are integers and non-decreasing).
        Flat layout per level: [x lo (W)][x hi1 (W)][y lo (H)][y hi1 (H)], levels concatenated."""
        out = []
        for l in range(self.num_levels):
            for ax, extent in ((self.x[l], self.width), (self.y[l], self.height)):
                px = np.arange(extent)
                lo = np.searchsorted(ax["end"], px, side="right")     # ends <= q are skipped
                hi1 = np.searchsorted(ax["start"], px, side="left")    # starts < c
                out += [lo.astype(np.uint16), hi1.astype(np.uint16)]
        return np.concatenate(out)

    def clean_level_mask(self):
        """Bit l set iff no node of level l is dropped or degenerate (end < start) on either axis."""
        mask = 0
        for l in range(self.num_levels):
            ok = all(((ax[l]["flags"] & FLAG_DROPPED) == 0).all() and (ax[l]["end"] > ax[l]["start"]).all()
                     for ax in (self.x, self.y))
            if ok:
                mask |= 1 << l
        return mask

    def candidate_level_mask(self):
        """Bit l set iff level l has a live node that is not forced to split by its size (a leaf candidate)."""
        mask = 0
        for l in range(self.num_levels):
            fx, fy = self.x[l]["flags"], self.y[l]["flags"]
            live_x, live_y = (fx & FLAG_DROPPED) == 0, (fy & FLAG_DROPPED) == 0
            small_x, small_y = live_x & ((fx & FLAG_BIG) == 0), live_y & ((fy & FLAG_BIG) == 0)
            if small_x.any() and small_y.any():
                mask |= 1 << l
        return mask

    def max_leaf_pixels(self):
        return int(min(self.max_tile_size, self.width) * min(self.max_tile_size, self.height))


def interleave_key(ix, iy, level):
    """BFS rank key of node (ix, iy) at `level`: child rank = 2*xbit + ybit, most significant level first."""
    key = 0
    for b in range(level - 1, -1, -1):
        key = (key << 2) | (((ix >> b) & 1) << 1) | ((iy >> b) & 1)
    return key


def enumerate_tree(tables, count_fn):
    """Reference BFS replayed on the tables (host, for tests): count_fn(r0, c0, w, h) -> number of member Gaussians.
    Returns (leaves, background) with leaves = [(r0, c0, w, h, level, ix, iy)] in BFS order."""
    leaves, background = [], []
    frontier = [(0, 0)]
    for l in range(tables.num_levels):
        nxt = []
        X, Y = tables.x[l], tables.y[l]
        for (ix, iy) in frontier:
            if (X["flags"][ix] | Y["flags"][iy]) & FLAG_DROPPED:
                continue
            r0, c0, w, h = int(Y["start"][iy]), int(X["start"][ix]), int(X["size"][ix]), int(Y["size"][iy])
            cnt = count_fn(r0, c0, w, h)
            if cnt <= 0:
                background.append((r0, c0, w, h))
                continue
            forced = bool((X["flags"][ix] | Y["flags"][iy]) & FLAG_BIG)
            if forced or cnt > tables.max_gaussians_per_tile:
                if l + 1 >= tables.num_levels:
                    raise NotImplementedError("quadtree deeper than the tabulated levels")
                nxt += [(2 * ix, 2 * iy), (2 * ix, 2 * iy + 1), (2 * ix + 1, 2 * iy), (2 * ix + 1, 2 * iy + 1)]
                continue
            leaves.append((r0, c0, w, h, l, ix, iy))
        frontier = nxt
        if not frontier:
            break
    return leaves, background

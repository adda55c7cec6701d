"""CPU: invariants the colour kernels rely on, checked by brute force.

* g2pc/quadtree.py pixel_luts(): the node range of a rect looked up at floor(min) / ceil(max) is exactly the set the
  reference's strict float compares select (gauss_render.py:308-310) — what csrc/s3_preprocess.cu indexes per level;
* the CUDA back-end's super-tile lists (csrc/s7_tiles.cu): the depth-ordered list of a 2x2 super-tile, filtered by the
  packed tile rect with the kernel's unsigned-compare test, is the per-tile list of rasterizer_impl.cu:69-137;
* bench.py's clock sampler degrades to an empty summary where nvidia-smi is missing (this container)."""
import numpy as np
import pytest


@pytest.mark.parametrize("W,H,extra", [(1280, 720, 0), (1920, 1080, 2), (720, 405, 1), (200, 112, 2), (257, 129, 1)])
def test_pixel_luts_equal_the_strict_compares(W, H, extra):
    from g2pc import quadtree as qt
    T = qt.QuadtreeTables(W, H, 60, 60000, extra_levels=extra)
    luts = T.pixel_luts()
    per_level = 2 * (W + H)
    assert luts.shape[0] == per_level * T.num_levels and luts.dtype == np.uint16
    rng = np.random.default_rng(W + 31 * H + extra)
    f = np.float32
    for l in range(T.num_levels):
        L = luts[l * per_level:(l + 1) * per_level]
        for ax, extent, off in ((T.x[l], W, 0), (T.y[l], H, 2 * W)):
            lo_t, hi_t = L[off:off + extent], L[off + extent:off + 2 * extent]
            live = ((ax["flags"] & qt.FLAG_DROPPED) == 0) & (ax["end"] > ax["start"])
            # rect bounds as the kernel sees them: clipped to [0, extent-1], min < max; integers and near-integers included
            a = rng.uniform(0, extent - 1, 300).astype(f)
            b = rng.uniform(0, extent - 1, 300).astype(f)
            a[:40] = np.floor(a[:40]); b[40:80] = np.floor(b[40:80])
            a[80:100] = np.nextafter(np.floor(a[80:100]) + f(1), f(0)).astype(f)
            r0, r1 = np.minimum(a, b), np.maximum(a, b)
            keep = r1 > r0
            for x0, x1 in zip(r0[keep], r1[keep]):
                lo = int(lo_t[int(np.floor(x0))])
                hi = int(hi_t[int(np.ceil(x1))]) - 1
                brute = [k for k in range(1 << l)
                         if min(x1, f(ax["end"][k])) > max(x0, f(ax["start"][k]))]
                got = list(range(lo, hi + 1)) if lo <= hi else []
                if ((T.clean_level_mask() >> l) & 1):
                    assert got == brute, (l, x0, x1, got, brute)
                else:  # dropped / degenerate nodes are filtered afterwards by the kernel's axis_member()
                    assert [k for k in got if live[k]] == [k for k in brute if live[k]], (l, x0, x1)


def _pack(x0, x1, y0, y1):
    return (np.uint32(x0) | (np.uint32(x1) << np.uint32(8)) | (np.uint32(y0) << np.uint32(16)) |
            (np.uint32(y1) << np.uint32(24)))


@pytest.mark.parametrize("gx,gy", [(80, 45), (45, 26), (13, 7), (1, 1), (2, 3)])
def test_super_tile_lists_filtered_by_tile_rect_are_the_tile_lists(gx, gy):
    rng = np.random.default_rng(gx * 100 + gy)
    n = 3000
    x0 = rng.integers(0, gx, n); y0 = rng.integers(0, gy, n)
    x1 = np.minimum(gx - 1, x0 + rng.geometric(0.45, n) - 1); y1 = np.minimum(gy - 1, y0 + rng.geometric(0.45, n) - 1)
    rect = _pack(x0, x1, y0, y1)                                   # record slot q2.w
    srange = _pack(x0 >> 1, x1 >> 1, y0 >> 1, y1 >> 1)             # sort value: the super-tile rect
    sgx, sgy = (gx + 1) // 2, (gy + 1) // 2
    # multisplit over the super-tile grid, Gaussians already in depth order (index order here)
    lists = {}
    for g in range(n):
        sx0, sx1 = int(srange[g] & 255), int((srange[g] >> 8) & 255)
        sy0, sy1 = int((srange[g] >> 16) & 255), int(srange[g] >> 24)
        for sy in range(sy0, sy1 + 1):
            for sx in range(sx0, sx1 + 1):
                assert sx < sgx and sy < sgy
                lists.setdefault((sx, sy), []).append(g)
    u32 = np.uint32
    old = np.seterr(over="ignore")  # the unsigned wrap-around IS the test
    for ty in range(gy):
        for tx in range(gx):
            want = [g for g in range(n) if x0[g] <= tx <= x1[g] and y0[g] <= ty <= y1[g]]
            got = []
            for g in lists.get((tx >> 1, ty >> 1), []):
                r = rect[g]
                # the kernel's membership test (blend_tiles_kernel): two unsigned compares per axis folded into one
                skip = (u32(tx) - (r & u32(255)) > ((r >> u32(8)) & u32(255)) - (r & u32(255)) or
                        u32(ty) - ((r >> u32(16)) & u32(255)) > (r >> u32(24)) - ((r >> u32(16)) & u32(255)))
                if not skip:
                    got.append(g)
            assert got == want, (tx, ty)
    np.seterr(**old)


def test_clock_sampler_without_nvidia_smi(monkeypatch):
    import shutil
    import bench
    if shutil.which("nvidia-smi"):
        pytest.skip("nvidia-smi present: covered on the GPU box by the bench line's `clocks`")
    with bench.ClockSampler(range(2)) as clk:
        clk.mark_begin()
        clk.mark_end()
    s = clk.summary()
    assert s["samples"] == 0 and s["sm_mhz"] is None and s["reasons"] == [] and s["gpus_watched"] == [0, 1]
    with bench.ClockSampler(range(1), enabled=False) as clk2:
        pass
    assert clk2.summary()["samples"] == 0

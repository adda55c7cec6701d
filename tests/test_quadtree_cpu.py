"""CPU: the separable per-level interval tables (g2pc/quadtree.py) against the oracle's explicit BFS
(oracle/render.py: quadtree_leaves, restating gauss_render.py:290-335), including count-driven splits, and the
kernel's leaf numbering rule (level-major, then interleaved child-rank key) against the BFS order."""
import numpy as np
import pytest


def _rects(rng, W, H, n):
    mx = rng.uniform(-20, W + 20, n).astype(np.float32)
    my = rng.uniform(-20, H + 20, n).astype(np.float32)
    rad = (3 * np.ceil(rng.gamma(1.5, 2.0, n))).astype(np.float32)
    f = np.float32
    return (np.clip(mx - rad, f(0), f(W - 1)), np.clip(my - rad, f(0), f(H - 1)),
            np.clip(mx + rad, f(0), f(W - 1)), np.clip(my + rad, f(0), f(H - 1)))


@pytest.mark.parametrize("W,H,mt,mg", [(200, 112, 60, 60000), (1280, 720, 60, 60000), (720, 405, 60, 60000),
                                       (1920, 1080, 60, 60000), (330, 245, 60, 300), (330, 185, 60, 150),
                                       (180, 101, 60, 60000), (257, 129, 32, 200)])
def test_tables_enumerate_the_reference_tree(W, H, mt, mg):
    from g2pc import quadtree as qt
    from oracle import render as orr
    rng = np.random.default_rng(W * 7 + H)
    rx0, ry0, rx1, ry1 = _rects(rng, W, H, 4000)
    leaves, bg = orr.quadtree_leaves(W, H, rx0, ry0, rx1, ry1, mt, mg)
    f = np.float32

    def count(r0, c0, w, h):
        return int(((np.minimum(rx1, f(c0 + w - 1)) > np.maximum(rx0, f(c0))) &
                    (np.minimum(ry1, f(r0 + h - 1)) > np.maximum(ry0, f(r0)))).sum())

    T = None
    for extra in range(0, 6):  # the renderer adds levels on demand (HDR_NEED_DEEPER)
        T = qt.QuadtreeTables(W, H, mt, mg, extra_levels=extra)
        try:
            l2, b2 = qt.enumerate_tree(T, count)
            break
        except NotImplementedError:
            continue
    else:
        pytest.fail("tree deeper than the supported levels")
    assert [(a[0], a[1], a[2], a[3]) for a in leaves] == [(a[0], a[1], a[2], a[3]) for a in l2]
    assert bg == b2
    # the tree kernel numbers leaves level-major, then by the interleaved child-rank key: must equal the BFS order
    keyed = sorted(range(len(l2)), key=lambda i: (l2[i][4], qt.interleave_key(l2[i][5], l2[i][6], l2[i][4])))
    assert keyed == list(range(len(l2)))
    # separable member range query == brute force, at every tabulated level
    g = rng.integers(0, rx0.shape[0], 50)
    for l in range(T.num_levels):
        X, Y = T.x[l], T.y[l]
        for i in g:
            mx = [k for k in range(1 << l) if not (X["flags"][k] & qt.FLAG_DROPPED) and X["end"][k] > X["start"][k]
                  and min(rx1[i], f(X["end"][k])) > max(rx0[i], f(X["start"][k]))]
            if mx:
                assert mx == list(range(mx[0], mx[-1] + 1)) or all(
                    (X["flags"][k] & qt.FLAG_DROPPED) or X["end"][k] <= X["start"][k]
                    for k in set(range(mx[0], mx[-1] + 1)) - set(mx)), "members must form a contiguous index range"


def test_candidate_level_mask():
    from g2pc import quadtree as qt
    T = qt.QuadtreeTables(1280, 720, 60, 60000)
    assert T.geo_depth == 5 and T.num_levels == 6 and T.candidate_level_mask() == 0b100000
    T2 = qt.QuadtreeTables(330, 185, 60, 150, extra_levels=2)
    assert T2.candidate_level_mask() == 0b111000

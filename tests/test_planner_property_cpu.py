"""CPU, hypothesis: invariants of the host-side sampling plan for arbitrary points-per-Gaussian histograms."""
import numpy as np
from hypothesis import given, settings, strategies as st


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=400), min_size=3, max_size=60), st.booleans(),
       st.integers(min_value=1, max_value=8))
def test_plan_invariants(hist, exact, attempts):
    from g2pc import sampler
    hist = np.asarray(hist, dtype=np.int64)
    if np.count_nonzero(hist) < 2:
        return  # the reference's heuristic needs two distinct values (numpy.gradient)
    bins = sampler.plan_bins(hist, exact)
    # bins are disjoint, ascending value ranges; every Gaussian is in at most one bin
    prev_end = -1.0
    covered = 0
    for (start, end, n, count) in bins:
        assert start >= prev_end and end > start and n >= 1 and count >= 1
        assert n == int(np.floor(start + (end - start) / 2))
        prev_end = end
        covered += count
    assert covered <= int(hist.sum())
    if exact:
        # every occurring value v >= 1 forms its own bin [v, next) and receives floor of the midpoint
        assert covered == int(hist[1:].sum()) + (int(hist[0]) if (hist[0] and bins and bins[0][0] == 0) else 0)
    plan = sampler.SamplePlan([(n - 1, c) for (_, _, n, c) in bins], attempts)
    assert plan.n == covered
    assert plan.capacity == sum(c * n for (_, _, n, c) in bins)
    # tiles tile [0, n) contiguously; every sample unit points at a tile's Gaussians
    j = 0
    for (j0, cnt, k, lpg) in plan.tiles:
        assert j0 == j and 1 <= cnt <= 256 // lpg
        j += cnt
    assert j == plan.n
    tiles = {(int(t[0]), int(t[1])) for t in plan.tiles}
    for (a, j0, cnt, k) in plan.units:
        if a >= 0:
            assert (int(j0), int(cnt)) in tiles and 0 <= a < attempts and k >= 1
    assert plan.unit_src.shape[0] == plan.units.shape[0]
    assert plan.unit_src.max(initial=-1) < plan.tiles.shape[0] * attempts + plan.centre_lens.shape[0]

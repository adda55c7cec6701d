"""CPU: host-side bin planning (g2pc/sampler.py) against the oracle's restatement of gauss_to_pc.py:105-138,308-343."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("n,P,exact", [(5000, 50000, False), (5000, 50000, True), (20000, 300000, False),
                                       (300, 1000, False), (1000, 100, False), (2000, 2000, True)])
def test_bins_match_oracle(n, P, exact):
    from g2pc import sampler, synth
    from oracle import gaussians as og, sampling as osamp
    sc = synth.make_scene(n, seed=1235)
    cov, _ = og.validate_covariances(og.build_covariance(sc["scales"], sc["rots"]))
    ppg = osamp.distribute_points(og.gaussian_magnitudes(cov, sc["opacities"]), P).to(torch.int32)
    ob = []
    for (s, e, k) in osamp.make_bins(ppg, exact):
        c = int(((ppg >= s) & (ppg < e)).sum())
        if c >= 1:
            ob.append((s, e, k, c))
    pb = sampler.plan_bins(torch.bincount(ppg.long()).numpy(), exact)
    assert pb == ob


def test_plan_tables_are_consistent():
    from g2pc import sampler
    bins = [(0, 700), (3, 513), (20, 40), (700, 3), (5000, 1)]
    plan = sampler.SamplePlan(bins, 5)
    assert plan.n == sum(c for _, c in bins)
    # tiles partition [0, n) in order, never straddle a bin, respect 256/lpg
    j = 0
    for (j0, cnt, k, lpg) in plan.tiles:
        assert j0 == j and 1 <= cnt <= 256 // lpg and lpg & (lpg - 1) == 0
        j += cnt
    assert j == plan.n
    # units: one centre unit per bin, attempts x tiles sample units for k > 0, in output order
    centre = plan.units[plan.units[:, 0] < 0]
    assert [int(c) for c in centre[:, 2]] == [c for _, c in bins]
    assert plan.capacity == sum(c * (k + 1) for k, c in bins)
    assert plan.units.shape[0] == len(bins) + 5 * sum(1 for (j0, cnt, k, lpg) in plan.tiles if k > 0)


def test_too_few_distinct_counts_raises_like_reference():
    from g2pc import sampler
    with pytest.raises(ValueError):
        sampler.plan_bins(np.array([0, 10]), False)  # a single distinct value: numpy.gradient needs >= 2

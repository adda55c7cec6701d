"""CPU: the host-built tile / unit tables (g2pc/sampler.py) drive a numpy emulation of the two S2 passes; the emulated
output must equal the oracle's point cloud — i.e. the tables encode the reference's bin / attempt / Gaussian order.
(The real kernels are checked against the oracle on the GPU; this test pins the host logic without one.)"""
import numpy as np
import pytest
import torch


def emulate(plan, perm, xyz, cov, colours, normals, gids, attempts, std, eps_fn):
    """count pass -> (xl, tile_totals); unit lengths -> unit_base; emit pass -> points, all in numpy/torch-CPU."""
    from oracle import sampling as osamp
    A = plan.attempts_stored
    n = plan.n
    xl = np.zeros((A, n), dtype=np.int64)
    tt = np.zeros((plan.tiles.shape[0], A), dtype=np.int64)
    L = torch.linalg.cholesky(cov)
    draws = {}
    for t, (j0, cnt, k, lpg) in enumerate(plan.tiles):
        if k <= 0:
            continue
        rows = perm[j0:j0 + cnt]
        added = np.zeros(cnt, dtype=np.int64)
        for a in range(attempts):
            todo = np.nonzero(added != k)[0]
            if todo.size == 0:
                break
            eps = torch.as_tensor(eps_fn(gids[rows[todo]], k, a))                       # (k, n', 3)
            x = xyz[rows[todo]].unsqueeze(0) + torch.matmul(L[rows[todo]].unsqueeze(0), eps.unsqueeze(-1)).squeeze(-1)
            samples = x.transpose(0, 1).contiguous().view(-1, 3)
            d = osamp.mahalanobis(torch.repeat_interleave(xyz[rows[todo]], k, dim=0), samples,
                                  torch.repeat_interleave(cov[rows[todo]], k, dim=0))
            counts = (d <= std).view(-1, k).sum(1).numpy()
            m = np.zeros(cnt, dtype=np.int64)
            m[todo] = np.minimum(k - added[todo], counts)
            added[todo] = np.minimum(k, added[todo] + counts)
            if a < A:
                xl[a, j0:j0 + cnt] = np.concatenate([[0], np.cumsum(m)[:-1]])
                tt[t, a] = m.sum()
            for li in todo:
                draws[(j0 + li, a)] = None
            draws[(t, a)] = (todo, x)  # x[s, i] = sample s of todo[i]
    lens = np.concatenate([tt.reshape(-1), plan.centre_lens])[plan.unit_src]
    pts, cols, nrms = [], [], []
    tile_of = {(int(t[0]), int(t[1])): i for i, t in enumerate(plan.tiles)}
    for (a, j0, cnt, k), ln in zip(plan.units, lens):
        rows = perm[j0:j0 + cnt]
        if a < 0:
            pts.append(xyz[rows]); cols.append(colours[rows]); nrms.append(normals[rows])
            continue
        if ln == 0:
            continue
        t = tile_of[(int(j0), int(cnt))]
        todo, x = draws[(t, a)]
        x_full = {int(li): x[:, i] for i, li in enumerate(todo)}
        m = np.diff(np.concatenate([xl[a, j0:j0 + cnt], [ln]]))
        for li in range(cnt):
            if m[li] > 0:
                pts.append(x_full[li][: m[li]])
                cols.append(colours[rows[li]].unsqueeze(0).repeat(int(m[li]), 1))
                nrms.append(normals[rows[li]].unsqueeze(0).repeat(int(m[li]), 1))
    return torch.cat(pts, 0), torch.cat(cols, 0), torch.cat(nrms, 0)


@pytest.mark.parametrize("n,P,exact,attempts", [(1500, 12000, False, 5), (600, 9000, True, 100), (400, 40000, False, 5)])
def test_tables_reproduce_reference_order(n, P, exact, attempts):
    from g2pc import config, sampler, synth
    from oracle import gaussians as og, philox, sampling as osamp
    sc = synth.make_scene(n, seed=1400 + n)
    cov, _ = og.validate_covariances(og.build_covariance(sc["scales"], sc["rots"]))
    nrm = og.calculate_normals(sc["scales"], sc["rots"])
    mags = og.gaussian_magnitudes(cov, sc["opacities"])
    colours = sc["colours"] * 255
    eps_fn = lambda g, k, a: philox.draw_eps(g, k, a, 11, 0)
    o = osamp.generate_pointcloud(sc["xyz"], cov, colours, nrm, mags, P, exact_num_points=exact,
                                  num_sample_attempts=attempts, eps_fn=eps_fn)
    ppg = o["ppg"].to(torch.int64)
    hist = torch.bincount(ppg).numpy()
    bins = sampler.plan_bins(hist, exact)
    lut = np.full((hist.shape[0],), len(bins), dtype=np.int64)
    for b, (s, e, _, _) in enumerate(bins):
        lut[int(np.ceil(s)):int(np.ceil(e))] = b
    bin_of = lut[ppg.numpy()]
    order = np.argsort(bin_of, kind="stable")
    perm = order[: sum(c for (_, _, _, c) in bins)]
    A = min(attempts, config.ATTEMPTS_STORED_FIRST)
    plan = sampler.SamplePlan([(k - 1, c) for (_, _, k, c) in bins], A)
    pts, cols, nr = emulate(plan, perm, sc["xyz"], cov, colours, nrm, np.arange(n), attempts, 2.0, eps_fn)
    assert pts.shape == o["points"].shape
    assert torch.equal(pts, o["points"]), "tile/unit tables do not reproduce the reference's output order"
    assert torch.equal(cols, o["colours"]) and torch.equal(nr, o["normals"])

"""Generate the golden vectors that pin the oracle: outputs of the UNMODIFIED reference (/root/reference) run on
CPU through oracle/ref_shim.py on small seeded scenes.  Run in the build container (the reference does not exist on
the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/sampling_*.npz (and colour_*.npz).  Inputs are regenerated from the seeds by g2pc.synth, so only
outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3dgs-to-pc_b200"))

from oracle import philox, ref_shim  # noqa: E402
from g2pc import synth  # noqa: E402

SAMPLING_CASES = {
    # name: (n_gaussians, scene_seed, num_points, exact_num_points, attempts, rng_seed)
    "sampling_a": (1500, 1301, 12000, False, 5, 42),
    "sampling_b": (800, 1302, 9000, True, 100, 43),
}


def make_sampling(name, n, scene_seed, num_points, exact, attempts, rng_seed):
    ref = ref_shim.load()
    sc = synth.make_scene(n, seed=scene_seed)
    eps_fn = lambda g, k, a: philox.draw_eps(g, k, a, rng_seed, 0)
    with ref_shim.cpu_redirect():
        G = ref.gauss_handler.Gaussians(sc["xyz"].clone(), sc["scales"].clone(), sc["rots"].clone(),
                                        sc["colours"].clone() * 255, sc["opacities"].clone())
        G.calculate_normals()
        cov0 = G.covariances.clone()
        keep = G.validate_covariances()
        mags = G.get_gaussian_magnitudes()
        ppg = ref.gauss_to_pc.distribute_points(mags, num_points).type(torch.int)
        with ref_shim.EpsInjector(ref, G.xyz, eps_fn) as inj:
            pts, cols, nrm = ref.gauss_to_pc.generate_pointcloud(
                G, num_points, mahalanobis_distance_std=2.0, exact_num_points=exact,
                num_sample_attempts=attempts, device="cpu", quiet=True)
            calls = np.array([(k, a, len(g)) for (k, a, g) in inj.log], dtype=np.int64)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        meta=np.array([n, scene_seed, num_points, int(exact), attempts, rng_seed], dtype=np.int64),
        cov0=cov0.numpy(), cov=G.covariances.numpy(), keep=keep.numpy(), normals=G.normals.numpy(),
        magnitudes=mags.numpy(), ppg=ppg.numpy(), points=pts.numpy(), colours=cols.numpy().astype(np.float32),
        point_normals=nrm.numpy().astype(np.float32), mvn_calls=calls)
    print(name, "points", tuple(pts.shape), "mvn calls", calls.shape[0])


if __name__ == "__main__":
    torch.manual_seed(0)
    for name, args in SAMPLING_CASES.items():
        make_sampling(name, *args)
    if "--colour" in sys.argv or True:
        try:
            from make_golden_colour import make_all
            make_all()
        except ImportError:
            pass

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


COLOUR_CASES = {
    # name: (n_gaussians, scene_seed, n_cameras, colour_resolution)
    "colour_a": (2500, 1310, 2, 200),
    "colour_b": (1200, 1311, 3, 180),
}


def make_colour(name, n, scene_seed, ncams, res):
    """GaussPythonRenderer of the reference (gauss_render.py:210-465), tile parameters pinned to (60, 60000)."""
    ref = ref_shim.load()
    sc = synth.make_scene(n, seed=scene_seed)
    cams, intr = synth.make_cameras(ncams)
    with ref_shim.cpu_redirect(pinned_tiles=(60, 60000)):
        G = ref.gauss_handler.Gaussians(sc["xyz"].clone(), sc["scales"].clone(), sc["rots"].clone(),
                                        sc["colours"].clone(), sc["opacities"].clone())
        R = ref.gauss_render.get_renderer("python", G.xyz, torch.unsqueeze(torch.clone(G.opacities), 1), G.colours,
                                          G.covariances, visible_gaussian_threshold=0.05)
        imgs = []
        for c2w, k in zip(cams, intr):
            cam = ref.camera_handler.get_camera("python", c2w.clone(), k, colour_resolution=res)
            img, _, _, _ = R(cam)
            imgs.append(img.numpy().astype(np.float32))
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            meta=np.array([n, scene_seed, ncams, res], dtype=np.int64),
                            max_contribution=R.gaussian_max_contribution.numpy(),
                            colours=R.gaussian_colours.numpy(), images=np.stack(imgs),
                            visible=R.get_visible_gaussians().numpy())
    print(name, "images", np.stack(imgs).shape, "seen", int((R.gaussian_max_contribution > 0).sum()))


def make_sh(name="sh_a", n=400, seed=1320):
    """eval_sh of the reference (gauss_render.py:43-99), degrees 0..3, on seeded coefficients / unit directions; the
    rendered colour is eval_sh + 0.5 clamped at 0 (forward.cu:65-72)."""
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(seed)
    sh = (0.4 * torch.randn(n, 3, 16, generator=g)).float()
    d = torch.randn(n, 3, generator=g)
    d = (d / d.norm(dim=1, keepdim=True)).float()
    out = {f"deg{deg}": ref.gauss_render.eval_sh(deg, sh[..., : (deg + 1) ** 2], d).numpy() for deg in range(4)}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array([n, seed], dtype=np.int64), **out)
    print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    make_sh()
    for name, args in SAMPLING_CASES.items():
        make_sampling(name, *args)
    for name, args in COLOUR_CASES.items():
        make_colour(name, *args)

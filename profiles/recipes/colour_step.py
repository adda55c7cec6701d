#!/usr/bin/env python
"""Driver for the ncu recipes: render a few cameras of a bench.py workload through the colour pipeline (the kernels the
captures name: preprocess_kernel, tree_kernel, ms_count / ms_scan / ms_scatter, blend_kernel, accumulate_kernel).

    python profiles/recipes/colour_step.py --workload c3 --cams 3 [--renderer python|cuda] [--strict]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3dgs-to-pc_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import camera_handler as ch  # noqa: E402
import gauss_handler as gh  # noqa: E402
import gauss_render as gr  # noqa: E402
from g2pc import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--cams", type=int, default=3)
ap.add_argument("--first-cam", type=int, default=0)
ap.add_argument("--renderer", default="python")
ap.add_argument("--strict", action="store_true")
ap.add_argument("--sync", action="store_true", help="confirm every frame before the next (no skipped launches after a replay)")
ap.add_argument("--surface", action="store_true", help="renderer cuda: also compute the surface distances (C4)")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = "cuda:0"
sc = bench._scene_for(wl)
d = {k: v.to(dev) for k, v in sc.items()}
G = gh.Gaussians(d["xyz"], d["scales"], d["rots"], d["colours"], d["opacities"])
R = gr.get_renderer(a.renderer, G.xyz, G.opacities.unsqueeze(1), G.colours, G.covariances,
                    shs=d["shs"] if wl["sh"] > 0 else None, visible_gaussian_threshold=0.05,
                    **(dict(calculate_surface_distance=True, surface_distance_std=2.0) if a.surface else {}))
if a.strict:
    R.t_stop = 0.0
R.async_mode = not a.sync
cams, intr = synth.make_cameras(wl["cams"])
for c, k in list(zip(cams, intr))[a.first_cam:a.first_cam + a.cams]:
    R(ch.get_camera(a.renderer, c, k, colour_resolution=wl["res"]))
R.flush()
torch.cuda.synchronize()
print(R.last_stats, "pairs", R.executed_pairs() if hasattr(R, "executed_pairs") else None)

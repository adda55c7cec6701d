"""torchrun -n 2 helper of tests/test_colour_gpu.py::test_sharded_equals_single_gpu_nccl: the sharded pipeline
(cameras sharded for the colour stage, Gaussian index ranges for the sampler, NCCL merges) against the single-process
pipeline on the same scene; rank 0 compares the 36-byte rows of both clouds as sets."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3dgs-to-pc_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import gauss_to_pc as g2p  # noqa: E402
from g2pc import dist as gd, sampler, synth  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
wl = dict(n=200_000, cams=6, points=800_000, res=720, sh=3, colours=True, seed=77)
st = bench.settings_for(wl, g2p, dev)
sc = synth.make_scene(wl["n"], seed=wl["seed"], sh_degree=3)
d = {k: v.to(dev) for k, v in sc.items()}
cams, intr = synth.make_cameras(wl["cams"])
tr = {f"c{i}": c for i, c in enumerate(cams)}
ik = {f"c{i}": k for i, k in enumerate(intr)}
sampler.reset_call_counter(0)
pc = gd.convert_gaussians_to_pc_sharded(d, tr, ik, st, render_shs=True)
cnt = torch.tensor([pc.points.shape[0]], device=dev)
allc = [torch.zeros_like(cnt) for _ in range(world)]
dist.all_gather(allc, cnt)
mx = int(max(c.item() for c in allc))
pad = torch.zeros((mx, 9), device=dev)
pad[: pc.points.shape[0]] = torch.cat([pc.points, pc.colours, pc.normals], 1)
allp = [torch.zeros_like(pad) for _ in range(world)]
dist.all_gather(allp, pad)
ok = True
if rank == 0:
    sharded = torch.cat([allp[r][: int(allc[r].item())] for r in range(world)], 0)
    sampler.reset_call_counter(0)
    ref, _ = g2p.convert_gaussians_to_pc(d["xyz"], d["scales"], d["rots"], d["colours"].clone(), d["opacities"], d["shs"],
                                         tr, ik, None, st, render_shs=True)
    single = torch.cat([ref.points, ref.colours, ref.normals], 1)

    def rows(x):
        return set(map(bytes, np.ascontiguousarray(x).view(np.uint8).reshape(x.shape[0], -1)))

    ra, rb = rows(sharded.cpu().numpy()), rows(single.cpu().numpy())
    print("points sharded", sharded.shape[0], "single", single.shape[0], "identical rows", len(ra & rb),
          "only sharded", len(ra - rb), "only single", len(rb - ra))
    ok = sharded.shape[0] == single.shape[0] and len(ra - rb) == 0 and len(rb - ra) == 0
    if ok:
        print("DIST_CHECK_OK")
dist.barrier()
dist.destroy_process_group()
sys.stdout.flush()
sys.exit(0 if ok else 1)

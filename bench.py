#!/usr/bin/env python
"""bench.py — Mpoints/s (sample + colour) of the 3DGS-to-PC hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's CPU path: the oracle port, host cores)

One "step" = one pass of the hot path over the whole synthetic scene: covariance build (S1) -> colour stage over all
cameras (S3-S6, renderer_type=python semantics, SH evaluated per camera) -> visibility cull -> validate covariances ->
magnitudes / points-per-Gaussian / bins -> sampling + Mahalanobis cull (S2).  `value` is measured with the scene
already resident in HBM; `e2e` goes through the same public call (gauss_to_pc.convert_gaussians_to_pc) with pinned
HOST buffers, host->device copies of the scene and the device->host read of the point cloud inside the timed region.

Workloads (BASELINE.json configs): c3 = 3M Gaussians / 200 cameras / 10M points / 1280x720 / SH deg 3 / visibility
0.05 (headline, default); c2 = 1M / 50 / 10M / 720x405 / SH deg 2; c1 = 10k Gaussians, 100k points, no colour stage.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "3dgs-to-pc_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: n_gaussians, n_cameras, num_points, colour_quality width, sh_degree, render colours
    "c3": dict(n=3_000_000, cams=200, points=10_000_000, res=1280, sh=3, colours=True, seed=1234 + 2),
    "c2": dict(n=1_000_000, cams=50, points=10_000_000, res=720, sh=2, colours=True, seed=1234 + 1),
    "c1": dict(n=10_000, cams=0, points=100_000, res=None, sh=0, colours=False, seed=1234 + 0),
    "tiny": dict(n=100_000, cams=4, points=400_000, res=720, sh=3, colours=True, seed=1234 + 9),
}
METRIC = "Mpoints/sec (sample+colour) at 3M Gaussians/200 cams, 1/2/4/8 B200 vs CPU ref"
UNIT = "Mpoints/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-gaussians", type=int, default=30000)
    ap.add_argument("--cpu-sample-cams", type=int, default=2)
    return ap.parse_args()


def settings_for(wl, g2p, device):
    return g2p.GaussPointCloudSettings(
        renderer_type="python", num_points=wl["points"], prioritise_visible_gaussians=True,
        mahalanobis_distance_std=2.0, camera_skip_rate=0, render_colours=wl["colours"], min_opacity=0.0,
        bounding_box_min=None, bounding_box_max=None, calculate_normals=True, cull_large_percentage=0.0,
        remove_unrendered_gaussians=True, colour_resolution=wl["res"], max_sh_degree=wl["sh"], exact_num_points=False,
        visibility_threshold=0.05, surface_distance_std=None, generate_mesh=False, quiet=True, device=device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.strip().split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from g2pc import build, capi, sampler, synth
    build.build()
    capi.load()
    import gauss_to_pc as g2p

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # torchrun pins OMP_NUM_THREADS=1: give every rank its share of the host cores for the (untimed) scene synthesis
    torch.set_num_threads(max(1, min(16, (os.cpu_count() or 1) // world)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    wl = WORKLOADS[args.workload]
    st = settings_for(wl, g2p, dev)

    sc = _scene_for(wl)
    cams, intr = synth.make_cameras(wl["cams"]) if wl["cams"] else ([], [])
    transforms = {f"cam{i:04d}": c for i, c in enumerate(cams)}
    intrinsics = {f"cam{i:04d}": k for i, k in enumerate(intr)}
    host = {k: v.pin_memory() for k, v in sc.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values()) + len(cams) * 64

    def upload():
        return {k: v.to(dev, non_blocking=True) for k, v in host.items()}

    if world > 1:
        from g2pc import dist as gdist
        runner = lambda d: gdist.convert_gaussians_to_pc_sharded(d, transforms, intrinsics, st, render_shs=wl["sh"] > 0)
    else:
        def runner(d):
            pc, _ = g2p.convert_gaussians_to_pc(d["xyz"], d["scales"], d["rots"], d["colours"].clone(), d["opacities"],
                                                d["shs"], transforms if wl["colours"] else None, intrinsics, None, st,
                                                render_shs=wl["colours"] and wl["sh"] > 0)
            return pc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = upload()
    torch.cuda.synchronize()

    def step_resident():
        sampler.reset_call_counter(0)
        return runner(resident)

    out_host = {}

    def step_e2e():
        sampler.reset_call_counter(0)
        d = upload()
        pc = runner(d)
        for name, t in (("points", pc.points), ("colours", pc.colours), ("normals", pc.normals)):
            if t is None:
                continue
            buf = out_host.get(name)
            if buf is None or buf.shape[0] < t.shape[0]:
                buf = torch.empty((int(t.shape[0] * 1.05) + 16, 3), dtype=t.dtype).pin_memory()
                out_host[name] = buf
            buf[: t.shape[0]].copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        return pc

    def timed(fn, steps):
        """CUDA-event timing of `steps` calls, barrier + synchronize on both sides, max over ranks."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pc = None
        for _ in range(steps):
            pc = fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        npts = torch.tensor([pc.points.shape[0]], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(npts, op=dist.ReduceOp.SUM)
        return float(ms.item()), int(npts.item()), pc

    for _ in range(max(args.warmup, 3)):
        step_resident()
    capi.LAUNCHES = 0
    capi.TIMING = {"g2pc_sample_emit": [], "g2pc_sample_count": [], "g2pc_blend": [], "g2pc_preprocess": []}
    with ClockSampler(local) as clk:
        ms, npts, pc = timed(step_resident, args.steps)
    launches = capi.LAUNCHES
    timing = {k: [a.elapsed_time(b) for (a, b) in v] for k, v in capi.TIMING.items()}
    capi.TIMING = None
    ms_step = ms / args.steps
    value = npts / (ms_step * 1e-3) / 1e6

    step_e2e()
    e2e_ms, e2e_pts, pc2 = timed(step_e2e, args.steps)
    e2e_step = e2e_ms / args.steps
    d2h_bytes = sum(t.numel() * t.element_size() for t in (pc2.points, pc2.colours, pc2.normals) if t is not None)

    # ---- roofline of the kernel north_star's HBM target names: the S2 emit kernel ----------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    emit_ms = float(np.mean(timing["g2pc_sample_emit"])) if timing["g2pc_sample_emit"] else None
    n_active = int(getattr(g2p, "LAST_SAMPLE_STATS", {}).get("n_active", 0))
    p_rank = int(pc.points.shape[0])
    alg_bytes = n_active * 44 + p_rank * 36  # 12 mu + 24 Sigma/L + 8 count/offset per Gaussian; 36 B/point
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_emit_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roof = None
    if emit_ms:
        ach = alg_bytes / (emit_ms * 1e-3) / 1e9
        roof = {"kernel": "sample_emit_kernel", "bound": "hbm", "achieved": round(ach, 1), "peak": peak_gbs,
                "unit": "GB/s", "frac": round(ach / peak_gbs, 4), "traffic": traffic,
                "peak_source": "measured" if peaks else "fallback", "alg_bytes_per_launch": alg_bytes,
                "avg_launch_ms": round(emit_ms, 4)}
    kernel_ms = {k.replace("g2pc_", ""): round(float(np.sum(v)) / args.steps, 3) for k, v in timing.items() if v}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_base = None
    if not args.no_cpu_baseline and world == 1:
        cpu_base = cpu_baseline(wl, args.cpu_sample_gaussians, args.cpu_sample_cams)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['n']} Gaussians, {wl['cams']} cameras, {wl['points']} points, "
                               f"width {wl['res']}, SH deg {wl['sh']}, visibility_threshold 0.05, renderer_type=python "
                               "semantics", "points_out": npts,
                   "l2": ("inputs larger than L2 (per-step working set >> 126 MB)" if h2d_bytes > 4 * 126e6 else
                          "working set below L2 and not flushed (non-headline workload)"),
                   "parallelism": "1 GPU" if world == 1 else f"cameras sharded x{world} (colour), Gaussians sharded x{world} (sampling)"},
        "e2e": {"value": round(e2e_pts / (e2e_step * 1e-3) / 1e6, 3), "unit": UNIT, "ms_per_step": round(e2e_step, 3),
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": roof,
        "kernel_ms_per_step": kernel_ms,
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------------
_SCENE_CACHE = {}


def _scene_for(wl):
    """The synthetic scene is a pure function of the workload: generate it once per process."""
    from g2pc import synth
    key = (wl["n"], wl["seed"], wl["sh"])
    if key not in _SCENE_CACHE:
        _SCENE_CACHE[key] = synth.make_scene(wl["n"], seed=wl["seed"], sh_degree=wl["sh"])
    return _SCENE_CACHE[key]


def cpu_sample_run(wl, n_s, cams_s, threads):
    """The oracle port (CPU restatement of the reference's python path) on a bounded sample of the workload:
    the first n_s Gaussians of the scene, the first cams_s cameras at full resolution, and num_points scaled by
    (n_s / n) * (cams_s / cams) so that the Gaussian-camera work per emitted point equals the full workload's."""
    from g2pc import synth
    from oracle import gaussians as og, render as orr, sampling as osamp
    torch.set_num_threads(threads)
    n_s = min(n_s, wl["n"])
    sc = {k: v[:n_s].clone() for k, v in _scene_for(wl).items()}
    cams, intr = synth.make_cameras(wl["cams"]) if wl["cams"] else ([], [])
    cams, intr = cams[:cams_s], intr[:cams_s]
    # keep the work per emitted point of the full workload: Gaussians x cameras / points is preserved
    frac = (n_s / wl["n"]) * ((len(cams) / wl["cams"]) if wl["cams"] else 1.0)
    points = max(200, int(round(wl["points"] * frac)))
    t0 = time.perf_counter()
    cov = og.build_covariance(sc["scales"], sc["rots"])
    nrm = og.calculate_normals(sc["scales"], sc["rots"])
    contrib = sc["opacities"]
    colours = sc["colours"] * 255
    keep = torch.ones(n_s, dtype=torch.bool)
    if wl["colours"] and cams:
        O = orr.PythonRendererOracle(sc["xyz"], sc["opacities"], sc["colours"], cov, shs=sc["shs"] if wl["sh"] > 0 else None,
                                     sh_degree=wl["sh"], dense=True)
        for c2w, k in zip(cams, intr):
            O(orr.Camera(c2w, k, colour_resolution=wl["res"]))
        colours = torch.as_tensor(O.get_gaussian_colours())
        mc = torch.as_tensor(O.gaussian_max_contribution)
        keep = mc > 0.05
        contrib = mc
    if int(keep.sum()) < 2:
        keep[:] = True
    cov_k, vkeep = og.validate_covariances(cov[keep])
    mags = og.gaussian_magnitudes(cov_k, contrib[keep])
    o = osamp.generate_pointcloud(sc["xyz"][keep], cov_k, colours[keep], nrm[keep], mags, points, std=2.0,
                                  num_sample_attempts=5, seed=42)
    dt = time.perf_counter() - t0
    return o["points"].shape[0], dt, dict(gaussians=n_s, cameras=len(cams), points_requested=points)


def host_threads():
    """Threads for the CPU arm: all cores up to 32 (beyond that the torch-CPU ops of this workload — thousands of
    small tile tensors — get slower, not faster: measured 240 s at 128 threads vs seconds at 8-32)."""
    return min(os.cpu_count() or 1, 32)


def cpu_baseline(wl, n_s, cams_s):
    threads = host_threads()
    npts, dt, desc = cpu_sample_run(wl, n_s, cams_s, threads)
    return {"value": round(npts / dt / 1e6, 5), "unit": UNIT, "cores": threads, "kind": "port",
            "seconds": round(dt, 2),
            "sample": f"first {desc['gaussians']} Gaussians, first {desc['cameras']} cameras at full resolution, "
                      f"{desc['points_requested']} points requested (oracle port, torch-CPU dense tile blend)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    threads = host_threads()
    for _ in range(min(args.warmup, 1)):
        cpu_sample_run(wl, max(2000, args.cpu_sample_gaussians // 10), 1, threads)
    tot_pts, tot_t = 0, 0.0
    desc = None
    for _ in range(args.steps):
        npts, dt, desc = cpu_sample_run(wl, args.cpu_sample_gaussians, args.cpu_sample_cams, threads)
        tot_pts += npts
        tot_t += dt
    value = tot_pts / tot_t / 1e6
    sample = (f"first {desc['gaussians']} Gaussians, first {desc['cameras']} cameras at full resolution, "
              f"{desc['points_requested']} points requested per step (oracle port of the reference's python path: the Python "
              "reference cannot travel to the GPU box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 5), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(tot_t / args.steps * 1e3, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} (bounded sample)", "sample": sample},
        "cpu_baseline": {"value": round(value, 5), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(a)

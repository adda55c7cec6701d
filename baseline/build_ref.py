#!/usr/bin/env python
"""Stage the UNMODIFIED reference into baseline/_ref/ so that it can be timed on the GPU box (which has no
/root/reference), per BASELINE.md §3.4 / §3.5:

    python baseline/build_ref.py          (run in the build container; /root/reference must be present)

  baseline/_ref/py/                      the reference's own Python modules, byte-for-byte (gauss_to_pc.py, ...)
  baseline/_ref/gaussian_pointcloud_rasterization/
        __init__.py                      the reference wrapper, byte-for-byte
        _C.cpython-*.so                  the reference CUDA rasterizer built for sm_100 (the throughput comparator)

baseline/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels to the
GPU box with the snapshot.  Nothing under baseline/_ref is imported by the product; only bench.py's reference legs
(`--impl reference`, the `ref_cuda` leg) and tests that compare against the reference on the box use it.

Build recipe of the extension (no source edit): copy gaussian-pointcloud-rasterization/ to a scratch dir and run its
own setup.py with  NVCC_APPEND_FLAGS="-include cstdint"  (rasterizer_impl.h:24,40-61 use std::uintptr_t / uint32_t
without <cstdint>; GCC 13 rejects that) and TORCH_CUDA_ARCH_LIST=10.0.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("G2PC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PY_FILES = ["gauss_to_pc.py", "gauss_handler.py", "gauss_render.py", "camera_handler.py", "gauss_dataloader.py",
            "transform_dataloader.py", "mask_dataloader.py", "mesh_handler.py"]
GPR = "gaussian-pointcloud-rasterization"


def _tree_hash(root):
    h = hashlib.sha256()
    for dp, dn, fn in sorted(os.walk(root)):
        if "third_party" in dp:
            continue
        for f in sorted(fn):
            if f.endswith((".cu", ".h", ".cpp", ".py")):
                h.update(f.encode())
                h.update(open(os.path.join(dp, f), "rb").read())
    return h.hexdigest()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "gauss_to_pc.py"))


def stage_python():
    dst = os.path.join(OUT, "py")
    os.makedirs(dst, exist_ok=True)
    for f in PY_FILES:
        shutil.copyfile(os.path.join(REF_ROOT, f), os.path.join(dst, f))
    pkg = os.path.join(OUT, "gaussian_pointcloud_rasterization")
    os.makedirs(pkg, exist_ok=True)
    shutil.copyfile(os.path.join(REF_ROOT, GPR, "gaussian_pointcloud_rasterization", "__init__.py"),
                    os.path.join(pkg, "__init__.py"))


def build_extension(verbose=False):
    """Build the reference's CUDA extension for sm_100 with its own setup.py in a scratch copy."""
    pkg = os.path.join(OUT, "gaussian_pointcloud_rasterization")
    stamp = os.path.join(pkg, ".src_hash")
    want = _tree_hash(os.path.join(REF_ROOT, GPR))
    if glob.glob(os.path.join(pkg, "_C*.so")) and os.path.exists(stamp) and open(stamp).read() == want:
        return glob.glob(os.path.join(pkg, "_C*.so"))[0]
    tmp = tempfile.mkdtemp(prefix="gpr_build_")
    src = os.path.join(tmp, "gpr")
    shutil.copytree(os.path.join(REF_ROOT, GPR), src)
    env = dict(os.environ)
    env["NVCC_APPEND_FLAGS"] = (env.get("NVCC_APPEND_FLAGS", "") + " -include cstdint").strip()
    env["TORCH_CUDA_ARCH_LIST"] = "10.0"
    env.setdefault("MAX_JOBS", "8")
    r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=src, env=env,
                       stdout=None if verbose else subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("reference extension build failed:\n" + (r.stdout.decode()[-4000:] if r.stdout else ""))
    so = glob.glob(os.path.join(src, "gaussian_pointcloud_rasterization", "_C*.so"))
    if not so:
        raise RuntimeError("reference extension build produced no _C*.so")
    os.makedirs(pkg, exist_ok=True)
    for old in glob.glob(os.path.join(pkg, "_C*.so")):
        os.remove(old)
    dst = os.path.join(pkg, os.path.basename(so[0]))
    shutil.copyfile(so[0], dst)
    open(stamp, "w").write(want)
    shutil.rmtree(tmp, ignore_errors=True)
    return dst


def build(verbose=False, extension=True):
    if not available():
        return None
    stage_python()
    return build_extension(verbose) if extension else None


if __name__ == "__main__":
    if not available():
        raise SystemExit(f"{REF_ROOT} not present: nothing to stage (the GPU box uses the prebuilt baseline/_ref)")
    print(build(verbose="-v" in sys.argv))

#!/usr/bin/env python
"""Recipe: compile the reference's own op sources, unmodified, into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package `de6d_b200`; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may load what this script builds.

The sources are compiled *where they lie* under /root/reference (never copied
into this repository); only object files and the three extension modules land
in oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the prebuilt modules travel
to the GPU box where /root/reference does not exist).

Modules produced (names = what the reference Python imports):
  pointnet2_batch_cuda   <- core/pcdet/ops/pointnet2/pointnet2_batch/src/*.{cpp,cu}
  iou3d_nms_cuda         <- core/pcdet/ops/iou3d_nms/src/*.{cpp,cu}
  roiaware_pool3d_cuda   <- core/pcdet/ops/roiaware_pool3d/src/*.{cpp,cu}

The only incompatibility with torch 2.11 is `#include <THC/THC.h>`; a two-line
shim header generated into oracle/_ref/shim/THC/THC.h is put first on the include
path.  The reference ships no build file, so flags are the
torch.utils.cpp_extension defaults: nvcc -O2 (fmad on), g++ -O2.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("DE6D_REFERENCE", "/root/reference")
OPS = os.path.join(REF, "core", "pcdet", "ops")

MODULES = {
    "pointnet2_batch_cuda": os.path.join(OPS, "pointnet2", "pointnet2_batch", "src"),
    "iou3d_nms_cuda": os.path.join(OPS, "iou3d_nms", "src"),
    "roiaware_pool3d_cuda": os.path.join(OPS, "roiaware_pool3d", "src"),
}


def _torch_paths():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    return ce.include_paths(), ce.library_paths()[0]


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise SystemExit("oracle/_ref build failed")
    return r.stdout


# The reference's own Python on top of the extension modules (what "drop-in" has to run unchanged): staged verbatim into
# oracle/_ref/py/ (git-ignored like the .so files, travels to the GPU box) so tests/test_dropin_gpu.py can execute it
# once over de6d_b200.compat and once over oracle/_ref/*.so.  No __init__.py is staged: oracle/ref_py.py builds the
# package tree by hand so that pcdet/__init__.py (needs the generated version.py) never runs.
PY_FILES = [
    "pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py",
    "pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py",
    "pcdet/ops/pointnet2/pointnet2_stack/pointnet2_utils.py",     # imported by pointnet2_modules.py:7
    "pcdet/ops/iou3d_nms/iou3d_nms_utils.py",
    "pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py",
    "pcdet/ops/pointnet2/pointnet2_stack/pointnet2_modules.py",   # imported by pointnet2_backbone.py:5
    "pcdet/models/model_utils/model_nms_utils.py",
    "pcdet/models/backbones_3d/pointnet2_backbone.py",            # PointNet2FSMSG: the SASA / 3DSSD / Det6D backbone
    "pcdet/utils/common_utils.py",
    "pcdet/utils/box_utils.py",
]
PY_OUT = os.path.join(OUT, "py")


def stage_python(force=False):
    """Copy the reference wrapper files, byte for byte, into oracle/_ref/py/ (only where /root/reference exists)."""
    import shutil
    core = os.path.join(REF, "core")
    if not os.path.isdir(core):
        return python_available()
    for rel in PY_FILES:
        dst = os.path.join(PY_OUT, rel)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(os.path.join(core, rel)):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(core, rel), dst)
    return True


def python_available():
    return all(os.path.exists(os.path.join(PY_OUT, rel)) for rel in PY_FILES)


def available():
    return all(os.path.exists(os.path.join(OUT, m + ".so")) for m in MODULES)


def build(force=False, jobs=None):
    if not os.path.isdir(OPS):
        return False  # GPU box: only the prebuilt modules exist
    stage_python(force)
    if available() and not force:
        return True
    os.makedirs(os.path.join(OUT, "shim", "THC"), exist_ok=True)
    with open(os.path.join(OUT, "shim", "THC", "THC.h"), "w") as f:
        f.write("#pragma once\nstruct THCState;\n")
    tinc, tlib = _torch_paths()
    pyinc = sysconfig.get_paths()["include"]
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    incs = ["-I" + os.path.join(OUT, "shim")] + ["-I" + p for p in tinc] + ["-I" + pyinc, "-I" + cuda + "/include"]
    jobs_list = []
    objs = {m: [] for m in MODULES}
    for mod, src in MODULES.items():
        odir = os.path.join(OUT, "obj", mod)
        os.makedirs(odir, exist_ok=True)
        defs = ["-DTORCH_EXTENSION_NAME=" + mod, "-DTORCH_API_INCLUDE_EXTENSION_H"]
        for fn in sorted(os.listdir(src)):
            p = os.path.join(src, fn)
            o = os.path.join(odir, fn + ".o")
            if fn.endswith(".cu"):
                cmd = ["nvcc", "-c", p, "-o", o, "-O2", "-std=c++17", "-Xcompiler", "-fPIC",
                       "-gencode", "arch=compute_100,code=sm_100", "-w",
                       "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                       "--expt-relaxed-constexpr"] + defs + incs
            elif fn.endswith(".cpp"):
                cmd = ["g++", "-c", p, "-o", o, "-O2", "-std=c++17", "-fPIC", "-w"] + defs + incs
            else:
                continue
            objs[mod].append(o)
            jobs_list.append(cmd)
    with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
        list(ex.map(_run, jobs_list))
    for mod in MODULES:
        so = os.path.join(OUT, mod + ".so")
        _run(["g++", "-shared", "-o", so] + objs[mod] +
             ["-L" + tlib, "-L" + cuda + "/lib64", "-Wl,-rpath," + tlib,
              "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"])
    return True


def load():
    """Import the three reference extension modules (torch must be imported first)."""
    import importlib.util
    import torch  # noqa: F401
    mods = {}
    for m in MODULES:
        path = os.path.join(OUT, m + ".so")
        spec = importlib.util.spec_from_file_location(m, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[m] = mod
    return mods


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference sources not present; nothing built",
          "| available:", available())

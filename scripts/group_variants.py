#!/usr/bin/env python
"""A/B of the staged grouping kernel's tuning knobs on the GPU: builds libde6d_b200 variants with -DDE6D_GS_U / -DDE6D_GS_MINB
into de6d_b200/build/variants/ (travels with gpurun, git-ignored) and times de6d_group_concat_t / de6d_group_points on the
bench shapes in a subprocess per variant (DE6D_LIB selects the library).
    python scripts/group_variants.py build      # here (nvcc, no GPU)
    python scripts/group_variants.py run        # on the GPU box
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VAR = os.path.join(ROOT, "de6d_b200", "build", "variants")
VARIANTS = [(1, 1), (1, 3), (2, 2), (2, 3), (4, 2), (4, 3), (8, 2)]


def build():
    from de6d_b200 import build as b
    b.build()
    os.makedirs(VAR, exist_ok=True)
    objs = [os.path.join(b.OBJDIR, os.path.basename(s) + ".o") for s in b.sources() if not s.endswith("group_gather.cu")]
    for u, mb in VARIANTS:
        o = os.path.join(VAR, "gg_u%d_b%d.o" % (u, mb))
        subprocess.check_call(["nvcc", "-c", os.path.join(b.CSRC, "group_gather.cu"), "-o", o, "-DDE6D_GS_U=%d" % u,
                               "-DDE6D_GS_MINB=%d" % mb] + b.NVCC_FLAGS)
        subprocess.check_call(["nvcc", "-shared", "-o", os.path.join(VAR, "lib_u%d_b%d.so" % (u, mb)), o] + objs +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print("built", len(VARIANTS), "variants in", VAR)


def time_one():
    import torch
    from de6d_b200 import pointnet2_utils as pu, synth
    B = 64
    out = []
    for (c, n, m, ns) in ((1, 16384, 4096, 32), (1, 16384, 4096, 64), (64, 4096, 1024, 32), (64, 4096, 1024, 64), (128, 1024, 512, 32),
                          (256, 512, 256, 16)):
        xyz = torch.from_numpy(synth.clouds(B, n, seed=1)).cuda()
        _, xyz_t = pu.gather_xyz(xyz, None)
        new_xyz = xyz[:, :m].contiguous()
        f = torch.randn(B, c, n, device="cuda")
        idx = torch.randint(0, n, (B, m, ns), dtype=torch.int32, device="cuda")
        for _ in range(3):
            pu.group_concat(xyz, new_xyz, f, idx, xyz_t=xyz_t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            o = pu.group_concat(xyz, new_xyz, f, idx, xyz_t=xyz_t)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        touched = min(n, m * ns)
        bytes_ = B * (4 * m * ns + 12 * touched + 12 * m + 4 * c * touched + 4 * (3 + c) * m * ns)
        out.append("%d,%d,%d,%d: %.4f ms %.0f GB/s" % (c, n, m, ns, ms, bytes_ / ms / 1e6))
    print(os.environ.get("DE6D_LIB", "default").split("/")[-1], " | ".join(out), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "one":
        time_one()
    else:
        for u, mb in VARIANTS:
            env = dict(os.environ, DE6D_LIB=os.path.join(VAR, "lib_u%d_b%d.so" % (u, mb)))
            subprocess.call([sys.executable, os.path.abspath(__file__), "one"], env=env)

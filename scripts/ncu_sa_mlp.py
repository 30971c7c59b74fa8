"""Workload for ncu: the fused SA scale kernel (csrc/sa_mlp.cu) on the SA1 / SA2 shapes of the bench, batch 64."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from de6d_b200 import pointnet2_utils as pu, sa_fused, synth  # noqa: E402

B = int(os.environ.get("DE6D_BATCH", "64"))
for n, m, c, r, ns, mlp in ((16384, 4096, 1, 0.8, 64, [32, 32, 64]), (4096, 1024, 64, 0.8, 32, [64, 64, 128]), (4096, 1024, 64, 1.6, 64, [64, 96, 128])):
    xyz = torch.from_numpy(synth.clouds(B, n, seed=1)).cuda()
    new_xyz = xyz[:, :m].contiguous()
    feats = torch.randn(B, c, n, device="cuda")
    widths = [c + 3] + mlp
    seq = []
    for a, b in zip(widths[:-1], widths[1:]):
        seq += [nn.Conv2d(a, b, 1, bias=False), nn.BatchNorm2d(b), nn.ReLU()]
    scale = sa_fused.FusedSAScale(r, ns, nn.Sequential(*seq).cuda().eval())
    for _ in range(2):
        out = scale(xyz, new_xyz, feats)
    torch.cuda.synchronize()
    print(n, m, c, ns, mlp, float(out.abs().max()), int(scale.status.item()))

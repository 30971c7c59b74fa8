import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from de6d_b200 import pointnet2_utils as pu
B = 64
for (C, N, M, ns) in ((64, 4096, 1024, 64), (64, 4096, 1024, 32), (128, 1024, 512, 32), (256, 512, 256, 32), (1, 16384, 4096, 64)):
    xyz = torch.rand(B, N, 3, device="cuda"); q = torch.rand(B, M, 3, device="cuda")
    f = torch.randn(B, C, N, device="cuda")
    for kind in ("random", "repeat3"):
        if kind == "random":
            idx = torch.randint(0, N, (B, M, ns), dtype=torch.int32, device="cuda")
        else:      # sparse balls: 3 hits repeated cyclically (what the uniform synthetic clouds produce)
            base = torch.randint(0, N, (B, M, 3), dtype=torch.int32, device="cuda")
            idx = base[:, :, torch.arange(ns, device="cuda") % 3].contiguous()
        for _ in range(4):
            pu.group_concat(xyz, q, f, idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pu.group_concat(xyz, q, f, idx)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byt = B * (4 * M * ns + 12 * min(N, M * ns) + 12 * M + 4 * C * min(N, M * ns) + 4 * (3 + C) * M * ns)
        print("%-8s G=%s waves=%s C=%d N=%d M=%d ns=%d : %.4f ms %.0f GB/s" % (kind, os.environ.get("DE6D_GS_G"), os.environ.get("DE6D_GS_WAVES"), C, N, M, ns, ms, byt / ms / 1e6), flush=True)

"""Fused F-FPS, 16 clouds (one wave of 6-CTA clusters), 4096 points x 64 channels: time against the number of samples m for the dense and the
pruned kernel -- separates the prologue (m = 1) from the per-sample cost (slope)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from de6d_b200 import synth, pointnet2_utils as pu

B = 32
big = torch.from_numpy(synth.clouds(B, 16384, 0)).cuda()
sub = pu.furthest_point_sample(big, 4096).long()
xyz = torch.gather(big, 1, sub[..., None].expand(-1, -1, 3)).contiguous()
for fscale in (1.0,):
    f = (torch.from_numpy(synth.features(B, 64, 4096, 10)).cuda() * fscale).permute(0, 2, 1)
    for prune, S in ((1, 4), (1, 44), (1, 6)):
        line = []
        for m in (1, 2, 65, 129, 257, 512):
            fn = lambda: pu.furthest_point_sample_features(xyz, f, 1.0, m, cluster_size=S, prune=prune)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            line.append("m=%d: %.4f" % (m, e0.elapsed_time(e1) / 10))
        print("fscale %.1f prune %d S %d | " % (fscale, prune, S) + "  ".join(line), flush=True)

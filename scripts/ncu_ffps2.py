"""One launch of the fused F-FPS kernel for an ncu capture: DE6D_PRUNE (0/1/2), DE6D_S (0/6/8), DE6D_FSCALE, DE6D_BATCH."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from de6d_b200 import synth, pointnet2_utils as pu
B = int(os.environ.get("DE6D_BATCH", "16"))
big = torch.from_numpy(synth.clouds(B, 16384, 0)).cuda()
sub = pu.furthest_point_sample(big, 4096).long()
xyz = torch.gather(big, 1, sub[..., None].expand(-1, -1, 3)).contiguous()
f = (torch.from_numpy(synth.features(B, 64, 4096, 10)).cuda() * float(os.environ.get("DE6D_FSCALE", "1.0"))).permute(0, 2, 1)
for _ in range(2):
    idx = pu.furthest_point_sample_features(xyz, f, 1.0, 512, cluster_size=int(os.environ.get("DE6D_S", "0")), prune=int(os.environ.get("DE6D_PRUNE", "0")))
torch.cuda.synchronize()
print(idx[0, :8])

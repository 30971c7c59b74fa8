import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth, pointnet2_utils as pu
B = int(os.environ.get("DE6D_BATCH", "16"))
xyz = torch.from_numpy(synth.clouds(B, 4096, seed=1)).cuda()
f = torch.from_numpy(synth.features(B, 64, 4096, seed=1)).cuda().permute(0, 2, 1)
for _ in range(2):
    idx = pu.furthest_point_sample_features(xyz, f, 1.0, 512)
torch.cuda.synchronize()
print(idx[0, :8])

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth, pointnet2_utils as pu
for B in (1, 4, 8, 9, 12, 14, 16, 17, 18, 20, 24, 32, 36, 64):
    xyz = torch.from_numpy(synth.clouds(B, 4096, seed=1)).cuda()
    f = torch.from_numpy(synth.features(B, 64, 4096, seed=1)).cuda().permute(0, 2, 1)
    for _ in range(2):
        pu.furthest_point_sample_features(xyz, f, 1.0, 512)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pu.furthest_point_sample_features(xyz, f, 1.0, 512)
    e1.record(); torch.cuda.synchronize()
    print("B=%2d clusters: %7.3f ms" % (B, e0.elapsed_time(e1) / 5), flush=True)

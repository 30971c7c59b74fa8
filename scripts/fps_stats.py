import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth, pointnet2_utils as pu
for maker in (synth.clouds, synth.lidar_clouds):
    for n, m in ((16384, 4096), (4096, 512), (512, 256)):
        xyz = torch.from_numpy(maker(1, n, seed=0)).cuda()
        pu.furthest_point_sample(xyz, m); torch.cuda.synchronize()

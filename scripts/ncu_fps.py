import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth, pointnet2_utils as pu
xyz = torch.from_numpy(synth.clouds(int(os.environ.get("DE6D_BATCH", "16")), 16384, seed=0)).cuda()
for _ in range(2):
    idx = pu.furthest_point_sample(xyz, 4096)
torch.cuda.synchronize()

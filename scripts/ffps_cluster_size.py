import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from de6d_b200 import synth, pointnet2_utils as pu
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8   # cluster size to pin: 6 or 8
ref = None
for B in (1, 8, 15, 16, 22, 24, 32, 64):
    xyz = torch.from_numpy(synth.clouds(B, 4096, seed=1)).cuda()
    f = torch.from_numpy(synth.features(B, 64, 4096, seed=1)).cuda().permute(0, 2, 1)
    for _ in range(2):
        out = pu.furthest_point_sample_features(xyz, f, 1.0, 512, cluster_size=S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pu.furthest_point_sample_features(xyz, f, 1.0, 512, cluster_size=S)
    e1.record(); torch.cuda.synchronize()
    print("S=%s B=%2d: %7.3f ms" % (S, B, e0.elapsed_time(e1) / 5), flush=True)
    np.save("/tmp/ff_idx_S%s_B%d.npy" % (S, B), out.cpu().numpy())

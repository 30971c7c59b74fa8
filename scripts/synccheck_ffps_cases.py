"""Two launches of the 8-CTA dense F-FPS kernel for compute-sanitizer --tool synccheck: argv[1] = 'waves' (4096 points, 18 clouds:
two waves of clusters, every CTA full) or 'empty' (3073 points, 3 clouds: one wave, the last CTAs of each cluster hold 1 / 0 points)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from de6d_b200 import synth, pointnet2_utils as pu
case = sys.argv[1]
if case == "dfps":          # no clusters: 160 one-CTA clouds on 148 SMs = two waves of plain CTAs
    xyz = torch.from_numpy(synth.clouds(160, 16384, seed=11)).cuda()
    out = pu.furthest_point_sample(xyz, 64)
    torch.cuda.synchronize()
    print(case, out[0, :6].tolist())
    sys.exit(0)
N, B, S = {"waves": (4096, 18, 8), "empty": (3073, 3, 8), "both": (3073, 18, 8), "six": (4096, 30, 6), "four": (4096, 40, 44),
           "six1": (4096, 20, 6), "four1": (4096, 30, 44)}[case]
xyz = torch.from_numpy(synth.clouds(B, N, seed=11)).cuda()
f = torch.from_numpy(synth.features(B, 64, N, seed=11)).cuda().permute(0, 2, 1)
out = pu.furthest_point_sample_features(xyz, f, 1.0, 100, cluster_size=S, prune=1)
torch.cuda.synchronize()
print(case, out[0, :6].tolist())

"""Fused F-FPS kernel forms on the chain's layer-2 shape (4096 points x 64 channels -> 512): dense vs pruned, 6- vs 8-CTA clusters,
per batch size, cloud generator and feature scale (the bound prunes by coordinates; features dominate the metric at scale 3)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from de6d_b200 import synth, pointnet2_utils as pu


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
for cloud in ("uniform", "lidar"):
    for B in (32, 64):
        gen = synth.lidar_clouds if cloud == "lidar" else synth.clouds
        big = torch.from_numpy(gen(B, 16384, 0)).cuda()
        sub = pu.furthest_point_sample(big, 4096).long()
        xyz = torch.gather(big, 1, sub[..., None].expand(-1, -1, 3)).contiguous()     # layer-2 input: D-FPS subset of the frame
        for fscale in (1.0,):
            f = (torch.from_numpy(synth.features(B, 64, 4096, 10)).cuda() * fscale).permute(0, 2, 1)
            ref = None
            for prune, S in ((1, 4), (1, 44), (1, 6), (0, 0)):
                out = pu.furthest_point_sample_features(xyz, f, 1.0, 512, cluster_size=S, prune=prune)
                ref = out if ref is None else ref
                assert torch.equal(out, ref)
                ms = timeit(lambda: pu.furthest_point_sample_features(xyz, f, 1.0, 512, cluster_size=S, prune=prune))
                rows.append({"cloud": cloud, "B": B, "fscale": fscale, "prune": prune, "S": S, "ms": round(ms, 4)})
                print(rows[-1], flush=True)
json.dump(rows, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ffps_variants.json", "w"), indent=1)

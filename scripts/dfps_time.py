"""D-FPS 16384 -> m x 64 clouds: default kernel (multi-sample rounds) vs one-sample-per-round kernel, CUDA events, mean of 10;
m = 1 / 2 isolate the prologue (Morton sort, bucket boxes) from the per-sample cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth
from de6d_b200._lib import call
B, n = 64, 16384
for name, maker in (("uniform", synth.clouds), ("lidar", synth.lidar_clouds)):
    xyz = torch.from_numpy(maker(B, n, seed=0)).cuda()
    temp = torch.empty((B, n), device="cuda")
    for m in (4096, 2):
        idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
        for impl in (0, 4, 7, 8):
            def run():
                temp.fill_(1e10)
                call("de6d_furthest_point_sampling_impl", B, n, m, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), impl,
                     torch.cuda.current_stream().cuda_stream)
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): run()
            e1.record(); torch.cuda.synchronize()
            if impl == 0: ref = idx.clone()
            else: assert torch.equal(ref, idx), "impl %d differs" % impl
            print("D-FPS %-8s m=%4d impl=%d : %7.3f ms" % (name, m, impl, e0.elapsed_time(e1) / 10), flush=True)

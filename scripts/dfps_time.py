"""D-FPS 16384 -> 4096 x 64 clouds: default kernel vs one-sample-per-round kernel (CUDA events, mean of 10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth
from de6d_b200._lib import call
B, n, m = 64, 16384, 4096
for name, maker in (("uniform", synth.clouds), ("lidar", synth.lidar_clouds)):
    xyz = torch.from_numpy(maker(B, n, seed=0)).cuda()
    temp = torch.empty((B, n), device="cuda"); idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    for impl in (0, 4):
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
        print("D-FPS %-8s impl=%d : %7.3f ms" % (name, impl, e0.elapsed_time(e1) / 10), flush=True)

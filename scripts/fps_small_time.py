"""Small-cloud FPS: automatic kernel choice (impl 0) vs the bucket kernel (impl 4), D-FPS and S-FPS, 64 clouds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from de6d_b200 import synth
from de6d_b200._lib import call
B = 64
print("| kind | N -> M | automatic (ms) | bucket kernel (ms) |\n|---|---|---|---|")
for n, m in ((64, 32), (512, 256), (1000, 256), (1024, 256), (2048, 512), (3000, 512), (4096, 512)):
    xyz = torch.from_numpy(synth.clouds(B, n, seed=0)).cuda()
    w = torch.from_numpy(synth.weights(B, n, seed=1)).cuda()
    temp = torch.empty((B, n), device="cuda"); idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    for kind in ("D", "S"):
        res = {}
        for impl in (0, 4):
            def run():
                temp.fill_(1e10)
                s = torch.cuda.current_stream().cuda_stream
                if kind == "D":
                    call("de6d_furthest_point_sampling_impl", B, n, m, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), impl, s)
                else:
                    call("de6d_furthest_point_sampling_weights_impl", B, n, m, xyz.data_ptr(), w.data_ptr(), temp.data_ptr(),
                         idx.data_ptr(), impl, s)
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): run()
            e1.record(); torch.cuda.synchronize()
            res[impl] = (e0.elapsed_time(e1) / 20, idx.clone())
        assert torch.equal(res[0][1], res[4][1])
        print("| %s-FPS | %d -> %d | %.4f | %.4f |" % (kind, n, m, res[0][0], res[4][0]), flush=True)

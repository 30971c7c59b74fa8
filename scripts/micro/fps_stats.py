"""Tuning aid: D-FPS multi-sample rounds -- rounds per cloud, samples accepted per round and why a round stopped, for 4 / 6 / 8
candidates per round.  Uses a -DDE6D_FPS_STATS build of fps.cu (scripts/micro/build_fps_stats.sh), not the product library."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from de6d_b200 import synth
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfps_stats.so"))
B, n, m = 16, 16384, 4096
for name, maker in (("uniform", synth.clouds), ("lidar", synth.lidar_clouds)):
    xyz = torch.from_numpy(maker(B, n, seed=0)).cuda()
    temp = torch.empty((B, n), device="cuda"); idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    for impl, K in ((0, 4), (7, 6), (8, 8)):
        out = (C.c_ulonglong * 8)()
        lib.de6d_fps_stats_read(out, 1)
        temp.fill_(1e10)
        rc = lib.de6d_furthest_point_sampling_impl(B, n, m, C.c_void_p(xyz.data_ptr()), C.c_void_p(temp.data_ptr()), C.c_void_p(idx.data_ptr()), impl, C.c_void_p(0))
        assert rc == 0
        lib.de6d_fps_stats_read(out, 1)
        r, acc, bnd, pair, lim, zero, full = [int(out[i]) for i in range(7)]
        print("%-8s K=%d: rounds/cloud %.0f  accepted/round %.2f  full rounds %.1f%%  stopped by: hidden-bound %.1f%%  pair-test %.1f%%  limit %.2f%%  zero %.2f%%" % (
            name, K, r / B, acc / max(r, 1), 100.0 * full / max(r, 1), 100.0 * bnd / max(r, 1), 100.0 * pair / max(r, 1), 100.0 * lim / max(r, 1), 100.0 * zero / max(r, 1)), flush=True)

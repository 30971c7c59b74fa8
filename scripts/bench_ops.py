"""Per-op A/B timings (CUDA events, mean of reps after warm-up) for kernel variants.  Development aid; the
numbers the judge reads come from bench.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from de6d_b200 import synth  # noqa: E402
from de6d_b200._lib import call  # noqa: E402
from de6d_b200 import pointnet2_utils as pu  # noqa: E402

B = int(os.environ.get("DE6D_BATCH", "64"))


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


s = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
for maker_name, maker in (("uniform", synth.clouds), ("lidar", synth.lidar_clouds)):
    for n, m in ((16384, 4096), (4096, 512), (512, 256)):
        xyz = cu(maker(B, n, seed=0))
        temp = torch.empty((B, n), device="cuda")
        idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
        res = {}
        for impl in (0, 4, 1):
            def run():
                temp.fill_(1e10)
                call("de6d_furthest_point_sampling_impl", B, n, m, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), impl, s())
            t = timeit(run)
            res[impl] = idx.clone()
            print("D-FPS %-8s B=%d n=%5d m=%4d impl=%d : %8.3f ms" % (maker_name, B, n, m, impl, t), flush=True)
        assert torch.equal(res[0], res[4]) and torch.equal(res[0], res[1])

xyz = cu(synth.clouds(B, 4096, seed=1))
f = cu(synth.features(B, 64, 4096, seed=1)).permute(0, 2, 1)
print("dist_matrix B=%d n=4096 c=64 : %8.3f ms" % (B, timeit(lambda: pu.calc_dist_matrix_for_sampling(xyz, f, 1.0))))
mat = pu.calc_dist_matrix_for_sampling(xyz, f, 1.0)
print("fps_matrix  B=%d n=4096 m=512 : %8.3f ms" % (B, timeit(lambda: pu.furthest_point_sample_matrix(mat, 512))))

# ---- the reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100) on the same inputs: context for
# DESIGN.md, not a bench value.  Only runs where the prebuilt modules are present.
try:
    from oracle import build_ref
    ref = build_ref.load() if build_ref.available() else None
except Exception as e:  # noqa: BLE001
    ref = None
    print("reference modules not loadable:", e)
if ref is not None:
    p2, iou3d = ref["pointnet2_batch_cuda"], ref["iou3d_nms_cuda"]
    from de6d_b200.compat import pointnet2_batch_cuda as mine
    from de6d_b200 import iou3d_nms_utils as iu
    xyz = cu(synth.clouds(B, 16384, seed=0))
    temp = torch.empty((B, 16384), device="cuda"); idx = torch.empty((B, 4096), dtype=torch.int32, device="cuda")

    def ref_fps():
        temp.fill_(1e10)
        p2.farthest_point_sampling_wrapper(B, 16384, 4096, xyz, temp, idx)
    print("REF  D-FPS B=%d 16384->4096 : %8.3f ms" % (B, timeit(ref_fps, reps=3, warm=1)))
    ref_idx = idx.clone()

    def my_fps():
        temp.fill_(1e10)
        mine.farthest_point_sampling_wrapper(B, 16384, 4096, xyz, temp, idx)
    print("OURS D-FPS B=%d 16384->4096 : %8.3f ms  equal=%s" % (B, timeit(my_fps), torch.equal(idx, ref_idx)))
    q = pu.gather_operation(xyz.transpose(1, 2).contiguous(), ref_idx).transpose(1, 2).contiguous()
    for r, ns in ((0.2, 32), (0.8, 64)):
        bi = torch.zeros((B, 4096, ns), dtype=torch.int32, device="cuda"); bc = torch.zeros((B, 4096), dtype=torch.int32, device="cuda")
        t_ref = timeit(lambda: p2.ball_query_cnt_wrapper(B, 16384, 4096, r, ns, q, xyz, bc, bi), reps=3, warm=1)
        ri, rc = bi.clone(), bc.clone()
        bi.zero_(); bc.zero_()
        t_my = timeit(lambda: mine.ball_query_cnt_wrapper(B, 16384, 4096, r, ns, q, xyz, bc, bi))
        print("ball_query_cnt r=%.1f ns=%d : REF %8.3f ms  OURS %8.3f ms  equal=%s" % (r, ns, t_ref, t_my, torch.equal(bi, ri) and torch.equal(bc, rc)))
    f = cu(synth.features(B, 64, 4096, seed=2)); gi = torch.randint(0, 4096, (B, 1024, 32), dtype=torch.int32, device="cuda")
    go = torch.empty((B, 64, 1024, 32), device="cuda")
    t_ref = timeit(lambda: p2.group_points_wrapper(B, 64, 4096, 1024, 32, f, gi, go), reps=5)
    g_ref = go.clone()
    t_my = timeit(lambda: mine.group_points_wrapper(B, 64, 4096, 1024, 32, f, gi, go))
    print("group_points C=64 N=4096 M=1024 ns=32 : REF %8.3f ms  OURS %8.3f ms  equal=%s" % (t_ref, t_my, torch.equal(go, g_ref)))
    bx, sc = synth.proposals(B, 512, seed=0)
    bx, sc = cu(bx), cu(sc)

    def ref_nms():  # what detector3d_template.post_processing does: one call (malloc + D2H + host sweep) per frame
        for fr in range(B):
            order = sc[fr].sort(0, descending=True)[1]
            keep = torch.empty(512, dtype=torch.int64)
            n = iou3d.nms_gpu(bx[fr][order].contiguous(), keep, 0.01)
            _ = order[keep[:n].cuda()]
    print("NMS 512 boxes x %d frames : REF per-frame loop %8.3f ms  OURS batched %8.3f ms" % (
        B, timeit(ref_nms, reps=3, warm=1), timeit(lambda: iu.nms_gpu_batched(bx, sc, 0.01))))

# ---- BASELINE configs[4]: 131072-point stress frames (FPS to 16384, ball_query ns = 64)
Bs = 4
xyz = cu(synth.lidar_clouds(Bs, 131072, seed=4))
temp = torch.empty((Bs, 131072), device="cuda"); idx = torch.empty((Bs, 16384), dtype=torch.int32, device="cuda")


def big_fps():
    temp.fill_(1e10)
    call("de6d_furthest_point_sampling", Bs, 131072, 16384, xyz.data_ptr(), temp.data_ptr(), idx.data_ptr(), s())
print("stress D-FPS B=%d 131072->16384 : %8.3f ms" % (Bs, timeit(big_fps, reps=2, warm=1)))
q = pu.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
print("stress ball_query_cnt r=0.2 ns=64 M=16384 : %8.3f ms" % timeit(lambda: pu.ball_query_cnt(0.2, 64, xyz, q), reps=3, warm=1))
xyz = cu(synth.clouds(B, 4096, seed=1))
f = cu(synth.features(B, 64, 4096, seed=1)).permute(0, 2, 1)
print("fused F-FPS B=%d n=4096 c=64 m=512 : %8.3f ms" % (B, timeit(lambda: pu.furthest_point_sample_features(xyz, f, 1.0, 512))))
# S-FPS at the layer-1 size (not on the SASA chain, where S-FPS samples 512 points): registers hold weights too
xyz = cu(synth.clouds(B, 16384, seed=0)); wts = cu(synth.weights(B, 16384, seed=1))
print("S-FPS B=%d 16384->4096 : %8.3f ms" % (B, timeit(lambda: pu.furthest_point_sample_weights(xyz, wts, 4096), reps=5)))
xyz = cu(synth.clouds(B, 4096, seed=0)); wts = cu(synth.weights(B, 4096, seed=1))
print("S-FPS B=%d 4096->512 : %8.3f ms" % (B, timeit(lambda: pu.furthest_point_sample_weights(xyz, wts, 512), reps=5)))

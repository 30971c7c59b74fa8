#!/usr/bin/env python
"""Every row of SURVEY.md section 8(a) measured on one B200: this library's kernel (CUDA events), the reference's own CUDA
kernel recompiled for sm_100 where oracle/_ref is present (context only), the CPU oracle on one host core (bounded
sample), algorithmic bytes and the resulting GB/s.  Writes a markdown table (stdout).
    python scripts/ops_table.py > gpurun_out/ops_table.md"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from de6d_b200 import synth, pointnet2_utils as pu, iou3d_nms_utils as iu, roiaware_pool3d_utils as ru  # noqa: E402
from de6d_b200.compat import pointnet2_batch_cuda as mine  # noqa: E402
from oracle import oracle as orc  # noqa: E402

try:
    from oracle import build_ref
    ref = build_ref.load() if build_ref.available() else None
except Exception:  # noqa: BLE001
    ref = None
p2 = ref["pointnet2_batch_cuda"] if ref else None
iou3d = ref["iou3d_nms_cuda"] if ref else None
roi = ref["roiaware_pool3d_cuda"] if ref else None
PEAK = 6551.0
B = 16


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def gpu_ms(fn, reps=10, warm=2):
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


def cpu_ms(fn):
    t0 = time.perf_counter()
    fn()
    return 1e3 * (time.perf_counter() - t0)


rows = []


def row(name, shape, ours, refk, cpu, cpu_scale, byt):
    rows.append((name, shape, ours, refk, cpu * cpu_scale if cpu is not None else None, byt))


xyz_h = synth.clouds(B, 16384, seed=0)
xyz = cu(xyz_h)
temp = torch.empty((B, 16384), device="cuda"); idx = torch.empty((B, 4096), dtype=torch.int32, device="cuda")


def fps_mine():
    temp.fill_(1e10); mine.farthest_point_sampling_wrapper(B, 16384, 4096, xyz, temp, idx)


def fps_ref():
    temp.fill_(1e10); p2.farthest_point_sampling_wrapper(B, 16384, 4096, xyz, temp, idx)


row("a1 D-FPS", "B=16, 16384->4096", gpu_ms(fps_mine), gpu_ms(fps_ref, 3, 1) if p2 else None,
    cpu_ms(lambda: orc.furthest_point_sample(xyz_h[:1], 4096)), B, B * (12 * 16384 + 4 * 4096))
sidx = pu.furthest_point_sample(xyz, 4096)
x2_h = synth.clouds(B, 4096, seed=1); x2 = cu(x2_h)
f2_h = synth.features(B, 64, 4096, seed=1); f2 = cu(f2_h)
mat = pu.calc_dist_matrix_for_sampling(x2, f2.permute(0, 2, 1), 1.0)
t2 = torch.empty((B, 4096), device="cuda"); i2 = torch.empty((B, 512), dtype=torch.int32, device="cuda")


def ffps_mine():
    t2.fill_(1e10); mine.furthest_point_sampling_matrix_wrapper(B, 4096, 512, mat, t2, i2)


def ffps_ref():
    t2.fill_(1e10); p2.furthest_point_sampling_matrix_wrapper(B, 4096, 512, mat, t2, i2)


mat_h = mat[:1].cpu().numpy()
row("a2 F-FPS (matrix)", "B=16, 4096->512", gpu_ms(ffps_mine), gpu_ms(ffps_ref, 3, 1) if p2 else None,
    cpu_ms(lambda: orc.furthest_point_sample_matrix(mat_h, 512)), B, B * (4 * 4096 * 512 + 4 * 512))
row("a2' dist matrix + F-FPS fused", "B=16, 4096 pts, 64 ch ->512", gpu_ms(lambda: pu.furthest_point_sample_features(x2, f2.permute(0, 2, 1), 1.0, 512)),
    None, None, 1, B * (12 * 4096 + 4 * 4096 * 64 + 4 * 512))
w_h = synth.weights(B, 4096, seed=2); wts = cu(w_h)
row("a3 S-FPS", "B=16, 4096->512", gpu_ms(lambda: pu.furthest_point_sample_weights(x2, wts, 512)),
    gpu_ms(lambda: p2.furthest_point_sampling_weights_wrapper(B, 4096, 512, x2, wts, t2.fill_(1e10), i2), 3, 1) if p2 else None,
    cpu_ms(lambda: orc.furthest_point_sample_weights(x2_h[:1], w_h[:1], 512)), B, B * (16 * 4096 + 4 * 512))
xt = xyz.transpose(1, 2).contiguous()
g_out = torch.empty((B, 3, 4096), device="cuda")
row("a4 gather_points", "B=16, C=3, 16384->4096", gpu_ms(lambda: mine.gather_points_wrapper(B, 3, 16384, 4096, xt, sidx, g_out)),
    gpu_ms(lambda: p2.gather_points_wrapper(B, 3, 16384, 4096, xt, sidx, g_out)) if p2 else None,
    cpu_ms(lambda: orc.gather_operation(xt[:1].cpu().numpy(), sidx[:1].cpu().numpy())), B, B * (4 * 4096 + 8 * 3 * 4096))
q = pu.gather_operation(xt, sidx).transpose(1, 2).contiguous()
q_h = q.cpu().numpy()
for tag, fn_m, fn_r, fn_c in (
        ("a5 ball_query", lambda bi, bc: mine.ball_query_wrapper(B, 16384, 4096, 0.4, 32, q, xyz, bi),
         lambda bi, bc: p2.ball_query_wrapper(B, 16384, 4096, 0.4, 32, q, xyz, bi), lambda: orc.ball_query(0.4, 32, xyz_h[:1], q_h[:1])),
        ("a6 ball_query_cnt", lambda bi, bc: mine.ball_query_cnt_wrapper(B, 16384, 4096, 0.4, 32, q, xyz, bc, bi),
         lambda bi, bc: p2.ball_query_cnt_wrapper(B, 16384, 4096, 0.4, 32, q, xyz, bc, bi), lambda: orc.ball_query_cnt(0.4, 32, xyz_h[:1], q_h[:1])),
        ("a7 ball_query_dilated", lambda bi, bc: mine.ball_query_dilated_wrapper(B, 16384, 4096, 0.2, 0.4, 32, q, xyz, bc, bi),
         lambda bi, bc: p2.ball_query_dilated_wrapper(B, 16384, 4096, 0.2, 0.4, 32, q, xyz, bc, bi), lambda: orc.ball_query_dilated(0.2, 0.4, 32, xyz_h[:1], q_h[:1]))):
    bi = torch.zeros((B, 4096, 32), dtype=torch.int32, device="cuda"); bc = torch.zeros((B, 4096), dtype=torch.int32, device="cuda")
    row(tag, "B=16, N=16384, M=4096, r=0.4, ns=32", gpu_ms(lambda: fn_m(bi, bc)), gpu_ms(lambda: fn_r(bi, bc), 3, 1) if p2 else None,
        cpu_ms(fn_c), B, B * (12 * 16384 + 12 * 4096 + 4 * 4096 * 32 + 4 * 4096))
gi = torch.randint(0, 4096, (B, 1024, 32), dtype=torch.int32, device="cuda")
go = torch.empty((B, 64, 1024, 32), device="cuda")
row("a8 group_points", "B=16, C=64, N=4096, M=1024, ns=32", gpu_ms(lambda: mine.group_points_wrapper(B, 64, 4096, 1024, 32, f2, gi, go)),
    gpu_ms(lambda: p2.group_points_wrapper(B, 64, 4096, 1024, 32, f2, gi, go)) if p2 else None,
    cpu_ms(lambda: orc.grouping_operation(f2_h[:1], gi[:1].cpu().numpy())), B, B * (4 * 1024 * 32 + 4 * 64 * 4096 + 4 * 64 * 1024 * 32))
unk_h = synth.clouds(B, 16384, seed=3); kn_h = synth.clouds(B, 4096, seed=4)
unk, kn = cu(unk_h), cu(kn_h)
d2 = torch.empty((B, 16384, 3), device="cuda"); i3 = torch.empty((B, 16384, 3), dtype=torch.int32, device="cuda")
row("a9 three_nn", "B=16, n=16384, m=4096", gpu_ms(lambda: mine.three_nn_wrapper(B, 16384, 4096, unk, kn, d2, i3)),
    gpu_ms(lambda: p2.three_nn_wrapper(B, 16384, 4096, unk, kn, d2, i3), 3, 1) if p2 else None,
    cpu_ms(lambda: orc.three_nn(unk_h[:1, :2048], kn_h[:1])), B * 8, B * (36 * 16384 + 12 * 4096))
wgt = torch.rand((B, 16384, 3), device="cuda"); fo = torch.empty((B, 64, 16384), device="cuda")
row("a10 three_interpolate", "B=16, C=64, m=4096, n=16384", gpu_ms(lambda: mine.three_interpolate_wrapper(B, 64, 4096, 16384, f2, i3, wgt, fo)),
    gpu_ms(lambda: p2.three_interpolate_wrapper(B, 64, 4096, 16384, f2, i3, wgt, fo)) if p2 else None,
    cpu_ms(lambda: orc.three_interpolate(f2_h[:1], i3[:1].cpu().numpy(), wgt[:1].cpu().numpy())), B, B * (24 * 16384 + 4 * 64 * 4096 + 4 * 64 * 16384))
bx_h, sc_h = synth.proposals(1, 512, seed=0)
bx, sc = cu(bx_h[0]), cu(sc_h[0])
o = torch.zeros((512, 512), device="cuda")
row("a11 boxes_iou_bev", "512 x 512", gpu_ms(lambda: iu.boxes_iou_bev(bx, bx)),
    gpu_ms(lambda: iou3d.boxes_iou_bev_gpu(bx, bx, o)) if iou3d else None, cpu_ms(lambda: orc.boxes_iou_bev(bx_h[0], bx_h[0])), 1, 28 * 1024 + 4 * 512 * 512)
row("a12 boxes_iou3d_gpu", "512 x 512 (reference: 7-kernel torch composition, not timed)", gpu_ms(lambda: iu.boxes_iou3d_gpu(bx, bx)), None,
    cpu_ms(lambda: orc.boxes_iou3d(bx_h[0], bx_h[0])), 1, 28 * 1024 + 4 * 512 * 512)


def ref_nms():
    order = sc.sort(0, descending=True)[1]
    keep = torch.empty(512, dtype=torch.int64)
    n = iou3d.nms_gpu(bx[order].contiguous(), keep, 0.01)
    return order[keep[:n].cuda()]


row("a13 nms_gpu (one frame, reference API incl. its host sync)", "512 boxes", gpu_ms(lambda: iu.nms_gpu(bx, sc, 0.01)),
    gpu_ms(ref_nms, 5, 2) if iou3d else None, cpu_ms(lambda: orc.nms_gpu(bx_h[0], sc_h[0], 0.01)), 1, 36 * 512)
bxb_h, scb_h = synth.proposals(64, 512, seed=1)
bxb, scb = cu(bxb_h), cu(scb_h)
op = iu.BatchedNMS(64, 512)
row("a13' nms batched, sync-free", "64 frames x 512 boxes", gpu_ms(lambda: op(bxb, scb, 0.01)), None, None, 1, 64 * 36 * 512)
row("a14 nms_normal_gpu (one frame)", "512 boxes", gpu_ms(lambda: iu.nms_normal_gpu(bx, sc, 0.01)), None,
    cpu_ms(lambda: orc.nms_gpu(bx_h[0], sc_h[0], 0.01, normal=True)), 1, 36 * 512)
b100_h = synth.boxes(1, 100, seed=5)[0]
row("a15 boxes_bev_iou_cpu (host tensors in/out)", "100 x 100", gpu_ms(lambda: iu.boxes_bev_iou_cpu(b100_h, b100_h)), None,
    cpu_ms(lambda: orc.boxes_bev_iou_cpu(b100_h, b100_h)), 1, 28 * 200 + 4 * 100 * 100)
pb_h = synth.boxes(B, 100, seed=6); pb = cu(pb_h)
po = torch.full((B, 16384), -1, dtype=torch.int32, device="cuda")
row("a16 points_in_boxes_gpu", "B=16, 16384 pts, 100 boxes", gpu_ms(lambda: ru.points_in_boxes_gpu(xyz, pb)),
    gpu_ms(lambda: roi.points_in_boxes_gpu(pb, xyz, po)) if roi else None,
    cpu_ms(lambda: orc.points_in_boxes_gpu(xyz_h[:1], pb_h[:1])), B, B * (16 * 16384 + 28 * 100))
row("a17 points_in_boxes_cpu (host arrays in/out)", "16384 pts, 100 boxes", gpu_ms(lambda: ru.points_in_boxes_cpu(xyz_h[0], pb_h[0])), None,
    cpu_ms(lambda: orc.points_in_boxes_cpu(xyz_h[0], pb_h[0])), 1, 12 * 16384 + 28 * 100 + 4 * 100 * 16384)

print("| SURVEY 8(a) row | shape | this library (ms) | reference CUDA kernel, sm_100 build (ms) | CPU oracle, 1 core (ms, scaled to the shape) | alg. bytes | GB/s | of %g GB/s |" % PEAK)
print("|---|---|---|---|---|---|---|---|")
for name, shape, ours, refk, cpu, byt in rows:
    gbs = byt / (ours * 1e-3) / 1e9
    print("| %s | %s | %.4f | %s | %s | %.2e | %.0f | %.1f %% |" % (name, shape, ours, "%.3f" % refk if refk is not None else "-",
                                                              "%.0f" % cpu if cpu is not None else "-", byt, gbs, 100 * gbs / PEAK))
print("\nhost cores: %d; reference kernels: %s" % (len(os.sched_getaffinity(0)), "oracle/_ref loaded" if ref else "not available"))

#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference ops compiled by
oracle/build_ref.py (oracle/_ref/*.so).

    python tests/golden/make_golden.py --cpu     # reference CPU entry points; runs anywhere the modules load
    python tests/golden/make_golden.py --cuda    # reference CUDA kernels; needs a GPU (run under gpurun),
                                                 # writes gpurun_out/golden_cuda.npz to be copied here
    python tests/golden/make_golden.py --cuda2   # reference python wrappers (autograd, boxes_iou3d_gpu, groupers) over
                                                 # the reference kernels -> gpurun_out/golden_cuda2.npz

Inputs come from de6d_b200.synth with fixed seeds and are stored next to the outputs so the fixtures are
self-contained (tests never regenerate them).  Sizes are kept small: the files are committed.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from de6d_b200 import synth  # noqa: E402
from oracle import build_ref  # noqa: E402


def cpu_vectors(ref):
    out = {}
    a = synth.proposals(1, 48, seed=3, clusters=6)[0][0]
    b = synth.proposals(1, 40, seed=3, clusters=5)[0][0]
    b[:20] = a[:20] + np.random.default_rng(1).normal(0, 0.2, (20, 7)).astype(np.float32)
    iou = torch.zeros(a.shape[0], b.shape[0])
    ref["iou3d_nms_cuda"].boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), iou)
    out.update(iou_a=a, iou_b=b, iou_bev_cpu=iou.numpy())
    bx = synth.boxes(1, 12, seed=4)[0]
    rng = np.random.default_rng(4)
    pts = (bx[rng.integers(0, 12, 600), :3] + rng.normal(0, 1.0, (600, 3))).astype(np.float32)
    m = torch.zeros(12, 600, dtype=torch.int32)
    ref["roiaware_pool3d_cuda"].points_in_boxes_cpu(torch.from_numpy(bx), torch.from_numpy(pts), m)
    out.update(pib_boxes=bx, pib_pts=pts, pib_cpu=m.numpy())
    return out


def cuda_vectors(ref):
    p2, iou3d, roi = ref["pointnet2_batch_cuda"], ref["iou3d_nms_cuda"], ref["roiaware_pool3d_cuda"]
    dev = "cuda"
    out = {}

    def T(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    # ---- FPS: power-of-two and ragged sizes, with duplicated points (tie rule) ----
    for tag, (B, N, M, dup) in {"a": (2, 1000, 64, 0.1), "b": (2, 2048, 128, 0.2), "c": (1, 300, 40, 0.0),
                                "d": (1, 4096, 256, 0.05)}.items():
        xyz = synth.clouds(B, N, seed=11, dup_frac=dup)
        temp = torch.full((B, N), 1e10, device=dev)
        idx = torch.zeros((B, M), dtype=torch.int32, device=dev)
        p2.farthest_point_sampling_wrapper(B, N, M, T(xyz), temp, idx)
        out["fps_%s_xyz" % tag] = xyz
        out["fps_%s_idx" % tag] = idx.cpu().numpy()
        out["fps_%s_temp" % tag] = temp.cpu().numpy()
        w = synth.weights(B, N, seed=12)
        w[:, :7] = 0.0  # exercises the max(w, 1e-12) double path
        temp = torch.full((B, N), 1e10, device=dev)
        idx = torch.zeros((B, M), dtype=torch.int32, device=dev)
        p2.furthest_point_sampling_weights_wrapper(B, N, M, T(xyz), T(w), temp, idx)
        out["sfps_%s_w" % tag] = w
        out["sfps_%s_idx" % tag] = idx.cpu().numpy()
    xyz = synth.clouds(2, 384, seed=13, dup_frac=0.1)
    mat = synth.dist_matrix(xyz, synth.features(2, 8, 384, seed=13))
    temp = torch.full((2, 384), 1e10, device=dev)
    idx = torch.zeros((2, 96), dtype=torch.int32, device=dev)
    p2.furthest_point_sampling_matrix_wrapper(2, 384, 96, T(mat), temp, idx)
    out.update(ffps_xyz=xyz, ffps_mat=mat, ffps_idx=idx.cpu().numpy())

    # ---- ball query (three variants), group, gather ----
    B, N, M, ns = 2, 1500, 96, 16
    xyz = synth.lidar_clouds(B, N, seed=21)
    new_xyz = xyz[:, ::15][:, :M].copy() + np.random.default_rng(2).normal(0, 0.05, (B, M, 3)).astype(np.float32)
    new_xyz[:, -3:] += 500.0  # empty balls
    out.update(bq_xyz=xyz, bq_new_xyz=new_xyz)
    for r in (0.5, 2.0):
        idx = torch.zeros((B, M, ns), dtype=torch.int32, device=dev)
        p2.ball_query_wrapper(B, N, M, r, ns, T(new_xyz), T(xyz), idx)
        out["bq_idx_r%g" % r] = idx.cpu().numpy()
        idx = torch.zeros((B, M, ns), dtype=torch.int32, device=dev); cnt = torch.zeros((B, M), dtype=torch.int32, device=dev)
        p2.ball_query_cnt_wrapper(B, N, M, r, ns, T(new_xyz), T(xyz), cnt, idx)
        out["bqc_idx_r%g" % r] = idx.cpu().numpy(); out["bqc_cnt_r%g" % r] = cnt.cpu().numpy()
        idx = torch.zeros((B, M, ns), dtype=torch.int32, device=dev); cnt = torch.zeros((B, M), dtype=torch.int32, device=dev)
        p2.ball_query_dilated_wrapper(B, N, M, r * 0.5, r, ns, T(new_xyz), T(xyz), cnt, idx)
        out["bqd_idx_r%g" % r] = idx.cpu().numpy(); out["bqd_cnt_r%g" % r] = cnt.cpu().numpy()

    # ---- three_nn / three_interpolate ----
    unknown = synth.clouds(2, 200, seed=31); known = synth.clouds(2, 50, seed=32)
    known[:, 10] = known[:, 3]  # equal distances: earliest index must win
    d2 = torch.zeros((2, 200, 3), device=dev); idx = torch.zeros((2, 200, 3), dtype=torch.int32, device=dev)
    p2.three_nn_wrapper(2, 200, 50, T(unknown), T(known), d2, idx)
    feats = synth.features(2, 5, 50, seed=33)
    wgt = np.random.default_rng(3).uniform(0, 1, (2, 200, 3)).astype(np.float32)
    wgt /= wgt.sum(-1, keepdims=True)
    o = torch.zeros((2, 5, 200), device=dev)
    p2.three_interpolate_wrapper(2, 5, 50, 200, T(feats), idx, T(wgt), o)
    out.update(nn_unknown=unknown, nn_known=known, nn_dist2=d2.cpu().numpy(), nn_idx=idx.cpu().numpy(),
               ti_feats=feats, ti_weight=wgt, ti_out=o.cpu().numpy())

    # ---- rotated IoU / NMS ----
    bx, sc = synth.proposals(1, 200, seed=41, clusters=25)
    bx, sc = bx[0], sc[0]
    a, b = bx[:90], bx[60:200]
    ov = torch.zeros((90, 140), device=dev); iou = torch.zeros((90, 140), device=dev)
    iou3d.boxes_overlap_bev_gpu(T(a), T(b), ov)
    iou3d.boxes_iou_bev_gpu(T(a), T(b), iou)
    out.update(iou_gpu_a=a, iou_gpu_b=b, overlap_gpu=ov.cpu().numpy(), iou_gpu=iou.cpu().numpy())
    order = np.argsort(-sc, kind="stable")
    sorted_boxes = np.ascontiguousarray(bx[order])
    for thr in (0.01, 0.1, 0.5):
        keep = torch.zeros(200, dtype=torch.int64)
        n = iou3d.nms_gpu(T(sorted_boxes), keep, thr)
        out["nms_keep_%g" % thr] = keep[:n].numpy().copy()
        keep = torch.zeros(200, dtype=torch.int64)
        n = iou3d.nms_normal_gpu(T(sorted_boxes), keep, thr)
        out["nmsn_keep_%g" % thr] = keep[:n].numpy().copy()
    out["nms_sorted_boxes"] = sorted_boxes

    # ---- points_in_boxes_gpu ----
    boxes = synth.boxes(2, 20, seed=51)
    boxes[:, -2:] = 0.0  # zero-padded gt boxes
    rng = np.random.default_rng(5)
    pts = np.stack([(boxes[i, rng.integers(0, 18, 800), :3] + rng.normal(0, 1.0, (800, 3))).astype(np.float32) for i in range(2)])
    o = torch.full((2, 800), -1, dtype=torch.int32, device=dev)
    roi.points_in_boxes_gpu(T(boxes), T(pts), o)
    out.update(pibg_boxes=boxes, pibg_pts=pts, pibg_out=o.cpu().numpy())
    return out


def cuda_vectors_wrappers():
    """Second fixture (round 2): outputs of the reference's own PYTHON wrappers over its own kernels -- the rows the first
    fixture lacks: boxes_iou3d_gpu (iou3d_nms_utils.py:48-81), gather / group forward AND backward (autograd Functions,
    pointnet2_utils.py:115-149, 232-273), three_interpolate backward (:184-229), QueryWithCntAndGroup (:390-424)."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_py
    tree = ref_py.load_tree("pcdet_ref", build_ref.load())
    pu, iu = tree.pointnet2_utils, tree.iou3d_nms_utils
    dev = "cuda"
    out = {}

    def T(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    bx, _ = synth.proposals(1, 160, seed=61, clusters=20)
    a, b = bx[0][:80].copy(), bx[0][50:160].copy()
    b[:, 2] += np.random.default_rng(6).normal(0, 0.4, len(b)).astype(np.float32)     # partial height overlaps
    out.update(iou3d_a=a, iou3d_b=b, iou3d=iu.boxes_iou3d_gpu(T(a), T(b)).cpu().numpy())

    B, C, N, M, ns = 2, 6, 700, 90, 8
    rng = np.random.default_rng(7)
    feats = synth.features(B, C, N, seed=62)
    gidx = rng.integers(0, N, (B, M)).astype(np.int32)
    f = T(feats).requires_grad_(True)
    y = pu.gather_operation(f, T(gidx))
    gy = torch.from_numpy(rng.normal(size=(B, C, M)).astype(np.float32)).to(dev)
    y.backward(gy)
    out.update(gg_feats=feats, gather_idx=gidx, gather_out=y.detach().cpu().numpy(), gather_gout=gy.cpu().numpy(),
               gather_grad=f.grad.cpu().numpy())
    qidx = rng.integers(0, N, (B, M, ns)).astype(np.int32)
    f = T(feats).requires_grad_(True)
    y = pu.grouping_operation(f, T(qidx))
    gy = torch.from_numpy(rng.normal(size=(B, C, M, ns)).astype(np.float32)).to(dev)
    y.backward(gy)
    out.update(group_idx=qidx, group_out=y.detach().cpu().numpy(), group_gout=gy.cpu().numpy(), group_grad=f.grad.cpu().numpy())

    m_known, n_unknown = 60, 150
    kf = synth.features(B, C, m_known, seed=63)
    tidx = rng.integers(0, m_known, (B, n_unknown, 3)).astype(np.int32)
    tw = rng.uniform(0, 1, (B, n_unknown, 3)).astype(np.float32)
    tw /= tw.sum(-1, keepdims=True)
    f = T(kf).requires_grad_(True)
    y = pu.three_interpolate(f, T(tidx), T(tw))
    gy = torch.from_numpy(rng.normal(size=(B, C, n_unknown)).astype(np.float32)).to(dev)
    y.backward(gy)
    out.update(ti2_feats=kf, ti2_idx=tidx, ti2_weight=tw, ti2_out=y.detach().cpu().numpy(), ti2_gout=gy.cpu().numpy(),
               ti2_grad=f.grad.cpu().numpy())

    xyz = synth.lidar_clouds(B, N, seed=64)
    new_xyz = np.ascontiguousarray(xyz[:, ::7][:, :M]) + np.float32(0.02)
    new_xyz[:, -2:] += 300.0
    cnt, nf = pu.QueryWithCntAndGroup(1.5, ns)(T(xyz), T(new_xyz), T(feats))
    out.update(qg_xyz=xyz, qg_new_xyz=new_xyz, qg_cnt=cnt.cpu().numpy(), qg_out=nf.cpu().numpy())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cuda2", action="store_true", help="reference python wrappers over the reference kernels (needs a GPU)")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    build_ref.build()
    ref = build_ref.load()
    here = os.path.dirname(os.path.abspath(__file__))
    if args.cpu:
        path = args.out or os.path.join(here, "golden_cpu.npz")
        np.savez_compressed(path, **cpu_vectors(ref))
        print("wrote", path)
    if args.cuda2:
        path = os.path.join(ROOT, "gpurun_out", "golden_cuda2.npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **cuda_vectors_wrappers())
        print("wrote", path)
    if args.cuda:
        path = args.out or os.path.join(ROOT, "gpurun_out", "golden_cuda.npz")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez_compressed(path, **cuda_vectors(ref))
        print("wrote", path)


if __name__ == "__main__":
    main()

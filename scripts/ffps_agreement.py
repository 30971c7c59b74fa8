#!/usr/bin/env python
"""How far the fused F-FPS (de6d_furthest_point_sampling_features: direct-difference distances, no (B,N,N) matrix) is
from what the reference PIPELINE selects on the same (xyz, features) -- VERDICT r1 weak #1.

Reference pipeline (pointnet2_modules.py:383-388): calc_dist_matrix_for_sampling = torch.cdist(xyz) + gamma *
torch.cdist(features) (pointnet2_utils.py:36-44; for N > 25 torch.cdist takes the |a|^2+|b|^2-2ab GEMM expansion in
fp32) -> furthest_point_sampling_matrix_wrapper (the reference kernel from oracle/_ref).  FPS index sequences are
discontinuous in the distances, so two evaluations of the same metric in different arithmetic diverge after the first
near-tie.  This script measures, per cloud:
    * position of the first differing index, fraction of identical positions, overlap of the selected SETS,
    * coverage quality of both selections (max over points of the distance to the nearest selected point, evaluated in
      float64 with the exact metric) -- what the sampler is for,
for these arms against the float64-exact greedy F-FPS ("exact": cdist in float64, the matrix rounded once to fp32):
    ref      reference python + reference kernels (torch.cdist fp32 GEMM expansion)
    ref_tf32 the same with torch.backends.cuda.matmul.allow_tf32 = True (what an Ampere+ default-flag run of older
             torch versions computes; shows the reference's own arithmetic spread)
    ours     de6d_b200 fused kernel
    ours2    de6d_b200 two-call route (de6d_dist_matrix + matrix kernel; must equal `ours` bit for bit)
    compat   torch.cdist + de6d_furthest_point_sampling_matrix (what an unmodified checkout gets through compat:
             must equal `ref` bit for bit)

    python scripts/ffps_agreement.py [--out profiles/r2_ffps_agreement.json] [--batch 16]
"""
import argparse
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def compare(a, b):
    """a, b (B, M) index sequences -> per-cloud (first difference, identical positions, set overlap)."""
    B, M = a.shape
    first, same, inter = [], [], []
    for i in range(B):
        d = np.nonzero(a[i] != b[i])[0]
        first.append(int(d[0]) if len(d) else M)
        same.append(float((a[i] == b[i]).mean()))
        inter.append(len(set(a[i].tolist()) & set(b[i].tolist())) / M)
    return {"first_diff_min": int(min(first)), "first_diff_median": float(np.median(first)),
            "identical_sequences": int(sum(f == M for f in first)), "clouds": B,
            "same_position_mean": float(np.mean(same)), "set_overlap_mean": float(np.mean(inter)),
            "set_overlap_min": float(np.min(inter))}


def coverage(exact64, idx):
    """max_i min_{s in idx} D[i, s] per cloud (float64 metric): the quantity farthest point sampling minimises greedily."""
    out = []
    for b in range(idx.shape[0]):
        sel = torch.from_numpy(idx[b].astype(np.int64)).to(exact64.device)
        out.append(float(exact64[b][:, sel].min(dim=1).values.max()))
    return float(np.mean(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_ffps_agreement.json"))
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    warnings.filterwarnings("ignore")
    from de6d_b200 import pointnet2_utils as pu, synth
    from oracle import ref_py
    ours_tree, ref_tree = ref_py.load_pair()
    rpu = ref_tree.pointnet2_utils
    cpu_ = ours_tree.pointnet2_utils            # reference python over compat
    torch.backends.cuda.matmul.allow_tf32 = False
    B = args.batch
    cases = []
    for name, n, c, m, cloud, fscale in (
            ("bench_l2_uniform", 4096, 64, 512, "uniform", 1.0),
            ("bench_l2_lidar", 4096, 64, 512, "lidar", 1.0),
            ("small_feature_scale", 4096, 64, 512, "lidar", 0.05),     # geometry-dominated metric
            ("l3_shape", 512, 128, 256, "uniform", 1.0),
            ("xyz_only", 4096, 0, 512, "lidar", 1.0)):
        if cloud == "uniform":
            xyz = synth.clouds(B, 16384, seed=3)[:, :n].copy()
        else:
            xyz = synth.lidar_clouds(B, 16384, seed=3)[:, :n].copy()
        xyz = torch.from_numpy(xyz).cuda()
        feats = None if c == 0 else torch.from_numpy(synth.features(B, c, n, seed=5) * fscale).cuda()   # (B, C, N) like the backbone
        f_nc = None if feats is None else feats.permute(0, 2, 1)
        with torch.no_grad():
            x64 = xyz.double()
            exact64 = torch.cdist(x64, x64, compute_mode="donot_use_mm_for_euclid_dist")
            if f_nc is not None:
                f64 = f_nc.double().contiguous()
                exact64 = exact64 + torch.cdist(f64, f64, compute_mode="donot_use_mm_for_euclid_dist") * 1.0
            arms = {}
            arms["exact"] = rpu.furthest_point_sample_matrix(exact64.float().contiguous(), m)
            mat_ref = rpu.calc_dist_matrix_for_sampling(xyz, f_nc, 1.0) if f_nc is not None else rpu.calc_dist_matrix_for_sampling(xyz)
            arms["ref"] = rpu.furthest_point_sample_matrix(mat_ref.contiguous(), m)
            arms["compat"] = cpu_.furthest_point_sample_matrix(mat_ref.contiguous(), m)
            torch.backends.cuda.matmul.allow_tf32 = True
            mat_tf32 = rpu.calc_dist_matrix_for_sampling(xyz, f_nc, 1.0) if f_nc is not None else rpu.calc_dist_matrix_for_sampling(xyz)
            arms["ref_tf32"] = rpu.furthest_point_sample_matrix(mat_tf32.contiguous(), m)
            torch.backends.cuda.matmul.allow_tf32 = False
            arms["ours"] = pu.furthest_point_sample_features(xyz, f_nc, 1.0, m)
            mat_ours = pu.calc_dist_matrix_for_sampling(xyz, f_nc, 1.0)
            arms["ours2"] = pu.furthest_point_sample_matrix(mat_ours, m)
            err = {"ref_abs_max": float((mat_ref.double() - exact64).abs().max()),
                   "ref_tf32_abs_max": float((mat_tf32.double() - exact64).abs().max()),
                   "ours_abs_max": float((mat_ours.double() - exact64).abs().max()),
                   "ref_diag_max": float(mat_ref.diagonal(dim1=1, dim2=2).abs().max()),
                   "ours_diag_max": float(mat_ours.diagonal(dim1=1, dim2=2).abs().max())}
            arms = {k: v.cpu().numpy() for k, v in arms.items()}
            rec = {"case": name, "n": n, "channels": c, "npoint": m, "cloud": cloud, "feature_scale": fscale, "matrix_error": err,
                   "ours_equals_two_call": bool(np.array_equal(arms["ours"], arms["ours2"])),
                   "compat_equals_ref": bool(np.array_equal(arms["compat"], arms["ref"])),
                   "vs_exact": {k: compare(arms[k], arms["exact"]) for k in ("ref", "ref_tf32", "ours")},
                   "ours_vs_ref": compare(arms["ours"], arms["ref"]),
                   "ref_tf32_vs_ref": compare(arms["ref_tf32"], arms["ref"]),
                   "coverage_radius": {k: coverage(exact64, arms[k]) for k in ("exact", "ref", "ref_tf32", "ours")}}
            cases.append(rec)
            print(json.dumps(rec))
        del exact64, mat_ref, mat_tf32, mat_ours
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "batch": B, "cases": cases}, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()

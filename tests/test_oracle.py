"""CPU suite: the oracle against the golden vectors produced by the reference's own code, and against
independent numpy statements of the rules it encodes."""
import os

import numpy as np
import pytest

from de6d_b200 import synth


def _load(golden_dir, name):
    p = os.path.join(golden_dir, name)
    if not os.path.exists(p):
        pytest.skip("%s not generated yet" % name)
    return np.load(p)


def test_oracle_vs_reference_cpu_functions(orc, golden_dir):
    g = _load(golden_dir, "golden_cpu.npz")
    iou = orc.boxes_bev_iou_cpu(g["iou_a"], g["iou_b"])
    assert (g["iou_bev_cpu"] > 0).sum() > 20
    np.testing.assert_array_equal(iou, g["iou_bev_cpu"])  # same arithmetic, same libm: bit-exact
    mask = orc.points_in_boxes_cpu(g["pib_pts"], g["pib_boxes"])
    assert g["pib_cpu"].sum() > 50
    np.testing.assert_array_equal(mask, g["pib_cpu"])


def test_oracle_vs_reference_cuda_fps(orc, golden_dir):
    g = _load(golden_dir, "golden_cuda.npz")
    for tag, m in (("a", 64), ("b", 128), ("c", 40), ("d", 256)):
        xyz = g["fps_%s_xyz" % tag]
        idx, temp = orc.furthest_point_sample(xyz, m, return_temp=True)
        np.testing.assert_array_equal(idx, g["fps_%s_idx" % tag])
        np.testing.assert_array_equal(temp, g["fps_%s_temp" % tag])
        sidx = orc.furthest_point_sample_weights(xyz, g["sfps_%s_w" % tag], m)
        np.testing.assert_array_equal(sidx, g["sfps_%s_idx" % tag])
    fidx = orc.furthest_point_sample_matrix(g["ffps_mat"], 96)
    np.testing.assert_array_equal(fidx, g["ffps_idx"])


def test_oracle_vs_reference_cuda_ball_query(orc, golden_dir):
    g = _load(golden_dir, "golden_cuda.npz")
    xyz, new_xyz = g["bq_xyz"], g["bq_new_xyz"]
    for r in (0.5, 2.0):
        np.testing.assert_array_equal(orc.ball_query(r, 16, xyz, new_xyz), g["bq_idx_r%g" % r])
        cnt, idx = orc.ball_query_cnt(r, 16, xyz, new_xyz)
        np.testing.assert_array_equal(cnt, g["bqc_cnt_r%g" % r]); np.testing.assert_array_equal(idx, g["bqc_idx_r%g" % r])
        cnt, idx = orc.ball_query_dilated(r * 0.5, r, 16, xyz, new_xyz)
        np.testing.assert_array_equal(cnt, g["bqd_cnt_r%g" % r]); np.testing.assert_array_equal(idx, g["bqd_idx_r%g" % r])
        assert (g["bqc_cnt_r%g" % r] == 0).sum() >= 6 and (g["bqc_cnt_r%g" % r] > 0).sum() > 50


def test_oracle_vs_reference_cuda_interpolate(orc, golden_dir):
    g = _load(golden_dir, "golden_cuda.npz")
    dist, idx = orc.three_nn(g["nn_unknown"], g["nn_known"])
    np.testing.assert_array_equal(idx, g["nn_idx"])
    np.testing.assert_array_equal(dist, np.sqrt(g["nn_dist2"]))
    out = orc.three_interpolate(g["ti_feats"], g["nn_idx"], g["ti_weight"])
    np.testing.assert_array_equal(out, g["ti_out"])  # the FMA shape of the reference build is restated exactly


def test_oracle_vs_reference_cuda_boxes(orc, golden_dir):
    g = _load(golden_dir, "golden_cuda.npz")
    a, b = g["iou_gpu_a"], g["iou_gpu_b"]
    ov, iou = orc.boxes_overlap_bev(a, b), orc.boxes_iou_bev(a, b)
    assert (g["iou_gpu"] > 0).sum() > 100
    # rotated-IoU tree: the reference GPU build contracts FMAs the host restatement does not -> 1e-5 relative
    np.testing.assert_allclose(ov, g["overlap_gpu"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(iou, g["iou_gpu"], rtol=1e-5, atol=1e-7)
    np.testing.assert_array_equal(ov > 0, g["overlap_gpu"] > 0)
    for thr in (0.01, 0.1, 0.5):
        np.testing.assert_array_equal(orc.nms_sorted(g["nms_sorted_boxes"], thr), g["nms_keep_%g" % thr])
        np.testing.assert_array_equal(orc.nms_sorted(g["nms_sorted_boxes"], thr, normal=True), g["nmsn_keep_%g" % thr])
    out = orc.points_in_boxes_gpu(g["pibg_pts"], g["pibg_boxes"])
    assert (g["pibg_out"] >= 0).sum() > 100
    np.testing.assert_array_equal(out, g["pibg_out"])


# ---- independent statements of the rules the oracle encodes ---------------------------------------------------

def _fps_numpy(xyz, m, bs):
    """Direct simulation of the reference CTA: strided per-thread scan + tie-to-lower-slot tree."""
    n = xyz.shape[0]
    temp = np.full(n, 1e10, np.float32)
    out = [0]
    old = 0
    for _ in range(1, m):
        d = xyz - xyz[old]
        dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
        # float32 fma emulated in float64 (products of float32 are exact in float64, one rounding at the end)
        t = (dy * dy).astype(np.float32)
        t = (dx.astype(np.float64) * dx.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        dist = (dz.astype(np.float64) * dz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
        temp = np.minimum(dist, temp)
        best = np.full(bs, -1.0, np.float32); besti = np.zeros(bs, np.int64)
        for k in range(n):
            s = k % bs
            if temp[k] > best[s]:
                best[s] = temp[k]; besti[s] = k
        half = bs // 2
        while half >= 1:
            for t_ in range(half):
                if best[t_ + half] > best[t_]:
                    best[t_] = best[t_ + half]; besti[t_] = besti[t_ + half]
            half //= 2
        old = int(besti[0])
        out.append(old)
    return np.array(out, np.int32)


@pytest.mark.parametrize("n,m,dup", [(64, 16, 0.3), (100, 24, 0.2), (257, 20, 0.0)])
def test_oracle_fps_tie_rule_small(orc, n, m, dup):
    xyz = synth.clouds(1, n, seed=5, dup_frac=dup)
    # double-rounding caveat: float64 emulation of fma is exact here because |values| < 2^24 ulp spacing
    got = orc.furthest_point_sample(xyz, m)[0]
    want = _fps_numpy(xyz[0], m, orc.opt_n_threads(n))
    np.testing.assert_array_equal(got, want)


def test_opt_n_threads_matches_reference_formula(orc):
    import math
    for n in list(range(1, 70)) + [127, 128, 129, 255, 256, 511, 512, 513, 1000, 1023, 1024, 1025, 4096, 16383, 16384, 131072]:
        p = int(math.log(float(n)) / math.log(2.0))
        assert orc.opt_n_threads(n) == max(min(1 << p, 1024), 1)


def test_ball_query_padding_rules(orc):
    xyz = np.zeros((1, 8, 3), np.float32); xyz[0, :, 0] = np.arange(8)
    q = np.array([[[2.0, 0, 0], [100.0, 0, 0]]], np.float32)
    idx = orc.ball_query(1.5, 5, xyz, q)
    np.testing.assert_array_equal(idx[0, 0], [1, 2, 3, 1, 1]); np.testing.assert_array_equal(idx[0, 1], [0] * 5)
    cnt, idx = orc.ball_query_cnt(1.5, 5, xyz, q)
    np.testing.assert_array_equal(cnt[0], [3, 0]); np.testing.assert_array_equal(idx[0, 0], [1, 2, 3, 1, 2])
    cnt, idx = orc.ball_query_dilated(0.5, 1.5, 4, xyz, q)
    np.testing.assert_array_equal(cnt[0], [2, 0]); np.testing.assert_array_equal(idx[0, 0], [1, 3, 1, 3])


def test_iou_known_answers(orc):
    a = np.array([[0, 0, 0, 4, 2, 1, 0.0]], np.float32)
    assert abs(orc.boxes_iou_bev(a, a)[0, 0] - 1.0) < 1e-6
    b = np.array([[2, 0, 0, 4, 2, 1, 0.0]], np.float32)       # half overlap: 4 / (8 + 8 - 4)
    assert abs(orc.boxes_iou_bev(a, b)[0, 0] - 1.0 / 3.0) < 1e-3  # MARGIN 1e-2 on corner containment
    c = np.array([[0, 0, 0, 2, 2, 1, np.pi / 4]], np.float32)  # diamond inside a 4x4 square
    d = np.array([[0, 0, 0, 4, 4, 1, 0.0]], np.float32)
    assert abs(orc.boxes_overlap_bev(c, d)[0, 0] - 4.0) < 1e-4
    far = np.array([[50, 50, 0, 4, 2, 1, 0.3]], np.float32)
    assert orc.boxes_iou_bev(a, far)[0, 0] == 0.0
    # 3-D: identical boxes shifted by half the height
    e = a.copy(); e[0, 2] = 0.5
    assert abs(orc.boxes_iou3d(a, e)[0, 0] - (8 * 0.5) / (8 + 8 - 4)) < 1e-3


def test_nms_sweep_properties(orc):
    bx, sc = synth.proposals(1, 256, seed=9)
    keep = orc.nms_gpu(bx[0], sc[0], 0.1)
    assert 0 < len(keep) < 256
    kept = bx[0][keep]
    iou = orc.boxes_iou_bev(kept, kept)
    np.fill_diagonal(iou, 0)
    assert iou.max() <= 0.1                       # survivors do not suppress each other
    keep2 = orc.nms_gpu(kept, sc[0][keep], 0.1)   # idempotence
    np.testing.assert_array_equal(np.sort(keep2), np.arange(len(keep)))


def test_dist_matrix_against_float64(orc):
    """orc.calc_dist_matrix_for_sampling (pointnet2_utils.py:36-44 with direct differences) against float64 torch.cdist."""
    import torch
    xyz = synth.clouds(2, 150, seed=9, dup_frac=0.1)
    f = np.ascontiguousarray(synth.features(2, 24, 150, seed=9).transpose(0, 2, 1))
    got = orc.calc_dist_matrix_for_sampling(xyz, f, 0.5)
    x64, f64 = torch.from_numpy(xyz).double(), torch.from_numpy(f).double()
    ref = (torch.cdist(x64, x64) + torch.cdist(f64, f64) * 0.5).numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-6)
    np.testing.assert_array_equal(got, got.transpose(0, 2, 1))
    assert (np.diagonal(got, axis1=1, axis2=2) == 0).all()


def test_points_in_boxes3d_against_scipy(orc):
    """orc.points_in_boxes3d against scipy's own Rotation.from_euler('zyx') + Delaunay.find_simplex, the two
    third-party pieces box_utils.points_in_boxes3d (box_utils.py:59-72,110-124) is made of; points closer than 1e-6 m
    to a box face are excluded (Delaunay's own tolerance decides those)."""
    from scipy.spatial import Delaunay
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    T, M = 14, 4000
    boxes = np.concatenate([synth.boxes(1, T, seed=8)[0][:, :7], rng.uniform(-0.4, 0.4, (T, 2)).astype(np.float32)], 1)
    boxes[5, :3] = boxes[4, :3] + 0.3                                  # overlapping boxes: the later one must win
    pts = (boxes[rng.integers(0, T, M), :3] + rng.normal(0, 1.2, (M, 3))).astype(np.float32)
    got = orc.points_in_boxes3d(pts, boxes)
    template = np.array([[1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1], [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1]]) / 2
    want = np.full(M, -1, np.int64)
    near = np.zeros(M, bool)
    for i in range(T):
        R = Rotation.from_euler("zyx", boxes[i, 6:9].astype(np.float64)).as_matrix()
        corners = (boxes[i, 3:6].astype(np.float64) * template) @ R.T + boxes[i, :3].astype(np.float64)
        want[Delaunay(corners).find_simplex(pts.astype(np.float64)) >= 0] = i
        local = (pts.astype(np.float64) - boxes[i, :3]) @ R
        near |= (np.abs(np.abs(local) - boxes[i, 3:6] / 2.0) < 1e-6).any(1)
    assert (want >= 0).sum() > 500 and (want == 5).sum() > 10
    np.testing.assert_array_equal(got[~near], want[~near])


def _rand_box9(rng, sigma=1.5, tilt=0.6):
    return np.concatenate([rng.normal(0, sigma, 3), rng.uniform(1, 4, 3), rng.uniform(-np.pi, np.pi, 1),
                           rng.uniform(-tilt, tilt, 2)]).astype(np.float32)


def _scipy_intersection_volume(a, b):
    """float64 volume of the intersection of two full-pose boxes with scipy only: half-spaces from
    Rotation.from_euler('zyx', (rz, ry, rx)) exactly as box_utils.boxes3d_to_corners_3d (pcdet/utils/box_utils.py:59-72),
    an interior point from a Chebyshev-centre LP, HalfspaceIntersection vertices, ConvexHull volume."""
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection
    from scipy.spatial.transform import Rotation
    hs = []
    for bx in (a, b):
        R = Rotation.from_euler("zyx", bx[6:9].astype(np.float64)).as_matrix()
        c, h = bx[:3].astype(np.float64), bx[3:6].astype(np.float64) / 2
        for j in range(3):
            n = R[:, j]
            hs.append(np.concatenate([n, [-(n @ c) - h[j]]]))
            hs.append(np.concatenate([-n, [(n @ c) - h[j]]]))
    hs = np.array(hs)
    A, rhs = hs[:, :3], -hs[:, 3]
    res = linprog([0, 0, 0, -1], A_ub=np.hstack([A, np.linalg.norm(A, axis=1)[:, None]]), b_ub=rhs,
                  bounds=[(None, None)] * 3 + [(0, None)])
    if res.status != 0 or res.x[3] < 1e-9:
        return 0.0
    return ConvexHull(HalfspaceIntersection(hs, res.x[:3]).intersections).volume


def test_full_pose_iou_oracle_vs_scipy(orc):
    """The 9-DoF intersection volume restated in the oracle (no reference function exists: SURVEY 8f rank 3) against
    scipy.spatial in float64, on generic pairs and on the degenerate configurations NMS meets (identical boxes, nested
    boxes, shared faces); for yaw-only boxes against the reference's own boxes_iou3d_gpu composition."""
    rng = np.random.default_rng(0)
    worst, nonzero = 0.0, 0
    for _ in range(250):
        a, b = _rand_box9(rng), _rand_box9(rng)
        v, w = orc.box9_intersection_volume(a, b), _scipy_intersection_volume(a, b)
        worst = max(worst, abs(v - w) / max(w, 1e-3))
        nonzero += w > 1e-6
    assert worst < 1e-9 and nonzero > 80
    a = _rand_box9(rng)
    vol = float(a[3]) * float(a[4]) * float(a[5])
    assert abs(orc.box9_intersection_volume(a, a) - vol) < 1e-9 * vol            # coplanar faces are not counted twice
    b = a.copy(); b[3:6] *= 0.5
    assert abs(orc.box9_intersection_volume(a, b) - vol / 8) < 1e-6 * vol and abs(orc.box9_intersection_volume(b, a) - vol / 8) < 1e-6 * vol
    b = a.copy(); b[3] *= 0.5                                                   # four shared face planes
    assert abs(orc.box9_intersection_volume(a, b) - vol / 2) < 1e-6 * vol
    far = a.copy(); far[0] += 50
    assert orc.box9_intersection_volume(a, far) == 0.0
    iou = orc.boxes_iou3d_9dof(np.stack([a, b, far]), np.stack([a, b, far]))
    np.testing.assert_allclose(np.diag(iou), 1.0, atol=1e-6)
    assert abs(iou[0, 1] - 0.5) < 1e-6 and iou[0, 2] == 0
    # yaw-only boxes: the reference composition (BEV clipping with 1e-2 m padded corner tests) approximates the same number
    A = np.stack([_rand_box9(rng) for _ in range(60)]); A[:, 7:] = 0; A[:, 2] *= 0.2
    i9, i7 = orc.boxes_iou3d_9dof(A, A), orc.boxes_iou3d(A[:, :7], A[:, :7])
    assert (i7 > 1e-3).sum() > 500 and np.abs(i9 - i7).max() < 5e-3
    keep = orc.nms_9dof(A, np.arange(60, 0, -1, dtype=np.float32), 0.1)
    assert keep[0] == 0 and 1 < len(keep) < 60


def test_oracle_vs_reference_wrapper_golden(orc, golden_dir):
    """Second fixture: the reference's python wrappers over its own kernels (tests/golden/make_golden.py --cuda2):
    boxes_iou3d_gpu, gather / group / three_interpolate forward + backward, QueryWithCntAndGroup."""
    g = _load(golden_dir, "golden_cuda2.npz")
    iou = orc.boxes_iou3d(g["iou3d_a"], g["iou3d_b"])
    assert (g["iou3d"] > 0.01).sum() > 40
    np.testing.assert_allclose(iou, g["iou3d"], rtol=1e-5, atol=1e-6)      # grazing overlaps (IoU ~ 1e-3) carry absolute, not relative, noise
    N = g["gg_feats"].shape[2]
    np.testing.assert_array_equal(orc.gather_operation(g["gg_feats"], g["gather_idx"]), g["gather_out"])
    np.testing.assert_allclose(orc.gather_operation_grad(g["gather_gout"], g["gather_idx"], N), g["gather_grad"], rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(orc.grouping_operation(g["gg_feats"], g["group_idx"]), g["group_out"])
    np.testing.assert_allclose(orc.grouping_operation_grad(g["group_gout"], g["group_idx"], N), g["group_grad"], rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(orc.three_interpolate(g["ti2_feats"], g["ti2_idx"], g["ti2_weight"]), g["ti2_out"])
    np.testing.assert_allclose(orc.three_interpolate_grad(g["ti2_gout"], g["ti2_idx"], g["ti2_weight"], g["ti2_feats"].shape[2]),
                               g["ti2_grad"], rtol=1e-5, atol=1e-5)
    cnt, idx = orc.ball_query_cnt(1.5, g["group_idx"].shape[2], g["qg_xyz"], g["qg_new_xyz"])
    np.testing.assert_array_equal(cnt, g["qg_cnt"])
    gx = orc.grouping_operation(np.ascontiguousarray(g["qg_xyz"].transpose(0, 2, 1)), idx) - g["qg_new_xyz"].transpose(0, 2, 1)[..., None]
    np.testing.assert_array_equal(np.concatenate([gx, orc.grouping_operation(g["gg_feats"], idx)], 1), g["qg_out"])

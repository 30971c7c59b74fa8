"""CPU suite: the two entry points the reference itself runs on the host (boxes_iou_bev_cpu, iou3d_cpu.cpp:232-252;
points_in_boxes_cpu, roiaware_pool3d.cpp:143-168) through the reference-facing wrappers of this package:
bit-identical to the golden vectors made by the reference functions, to the reference build run live (when
oracle/_ref is present), and usable inside forked DataLoader workers (database_sampler.py:232-233,
kitti_dataset.py:248 call them there; a CUDA context cannot be created in a forked child)."""
import os

import numpy as np
import pytest
import torch


def _boxes(n, rng):
    b = np.zeros((n, 7), np.float32)
    b[:, 0] = rng.uniform(0, 30, n); b[:, 1] = rng.uniform(-15, 15, n); b[:, 2] = rng.uniform(-2, 0, n)
    b[:, 3:6] = np.abs(rng.normal((3.9, 1.6, 1.56), 0.4, (n, 3))) + 0.05
    b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    return b


def test_host_ops_match_reference_golden(lib, golden_dir):
    from de6d_b200 import iou3d_nms_utils, roiaware_pool3d_utils
    g = np.load(os.path.join(golden_dir, "golden_cpu.npz"))
    np.testing.assert_array_equal(iou3d_nms_utils.boxes_bev_iou_cpu(g["iou_a"], g["iou_b"]), g["iou_bev_cpu"])
    np.testing.assert_array_equal(roiaware_pool3d_utils.points_in_boxes_cpu(g["pib_pts"], g["pib_boxes"]), g["pib_cpu"])
    # torch CPU tensors in -> torch out
    out = iou3d_nms_utils.boxes_bev_iou_cpu(torch.from_numpy(g["iou_a"]), torch.from_numpy(g["iou_b"]))
    assert isinstance(out, torch.Tensor) and np.array_equal(out.numpy(), g["iou_bev_cpu"])


def test_host_ops_match_oracle_on_edge_cases(lib, orc):
    """Identical boxes, touching boxes, zero-size (padded gt) boxes, axis-aligned pairs sharing edges (the EPS branch of
    the segment intersection), empty inputs; row-threaded evaluation gives the same bits as the single-thread one."""
    from de6d_b200 import iou3d_nms_utils, roiaware_pool3d_utils
    from de6d_b200.compat import iou3d_nms_cuda, roiaware_pool3d_cuda
    rng = np.random.default_rng(5)
    a = _boxes(120, rng)
    b = np.concatenate([a[:30], _boxes(60, rng)])
    b[30:40, 6] = 0.0; b[30:40, 3:5] = (4.0, 2.0)
    b[40:50] = b[30:40]; b[40:50, 0] += 4.0            # shares an edge with its neighbour
    b[50:55, 3:6] = 0.0                                 # zero-padded gt boxes
    a[100:110] = b[30:40]
    got = iou3d_nms_utils.boxes_bev_iou_cpu(a, b)
    np.testing.assert_array_equal(got, orc.boxes_bev_iou_cpu(a, b))
    assert np.isfinite(got).all() and (got > 0.99).sum() >= 30
    pts = np.concatenate([rng.uniform(-1, 31, (5000, 3)) * [1, 1, 0.1] - [0, 15.5, 1.0],
                          a[:50, :3] + rng.normal(0, 1.0, (50, 3))]).astype(np.float32)
    np.testing.assert_array_equal(roiaware_pool3d_utils.points_in_boxes_cpu(pts, a), orc.points_in_boxes_cpu(pts, a))
    assert iou3d_nms_utils.boxes_bev_iou_cpu(a[:0], b).shape == (0, len(b))
    assert roiaware_pool3d_utils.points_in_boxes_cpu(pts[:0], a).shape == (len(a), 0)
    for mod in (iou3d_nms_cuda, roiaware_pool3d_cuda):
        mod.HOST_THREADS = 4
    try:
        np.testing.assert_array_equal(iou3d_nms_utils.boxes_bev_iou_cpu(a, b), got)
        np.testing.assert_array_equal(roiaware_pool3d_utils.points_in_boxes_cpu(pts, a), orc.points_in_boxes_cpu(pts, a))
    finally:
        for mod in (iou3d_nms_cuda, roiaware_pool3d_cuda):
            mod.HOST_THREADS = 1
    with pytest.raises(TypeError):
        iou3d_nms_cuda.boxes_iou_bev_cpu(torch.zeros(2, 7, dtype=torch.float64), torch.zeros(2, 7), torch.zeros(2, 2))


def test_host_ops_match_live_reference_build(lib):
    """The unmodified reference wrappers over compat vs over the reference's own extension modules, on the CPU."""
    from oracle import build_ref, ref_py
    if not (build_ref.available() and ref_py.available()):
        pytest.skip("oracle/_ref not present")
    import warnings
    warnings.filterwarnings("ignore")
    ours, theirs = ref_py.load_pair()
    rng = np.random.default_rng(11)
    a, b = _boxes(150, rng), _boxes(90, rng)
    np.testing.assert_array_equal(ours.iou3d_nms_utils.boxes_bev_iou_cpu(a, b), theirs.iou3d_nms_utils.boxes_bev_iou_cpu(a, b))
    pts = (rng.uniform(-1, 31, (8000, 4)) * [1, 1, 0.1, 1] - [0, 15.5, 1.0, 0]).astype(np.float32)
    np.testing.assert_array_equal(ours.roiaware_pool3d_utils.points_in_boxes_cpu(pts[:, :3], a),
                                  theirs.roiaware_pool3d_utils.points_in_boxes_cpu(pts[:, :3], a))
    np.testing.assert_array_equal(ours.box_utils.remove_points_in_boxes3d(pts, a), theirs.box_utils.remove_points_in_boxes3d(pts, a))


class _Aug(torch.utils.data.Dataset):
    """What database_sampler / kitti_dataset do per sample inside a worker."""

    def __len__(self):
        return 4

    def __getitem__(self, i):
        from de6d_b200 import iou3d_nms_utils, roiaware_pool3d_utils
        rng = np.random.default_rng(i)
        a, b = _boxes(20, rng), _boxes(15, rng)
        pts = rng.uniform(0, 30, (500, 3)).astype(np.float32)
        return iou3d_nms_utils.boxes_bev_iou_cpu(a, b), roiaware_pool3d_utils.points_in_boxes_cpu(pts, a)


def test_host_ops_work_in_forked_dataloader_workers(lib, orc):
    import multiprocessing as mp
    if torch.cuda.is_available():
        torch.zeros(1, device="cuda")            # a CUDA context in the parent is what makes forked children unable to use CUDA
    dl = torch.utils.data.DataLoader(_Aug(), batch_size=None, num_workers=2, multiprocessing_context=mp.get_context("fork"))
    for i, (iou, mask) in enumerate(dl):
        rng = np.random.default_rng(i)
        a, b = _boxes(20, rng), _boxes(15, rng)
        pts = rng.uniform(0, 30, (500, 3)).astype(np.float32)
        np.testing.assert_array_equal(np.asarray(iou), orc.boxes_bev_iou_cpu(a, b))
        np.testing.assert_array_equal(np.asarray(mask), orc.points_in_boxes_cpu(pts, a))

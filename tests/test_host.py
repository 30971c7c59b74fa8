"""CPU suite: host-side logic -- sharding, the gloo gather path (world_size 2), chain configuration, the compat
module registration, and that the ops refuse to run without CUDA tensors (no silent CPU path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from de6d_b200 import dist as ddist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for total in (0, 1, 7, 16, 512, 513):
        for world in (1, 2, 4, 8):
            spans = [ddist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = ddist.init_from_env(backend="gloo")
    lo, hi = ddist.shard_range(total, r, w)
    per = -(-total // w)
    keep = torch.zeros((per, 4), dtype=torch.int64); num = torch.zeros(per, dtype=torch.int32)
    for i, f in enumerate(range(lo, hi)):           # "detections" of frame f: f, f+1, ...
        keep[i, : (f % 4) + 1] = torch.arange(f, f + (f % 4) + 1); num[i] = (f % 4) + 1
    keep_all, num_all = ddist.gather_detections(keep, num)
    t = ddist.max_over_ranks(float(rank + 1))
    # the full detection gather: boxes of frame f are f + 0.01 * proposal index, scores the proposal index
    boxes = torch.stack([(f + 0.01 * torch.arange(16, dtype=torch.float32)).unsqueeze(1).expand(16, 7) for f in range(lo, hi)])
    scores = torch.arange(16, dtype=torch.float32).unsqueeze(0).expand(hi - lo, 16).contiguous()
    kidx = torch.stack([torch.arange(4) + (f % 3) for f in range(lo, hi)])
    b_all, s_all, n_all = ddist.gather_detection_boxes(boxes, scores, kidx, num)
    if rank == 0:
        q.put((keep_all.numpy(), num_all.numpy(), t, b_all.numpy(), s_all.numpy(), n_all.numpy()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_gather_detections_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, total, world = _free_port(), 6, 2
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    keep_all, num_all, t, b_all, s_all, n_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0
    assert keep_all.shape == (6, 4)
    for f in range(total):   # concatenated shards == what a single rank would have produced
        n = (f % 4) + 1
        assert num_all[f] == n
        np.testing.assert_array_equal(keep_all[f, :n], np.arange(f, f + n))
        assert n_all[f] == n and b_all.shape == (6, 4, 7) and s_all.shape == (6, 4)
        want = np.arange(4) + (f % 3)
        np.testing.assert_allclose(s_all[f, :n], want[:n])
        np.testing.assert_allclose(b_all[f, :n, 0], f + 0.01 * want[:n], rtol=1e-6)
        assert (b_all[f, n:] == 0).all() and (s_all[f, n:] == 0).all()


def test_chain_config_shapes():
    from de6d_b200 import chain
    cfg = chain.ChainConfig()
    n = cfg.n_points
    for layer in cfg.layers:
        assert len(layer.npoints) == len(layer.methods) == len(layer.ranges)
        for (lo, hi) in layer.ranges:
            assert 0 <= lo < hi <= n
        n = sum(layer.npoints)
    assert n == 512 and cfg.n_votes <= n
    host = chain.make_inputs(chain.small_config(), batch=2, seed=0, pinned=False)
    assert host["xyz"].shape == (2, 2048, 3) and host["boxes"].shape == (2, 128, 7)
    again = chain.make_inputs(chain.small_config(), batch=2, seed=0, pinned=False)
    assert all(torch.equal(host[k], again[k]) for k in host)


def test_compat_install_registers_reference_module_paths(lib):
    from de6d_b200 import compat
    names = compat.install()
    assert set(names) == {"pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda", "pcdet.ops.iou3d_nms.iou3d_nms_cuda",
                          "pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda"}
    p2 = sys.modules["pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda"]
    for fn in ("ball_query_wrapper", "ball_query_cnt_wrapper", "ball_query_dilated_wrapper", "group_points_wrapper",
               "group_points_grad_wrapper", "gather_points_wrapper", "gather_points_grad_wrapper",
               "farthest_point_sampling_wrapper", "furthest_point_sampling_matrix_wrapper",
               "furthest_point_sampling_weights_wrapper", "three_nn_wrapper", "three_interpolate_wrapper",
               "three_interpolate_grad_wrapper"):
        assert callable(getattr(p2, fn))
    for k in list(names):
        sys.modules.pop(k, None)


def test_ops_reject_cpu_tensors_instead_of_falling_back(lib):
    from de6d_b200.compat import pointnet2_batch_cuda as p2, iou3d_nms_cuda as iou
    xyz = torch.zeros(1, 8, 3); temp = torch.zeros(1, 8); idx = torch.zeros(1, 2, dtype=torch.int32)
    with pytest.raises(ValueError, match="CUDA tensor"):
        p2.farthest_point_sampling_wrapper(1, 8, 2, xyz, temp, idx)
    with pytest.raises(ValueError, match="CUDA tensor"):
        iou.boxes_iou_bev_gpu(torch.zeros(2, 7), torch.zeros(2, 7), torch.zeros(2, 2))


def test_mirror_modules_expose_reference_names(lib):
    from de6d_b200 import pointnet2_utils as pu, iou3d_nms_utils as iu, roiaware_pool3d_utils as ru
    for name in ("furthest_point_sample", "farthest_point_sample", "furthest_point_sample_matrix", "furthest_point_sample_weights",
                 "calc_dist_matrix_for_sampling", "gather_operation", "three_nn", "three_interpolate", "grouping_operation",
                 "ball_query", "ball_query_cnt", "ball_query_dilated", "QueryAndGroup", "QueryWithCntAndGroup",
                 "QueryAndGroupDilated", "GroupAll"):
        assert hasattr(pu, name), name
    for name in ("boxes_bev_iou_cpu", "boxes_iou_bev", "boxes_iou3d_gpu", "nms_gpu", "nms_normal_gpu"):
        assert hasattr(iu, name), name
    for name in ("points_in_boxes_cpu", "points_in_boxes_gpu"):
        assert hasattr(ru, name), name


def test_small_cloud_fps_register_layout_is_the_reference_tie_order():
    """fps_small.cu keeps point k = c*B + (bitrev(a) << 5) + lane in slot J = a*C + c of its lane and relies on two facts:
    every point of the cloud appears in exactly one (lane, slot), and within a lane the reference's tie priority
    (common.cuh: fps_prio = bit-reversed (k mod B), then k / B) ascends with the slot -- so a strict '>' scan in slot
    order picks the same point the reference's strided in-thread scan + tree reduction picks.  Checked here for every
    layout class the launcher can choose (restating the index arithmetic of the kernel)."""
    import math

    def brev(x, bits):
        return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0

    for n in (32, 33, 40, 63, 64, 100, 300, 511, 512, 640, 777, 1000, 1023, 1024, 1025, 1500, 2048, 3000, 3072, 4095, 4096):
        log2b = min(int(math.log(n) / math.log(2.0)), 10)
        B, hbits = 1 << log2b, log2b - 5
        C = (n + B - 1) // B
        need = (B // 32) * C
        assert need <= 128 and C <= 4
        cinv = (65536 + C - 1) // C
        seen = set()
        for lane in range(32):
            last = -1
            for J in range(need):
                a, c = J // C, J % C
                assert (J * cinv) >> 16 == a                     # the multiply-shift division used in the sample loop
                k = c * B + (brev(a, hbits) << 5) + lane
                if k >= n:
                    continue
                seen.add(k)
                prio = (brev(k % B, log2b) << 22) | (k // B)      # fps_prio
                assert prio == ((brev(lane, 5) << hbits | a) << 22) | c
                assert prio > last
                last = prio
        assert seen == set(range(n))


def test_kernel_form_planning_is_host_side():
    """Shape planning of the cluster F-FPS kernels and of the fused SA scale runs on the host (no CUDA call): which shapes fit, how
    an oversize last layer is split, and that bad selectors are refused before anything is launched."""
    import ctypes as C
    from de6d_b200 import _lib, sa_fused
    lib = _lib.load()
    fits = lib.de6d_furthest_point_sampling_features_fits
    assert fits(4096, 64) == 1 and fits(3073, 64) == 1 and fits(512, 128) == 1
    assert fits(130, 128) == 1            # only the pruned form covers few points with many channels
    assert fits(0, 64) == 0 and fits(100000, 64) == 0 and fits(16384, 64) == 0
    plan = sa_fused.FusedSAScale.plan
    assert plan([4, 16, 16, 32], 32) == (1, 1)              # SA1: one CTA holds all weights
    assert plan([131, 128, 128, 256], 32) == (2, 1)         # SA3 scale 1: a cta_group::2 pair
    assert plan([131, 128, 256, 256], 32) == (2, 2)         # SA3 scale 3: the last layer in two launches
    assert plan([259, 256, 256, 512], 16) == (0, 0)         # vote head: the first layers alone exceed an SM pair
    assert plan([35, 24, 32], 32) == (0, 0) and plan([35, 32, 32], 24) == (0, 0)
    null = C.c_void_p(0)
    for cluster, prune in ((5, 0), (0, 4), (7, 1), (0, -1)):
        rc = lib.de6d_furthest_point_sampling_features_impl(1, 4096, 64, 8, null, null, 0, 0, 0, 1.0, null, null, cluster, prune, null)
        assert rc == 1 and b"must be" in lib.de6d_last_error_string()

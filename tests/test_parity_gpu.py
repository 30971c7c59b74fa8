"""GPU parity suite (pytest -m gpu): the sm_100a kernels, called through the Python mirror -> compat -> C ABI,
against (1) the CPU oracle on the same seeded inputs, (2) the committed golden vectors produced by the
reference's own CUDA kernels, (3) the reference's kernels run live when oracle/_ref is present on the box, and
(4) size-independent properties at BASELINE.json's full sizes.

Bars (BASELINE.json north_star): indices, counts and keep lists bit-exact; copies bit-exact; interpolated
features and IoUs within 1e-5 relative (three_interpolate is in fact bit-exact: its FMA shape is pinned).
"""
import os

import numpy as np
import pytest
import torch

from de6d_b200 import synth

pytestmark = pytest.mark.gpu

RTOL_IOU = 1e-5


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def ops(lib):
    from de6d_b200 import iou3d_nms_utils, pointnet2_utils, roiaware_pool3d_utils
    return pointnet2_utils, iou3d_nms_utils, roiaware_pool3d_utils


@pytest.fixture(scope="module")
def golden(golden_dir):
    p = os.path.join(golden_dir, "golden_cuda.npz")
    if not os.path.exists(p):
        pytest.skip("golden_cuda.npz not generated yet")
    return np.load(p)


def _fps_impl(xyz, m, impl, weights=None):
    from de6d_b200._lib import call
    B, N, _ = xyz.shape
    x = cu(xyz)
    temp = torch.full((B, N), 1e10, device="cuda")
    idx = torch.zeros((B, m), dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    if weights is None:
        call("de6d_furthest_point_sampling_impl", B, N, m, x.data_ptr(), temp.data_ptr(), idx.data_ptr(), impl, s)
    else:
        w = cu(weights)
        call("de6d_furthest_point_sampling_weights_impl", B, N, m, x.data_ptr(), w.data_ptr(), temp.data_ptr(), idx.data_ptr(), impl, s)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), temp.cpu().numpy()


# ------------------------------------------------------------------------------------------------ FPS
@pytest.mark.parametrize("B,N,M,dup", [(2, 1000, 64, 0.1), (3, 2048, 300, 0.2), (1, 300, 299, 0.0), (2, 4096, 512, 0.05),
                                       (1, 33, 9, 0.3), (1, 1, 1, 0.0), (2, 5000, 200, 0.1), (1, 16383, 160, 0.1)])
@pytest.mark.parametrize("impl", [0, 1, 2, 4, 5])
def test_dfps_vs_oracle(orc, lib, B, N, M, dup, impl):
    xyz = synth.clouds(B, N, seed=N + M, dup_frac=dup)
    want_idx, want_temp = orc.furthest_point_sample(xyz, M, return_temp=True)
    got_idx, got_temp = _fps_impl(xyz, M, impl)
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_temp, want_temp)  # temp is an in/out tensor of the op


@pytest.mark.parametrize("B,N,M", [(2, 1000, 64), (2, 2048, 200), (1, 4096, 256), (1, 16384, 200), (1, 40, 40)])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_sfps_vs_oracle(orc, lib, B, N, M, impl):
    xyz = synth.clouds(B, N, seed=7, dup_frac=0.1)
    w = synth.weights(B, N, seed=8)
    w[:, :5] = 0.0          # max(w, 1e-12) double path
    w[:, 5:9] = 1e-13
    want_idx, want_temp = orc.furthest_point_sample_weights(xyz, w, M, return_temp=True)
    got_idx, got_temp = _fps_impl(xyz, M, impl, weights=w)
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_temp, want_temp)


@pytest.mark.parametrize("B,N,M,dup", [(2, 16385, 70, 0.1), (1, 40000, 300, 0.2), (2, 131072, 200, 0.05), (1, 100000, 64, 0.0)])
def test_dfps_cluster_large_clouds(orc, lib, B, N, M, dup):
    """Clouds larger than one SM (BASELINE configs[4]: 131072 points): thread-block-cluster kernel (one CTA per
    16384-point slice, candidates exchanged through distributed shared memory) == generic kernel == oracle,
    indices and the written-back min-distances."""
    xyz = synth.lidar_clouds(B, N, seed=N, beams=64) if N % 64 == 0 else synth.clouds(B, N, seed=N, dup_frac=dup)
    if N % 64 == 0:
        xyz[:, 17000] = xyz[:, 5]; xyz[:, N - 1] = xyz[:, 16384]     # exact ties across CTA slices
    want, wtemp = orc.furthest_point_sample(xyz, M, return_temp=True)
    for impl in (0, 2):
        idx, temp = _fps_impl(xyz, M, impl)
        np.testing.assert_array_equal(idx, want)
        np.testing.assert_array_equal(temp, wtemp)


def test_dfps_cluster_stress_shape(lib):
    """131072 -> 16384 (the stress config): cluster kernel == generic kernel."""
    xyz = synth.lidar_clouds(1, 131072, seed=9)
    a, ta = _fps_impl(xyz, 16384, 0)
    b, tb = _fps_impl(xyz, 16384, 2)
    np.testing.assert_array_equal(a, b); np.testing.assert_array_equal(ta, tb)
    assert len(set(a[0].tolist())) == 16384


@pytest.mark.parametrize("N,M,kind", [(16384, 4096, "uniform"), (16384, 4096, "lidar"), (4096, 4096, "dup"), (3000, 1500, "grid"),
                                      (512, 512, "uniform"), (16384, 600, "clusters"), (16384, 5000, "dup"), (12000, 3000, "grid16")])
def test_dfps_multi_sample_rounds_exact(orc, lib, N, M, kind):
    """The default D-FPS kernel takes up to 4 samples per barrier round (speculation that is only accepted when it
    provably equals the sequential choice).  Stress the acceptance logic: exhaustive sampling (M == N: rounds where
    all remaining min-distances tie or are zero), heavy duplication, lattice clouds (masses of exact distance ties),
    tight clusters (several top candidates inside one bucket / one warp), against the oracle and the
    one-sample-per-round kernel, indices and written-back min-distances."""
    rng = np.random.default_rng(N + M)
    if kind == "uniform":
        xyz = synth.clouds(2, N, seed=N)
    elif kind == "lidar":
        xyz = synth.lidar_clouds(2, N, seed=N)
    elif kind == "dup":
        xyz = synth.clouds(2, N, seed=N, dup_frac=0.6)
    elif kind == "grid":
        g = np.stack(np.meshgrid(np.arange(15), np.arange(20), np.arange(10), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        xyz = np.stack([g[rng.permutation(N)] * np.float32(0.5), g[rng.permutation(N)] * np.float32(0.25)])
    elif kind == "grid16":
        g = np.stack(np.meshgrid(np.arange(30), np.arange(40), np.arange(10), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        xyz = np.stack([g[rng.permutation(N)] * np.float32(0.5), g[rng.permutation(N)] * np.float32(0.25)])
    else:
        centres = rng.uniform(0, 50, (12, 3))
        xyz = (centres[rng.integers(0, 12, (2, N))] + rng.normal(0, 0.05, (2, N, 3))).astype(np.float32)
    want, wtemp = orc.furthest_point_sample(xyz, M, return_temp=True)
    for impl in (0, 4, 5, 7, 8):     # automatic, one sample per round, multi-sample rounds forced at every size, 6 / 8 samples per round
        idx, temp = _fps_impl(xyz, M, impl)
        np.testing.assert_array_equal(idx, want)
        np.testing.assert_array_equal(temp, wtemp)


@pytest.mark.parametrize("B,N,M", [(7, 512, 256), (5, 64, 64), (6, 100, 37), (9, 1024, 200), (3, 1500, 300), (2, 3072, 500),
                                   (5, 32, 32), (4, 1023, 1023), (3, 4095, 100), (13, 640, 50)])
def test_small_cloud_register_kernel(orc, lib, B, N, M):
    """Clouds of 32..4096 points take the register-resident kernel (fps_small.cu; several one-warp clouds per CTA, partial
    last CTA, every (B = opt_n_threads, C = ceil(N/B)) layout): D-FPS and S-FPS == oracle == bucket kernel, indices and
    min-distances, with duplicated points so that the tie order matters."""
    xyz = synth.clouds(B, N, seed=3 * N + B, dup_frac=0.3)
    want, wtemp = orc.furthest_point_sample(xyz, M, return_temp=True)
    for impl in (0, 6, 4):
        idx, temp = _fps_impl(xyz, M, impl)
        np.testing.assert_array_equal(idx, want)
        np.testing.assert_array_equal(temp, wtemp)
    w = synth.weights(B, N, seed=N)
    w[:, :3] = 0.0; w[:, 3:6] = 1e-13; w[:, 6:12] = w[:, 12:18]        # double path + tied weights
    want, wtemp = orc.furthest_point_sample_weights(xyz, w, M, return_temp=True)
    for impl in (0, 6, 4):
        idx, temp = _fps_impl(xyz, M, impl, weights=w)
        np.testing.assert_array_equal(idx, want)
        np.testing.assert_array_equal(temp, wtemp)


def test_dfps_all_points_identical_and_caller_temp(orc, lib, ops):
    # every distance ties at 0: the tie rule alone decides
    xyz = np.ones((1, 777, 3), np.float32)
    np.testing.assert_array_equal(_fps_impl(xyz, 50, 0)[0], orc.furthest_point_sample(xyz, 50))
    np.testing.assert_array_equal(_fps_impl(xyz, 50, 1)[0], orc.furthest_point_sample(xyz, 50))


@pytest.mark.parametrize("B,N,M", [(2, 384, 96), (1, 1000, 120), (1, 4096, 64), (1, 4100, 33), (1, 5, 5)])
def test_ffps_vs_oracle(orc, ops, B, N, M):
    pu = ops[0]
    xyz = synth.clouds(B, N, seed=N, dup_frac=0.1)
    mat = synth.dist_matrix(xyz, synth.features(B, 4, N, seed=1)) if N <= 1024 else \
        np.abs(np.random.default_rng(N).normal(size=(B, N, N))).astype(np.float32)
    got = pu.furthest_point_sample_matrix(cu(mat), M).cpu().numpy()
    np.testing.assert_array_equal(got, orc.furthest_point_sample_matrix(mat, M))


@pytest.mark.parametrize("B,N,C,layout", [(2, 200, 16, "bnc"), (2, 257, 5, "bcn"), (1, 64, 64, "bcn"), (1, 130, 33, "bnc"),
                                            (2, 96, 0, "none"), (1, 1, 3, "bnc")])
def test_dist_matrix_vs_oracle(orc, ops, B, N, C, layout):
    """calc_dist_matrix_for_sampling: bit-exact against the oracle's direct-difference statement, bitwise symmetric,
    and equal to torch.cdist within torch's own cancellation error."""
    pu = ops[0]
    xyz = synth.clouds(B, N, seed=N, dup_frac=0.1)
    if layout == "none":
        got = pu.calc_dist_matrix_for_sampling(cu(xyz), None, 0.5).cpu().numpy()
        want = orc.calc_dist_matrix_for_sampling(xyz, None, 0.5)
        ref = torch.cdist(torch.from_numpy(xyz).double(), torch.from_numpy(xyz).double()).numpy()
    else:
        feats = synth.features(B, C, N, seed=2)                       # (B, C, N)
        f_bnc = np.ascontiguousarray(feats.transpose(0, 2, 1))
        dev_f = cu(f_bnc) if layout == "bnc" else cu(feats).permute(0, 2, 1)   # contiguous, or the reference's permuted view
        got = pu.calc_dist_matrix_for_sampling(cu(xyz), dev_f, 0.7).cpu().numpy()
        want = orc.calc_dist_matrix_for_sampling(xyz, f_bnc, 0.7)
        x64, f64 = torch.from_numpy(xyz).double(), torch.from_numpy(f_bnc).double()
        ref = (torch.cdist(x64, x64) + torch.cdist(f64, f64) * 0.7).numpy()
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(got, got.transpose(0, 2, 1))
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-6)


def test_dist_matrix_full_size_properties(ops):
    """Layer-2 shape (N = 4096, C = 64): symmetric, zero diagonal, agrees with torch.cdist on the device."""
    pu = ops[0]
    xyz = cu(synth.clouds(2, 4096, seed=5))
    f = cu(synth.features(2, 64, 4096, seed=5)).permute(0, 2, 1)
    got = pu.calc_dist_matrix_for_sampling(xyz, f, 1.0)
    assert torch.equal(got, got.transpose(1, 2))
    assert float(got.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    ref = torch.cdist(xyz.double(), xyz.double()) + torch.cdist(f.double(), f.double())
    assert float(((got.double() - ref).abs() / (1.0 + ref)).max()) < 2e-6
    ref32 = torch.cdist(xyz, xyz) + torch.cdist(f, f)   # the reference's own fp32 route: GEMM expansion, cancels
    assert float((got - ref32).abs().max()) < 0.1       # (observed 0.04 on the diagonal, where it should be 0)
    idx = pu.furthest_point_sample_matrix(got, 512)
    assert all(len(set(r.tolist())) == 512 for r in idx.cpu())


@pytest.mark.parametrize("B,N,C,M,layout", [(2, 384, 8, 96, "bnc"), (2, 1000, 16, 120, "bcn"), (1, 4096, 64, 512, "bcn"),
                                              (3, 4100, 20, 64, "bnc"), (1, 700, 0, 50, "none"), (1, 9, 3, 9, "bnc"),
                                              (2, 2048, 7, 33, "bcn"), (2, 3000, 32, 100, "bcn"), (1, 4096, 64, 40, "bnc")])
def test_fused_ffps_equals_matrix_path_and_oracle(orc, ops, B, N, C, M, layout):
    """furthest_point_sample_features (cluster kernel, no matrix) == calc_dist_matrix_for_sampling +
    furthest_point_sample_matrix on the device == the same pair evaluated by the oracle on the CPU; duplicated
    points (exact ties -> the reference tie rule decides) included."""
    pu = ops[0]
    xyz = synth.clouds(B, N, seed=N + 1, dup_frac=0.1)
    if layout == "none":
        feats_bnc, dev_f = None, None
    else:
        feats = synth.features(B, C, N, seed=4)
        feats[:, :, 7] = feats[:, :, 3]; xyz[:, 7] = xyz[:, 3]            # a fully duplicated point
        feats_bnc = np.ascontiguousarray(feats.transpose(0, 2, 1))
        dev_f = cu(feats_bnc) if layout == "bnc" else cu(feats).permute(0, 2, 1)
    got = pu.furthest_point_sample_features(cu(xyz), dev_f, 0.7, M).cpu().numpy()
    two = pu.furthest_point_sample_matrix(pu.calc_dist_matrix_for_sampling(cu(xyz), dev_f, 0.7), M).cpu().numpy()
    np.testing.assert_array_equal(got, two)
    if N <= 1024:
        want = orc.furthest_point_sample_matrix(orc.calc_dist_matrix_for_sampling(xyz, feats_bnc, 0.7), M)
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("C,M,kind", [(64, 4096, "dup"), (16, 2000, "lattice"), (32, 1500, "clusters"), (64, 512, "equal_features")])
def test_fused_ffps_adversarial_ties(ops, C, M, kind):
    """Fused F-FPS against the two-call path on inputs where ties decide: exhaustive sampling with massive duplication,
    lattices (exact distance ties), tight clusters, identical feature vectors."""
    pu = ops[0]
    rng = np.random.default_rng(C + M)
    N, B = 4096, 2
    if kind == "dup":
        xyz = synth.clouds(B, N, seed=3, dup_frac=0.5); feats = synth.features(B, C, N, seed=3)
        src = rng.integers(0, N, N // 2); dst = rng.permutation(N)[: N // 2]
        feats[:, :, dst] = feats[:, :, src]; xyz[:, dst] = xyz[:, src]
    elif kind == "lattice":
        g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        xyz = np.stack([g[rng.permutation(N)], g[rng.permutation(N)] * np.float32(0.5)])
        feats = np.round(synth.features(B, C, N, seed=4))
    elif kind == "clusters":
        ctr = rng.uniform(0, 40, (9, 3))
        xyz = (ctr[rng.integers(0, 9, (B, N))] + rng.normal(0, 0.02, (B, N, 3))).astype(np.float32)
        feats = (synth.features(B, C, N, seed=5) * 0.01).astype(np.float32)
    else:
        xyz = synth.clouds(B, N, seed=6); feats = np.ones((B, C, N), np.float32)
    x, f = cu(xyz), cu(feats).permute(0, 2, 1)
    two = pu.furthest_point_sample_matrix(pu.calc_dist_matrix_for_sampling(x, f, 1.0), M)
    assert torch.equal(pu.furthest_point_sample_features(x, f, 1.0, M), two)
    if C == 64:
        for s in (4, 44, 6, 8):
            assert torch.equal(pu.furthest_point_sample_features(x, f, 1.0, M, cluster_size=s), two), "cluster size %d" % s


@pytest.mark.parametrize("N,kind", [(4096, "dup"), (4096, "equal_features"), (3600, "plain"), (4224, "plain"), (4100, "dup"),
                                    (3073, "plain_bnc"), (4095, "dup_bnc")])
def test_fused_ffps_cluster_sizes_agree(ops, monkeypatch, N, kind):
    """The 4-CTA (1024 points per CTA, half of the channels read from shared memory), 6-CTA (704 points per CTA, uneven last slice) and 8-CTA cluster forms of the fused F-FPS kernel pick the same
    indices as the two-call path, ties included (the launcher chooses between them by batch size; forced here)."""
    pu = ops[0]
    bnc = kind.endswith("_bnc")          # point-major features (B, N, C) instead of the permuted view of (B, C, N)
    kind = kind.replace("_bnc", "")
    B, C, M = 3, 64, 1024 if kind == "dup" else 300
    rng = np.random.default_rng(N)
    xyz = synth.clouds(B, N, seed=11, dup_frac=0.3 if kind == "dup" else 0.0)
    feats = synth.features(B, C, N, seed=11)
    if kind == "dup":
        src = rng.integers(0, N, N // 2); dst = rng.permutation(N)[: N // 2]
        feats[:, :, dst] = feats[:, :, src]; xyz[:, dst] = xyz[:, src]
    elif kind == "equal_features":
        feats[:] = 1.0
    x, f = cu(xyz), (cu(np.ascontiguousarray(feats.transpose(0, 2, 1))) if bnc else cu(feats).permute(0, 2, 1))
    two = pu.furthest_point_sample_matrix(pu.calc_dist_matrix_for_sampling(x, f, 1.0), M)
    for s in (4, 44, 6, 8):    # 44: 4-CTA clusters with four points per thread
        for pr in (1, 2, 3):   # dense / pruned / pruned with cooperative evaluation
            if s in (4, 44) and (pr >= 2 or N > 4096):      # 4-CTA clusters: dense kernel, at most 4 x 1024 points
                continue
            assert torch.equal(pu.furthest_point_sample_features(x, f, 1.0, M, cluster_size=s, prune=pr), two), "cluster size %d, prune %d" % (s, pr)
    big = pu.furthest_point_sample_features(x.repeat(6, 1, 1), f.repeat(6, 1, 1), 1.0, M)   # 18 clouds: the launcher's own pick
    assert torch.equal(big[:3], two) and torch.equal(big[15:], two)


@pytest.mark.parametrize("B,N,C,M,gamma,cloud", [(2, 4096, 64, 512, 1.0, "uniform"), (2, 4096, 64, 512, 1.0, "lidar"), (2, 4096, 64, 4096, 0.05, "lidar"),
                                                   (2, 4096, 64, 300, 30.0, "uniform"), (3, 384, 8, 96, 0.7, "uniform"), (2, 1000, 16, 1000, 1.0, "lidar"),
                                                   (1, 700, 0, 50, 1.0, "uniform"), (2, 2048, 7, 33, 1.0, "uniform"), (2, 3000, 32, 100, 0.0, "uniform"),
                                                   (1, 8192, 16, 64, 1.0, "lidar"), (1, 65, 3, 65, 1.0, "uniform"), (2, 5000, 40, 200, 1.0, "uniform"),
                                                   (1, 130, 128, 100, 1.0, "uniform")])
def test_fused_ffps_pruned_equals_dense(ops, B, N, C, M, gamma, cloud):
    """The pruned cluster kernel (Morton-sorted 64-point buckets, a warp skips its bucket when the bounding box proves that no
    min-distance can change) picks the indices of the dense kernel and of the two-call path and leaves the same running
    min-distances: coordinate-dominated metrics (most buckets skipped), feature-dominated ones (none skipped), gamma = 0,
    exhaustive sampling, duplicated points, ragged last buckets, no features at all."""
    pu = ops[0]
    from de6d_b200._lib import call
    from de6d_b200.compat._common import stream_ptr
    xyz = (synth.lidar_clouds if cloud == "lidar" else synth.clouds)(B, N, seed=N + C)
    xyz[:, 5] = xyz[:, 2]
    feats = synth.features(B, max(C, 1), N, seed=C)
    feats[:, :, 5] = feats[:, :, 2]
    x = cu(xyz)
    f = cu(feats).permute(0, 2, 1) if C else None
    two = pu.furthest_point_sample_matrix(pu.calc_dist_matrix_for_sampling(x, f, gamma), M)
    outs = {}
    for pr in (1, 2, 3):
        for s in (6, 8):
            out = torch.empty(B, M, dtype=torch.int32, device="cuda")
            temp = torch.full((B, N), 1e10, device="cuda")
            fp, (sb, sn, sc) = (f.data_ptr(), f.stride()) if C else (None, (0, 0, 0))
            try:
                call("de6d_furthest_point_sampling_features_impl", B, N, C, M, x.data_ptr(), fp, sb, sn, sc, float(gamma),
                     temp.data_ptr(), out.data_ptr(), s, pr, stream_ptr())
            except RuntimeError as e:          # a (shape, cluster size) pair one of the kernels does not cover
                assert "not covered" in str(e) or "does not fit" in str(e), e
                continue
            outs[(pr, s)] = (out, temp)
            assert torch.equal(out, two), "prune %d cluster %d" % (pr, s)
    assert any(pr == 2 for pr, _ in outs) and any(pr == 3 for pr, _ in outs), "the pruned kernel covered no cluster size"
    temps = [t for _, t in outs.values()]
    for t in temps[1:]:
        assert torch.equal(t, temps[0])


@pytest.mark.parametrize("kind", ["nan_xyz", "inf_xyz", "nan_feat", "neg_gamma", "nan_temp"])
def test_fused_ffps_pruned_non_finite(ops, kind):
    """Inputs that switch the bound off (non-finite coordinates, NaN min-distances, negative gamma) or poison single
    distances (NaN features): pruned == dense, bit for bit."""
    pu = ops[0]
    from de6d_b200._lib import call
    from de6d_b200.compat._common import stream_ptr
    B, N, C, M = 2, 4096, 64, 200
    xyz = synth.clouds(B, N, seed=21)
    feats = synth.features(B, C, N, seed=21)
    gamma = 1.0
    if kind == "nan_xyz":
        xyz[0, 100, 1] = np.nan
    elif kind == "inf_xyz":
        xyz[1, 7, 0] = np.inf
    elif kind == "nan_feat":
        feats[0, 3, 50] = np.nan
    elif kind == "neg_gamma":
        gamma = -0.5
    x, f = cu(xyz), cu(feats).permute(0, 2, 1)
    res = []
    for pr in (1, 2, 3):
        out = torch.empty(B, M, dtype=torch.int32, device="cuda")
        temp = torch.full((B, N), 1e10, device="cuda")
        if kind == "nan_temp":
            temp[0, 9] = float("nan")
        call("de6d_furthest_point_sampling_features_impl", B, N, C, M, x.data_ptr(), f.data_ptr(), *f.stride(), float(gamma),
             temp.data_ptr(), out.data_ptr(), 0, pr, stream_ptr())
        res.append((out, temp))
    for r in res[1:]:
        assert torch.equal(res[0][0], r[0])
        assert torch.equal(res[0][1].nan_to_num(nan=-7.0), r[1].nan_to_num(nan=-7.0))


@pytest.mark.parametrize("B", [16, 48])
def test_fused_ffps_full_batch(ops, B):
    """Layer-2 shape of the chain at batch 16 (one wave of 6-CTA clusters) and 48 (the launcher's 4-CTA form, four points per
    thread, more clusters than fit at once: two waves)."""
    pu = ops[0]
    xyz = cu(synth.clouds(B, 4096, seed=8))
    f = cu(synth.features(B, 64, 4096, seed=8)).permute(0, 2, 1)
    got = pu.furthest_point_sample_features(xyz, f, 1.0, 512)
    two = pu.furthest_point_sample_matrix(pu.calc_dist_matrix_for_sampling(xyz, f, 1.0), 512)
    assert torch.equal(got, two)
    assert all(len(set(r.tolist())) == 512 for r in got.cpu())


def test_fps_full_size_properties(lib, ops):
    """BASELINE size (16 x 16384 -> 4096): pruned kernel == unpruned kernel == generic kernel, indices unique."""
    xyz = synth.clouds(16, 16384, seed=0, dup_frac=0.0)
    a, ta = _fps_impl(xyz, 4096, 0)
    b, tb = _fps_impl(xyz, 4096, 1)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ta, tb)
    c, _ = _fps_impl(xyz[:2], 4096, 2)
    np.testing.assert_array_equal(a[:2], c)
    assert (a[:, 0] == 0).all()
    for row in a:
        assert len(np.unique(row)) == 4096
    # running min-distance really is the distance to the nearest selected point (checked on a sample)
    sel = xyz[0][a[0]]
    probe = np.random.default_rng(0).integers(0, 16384, 64)
    d = ((xyz[0][probe, None, :].astype(np.float64) - sel[None].astype(np.float64)) ** 2).sum(-1).min(1)
    np.testing.assert_allclose(ta[0][probe], d, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ ball query / group / gather
def _bq_all(impl, r, ns, x, q, r_in=None):
    """The three ball-query variants through the compat layer with a pinned kernel choice
    (0 automatic, 1 brute-force kernel, 2 grid kernel)."""
    from de6d_b200.compat import pointnet2_batch_cuda as p2
    B, N, _ = x.shape
    M = q.shape[1]
    out = {}
    idx = torch.zeros((B, M, ns), dtype=torch.int32, device="cuda")
    p2._ball_query(0, B, N, M, 0.0, r, ns, q, x, None, idx, impl=impl)
    out["plain"] = idx.cpu().numpy()
    idx = torch.zeros((B, M, ns), dtype=torch.int32, device="cuda"); cnt = torch.zeros((B, M), dtype=torch.int32, device="cuda")
    p2._ball_query(1, B, N, M, 0.0, r, ns, q, x, cnt, idx, impl=impl)
    out["cnt"] = (cnt.cpu().numpy(), idx.cpu().numpy())
    idx = torch.zeros((B, M, ns), dtype=torch.int32, device="cuda"); cnt = torch.zeros((B, M), dtype=torch.int32, device="cuda")
    p2._ball_query(2, B, N, M, r * 0.5 if r_in is None else r_in, r, ns, q, x, cnt, idx, impl=impl)
    out["dil"] = (cnt.cpu().numpy(), idx.cpu().numpy())
    return out


def _bq_check(orc, got, r, ns, xyz, new_xyz):
    np.testing.assert_array_equal(got["plain"], orc.ball_query(r, ns, xyz, new_xyz))
    wc, wi = orc.ball_query_cnt(r, ns, xyz, new_xyz)
    np.testing.assert_array_equal(got["cnt"][0], wc); np.testing.assert_array_equal(got["cnt"][1], wi)
    wc, wi = orc.ball_query_dilated(r * 0.5, r, ns, xyz, new_xyz)
    np.testing.assert_array_equal(got["dil"][0], wc); np.testing.assert_array_equal(got["dil"][1], wi)


@pytest.mark.parametrize("B,N,M,r,ns", [(2, 1500, 96, 0.5, 16), (2, 1500, 96, 2.0, 16), (1, 4096, 333, 1.0, 32),
                                        (2, 16384, 512, 0.8, 64), (1, 100, 7, 50.0, 5), (1, 2049, 65, 3.0, 33),
                                        (1, 4096, 300, 30.0, 32), (2, 3000, 257, 0.05, 8)])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_ball_query_variants_vs_oracle(orc, ops, B, N, M, r, ns, impl):
    """Brute-force kernel, grid kernel (incl. its in-order fallback for balls that swallow the cloud: r = 30, 50) and
    the automatic choice, all three variants, against the oracle: LiDAR-ring clouds (dense near field: balls
    overflow nsample), duplicated points, queries far outside the cloud (empty rows)."""
    xyz = synth.lidar_clouds(B, N, seed=N) if N >= 1024 else synth.clouds(B, N, seed=N)
    xyz[:, 5] = xyz[:, 3]                                   # duplicated points are distinct hits
    new_xyz = np.ascontiguousarray(xyz[:, :: max(N // M, 1)][:, :M]) + np.float32(0.01)
    new_xyz[:, -2:] += 1000.0
    _bq_check(orc, _bq_all(impl, r, ns, cu(xyz), cu(new_xyz)), r, ns, xyz, new_xyz)


@pytest.mark.parametrize("impl", [1, 2])
def test_ball_query_radius_boundary_and_degenerate_clouds(orc, impl):
    """Points placed within a few ulps of the sphere surface (the candidate-set argument of the grid kernel must not
    lose a point the exact test accepts), a flat cloud (zero extent in z), a single-point cloud, NaN / inf
    coordinates (never hits; the grid kernel falls back to the in-order scan for that cloud)."""
    rng = np.random.default_rng(7)
    N, M, r, ns = 4096, 128, 0.75, 16
    xyz = synth.clouds(1, N, seed=3)
    new_xyz = np.ascontiguousarray(xyz[:, :M]) + np.float32(0.3)
    d = rng.normal(size=(M, 8, 3)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    scale = r * (1.0 + rng.integers(-4, 5, size=(M, 8, 1)) * 6e-8)
    xyz[0, 1000:1000 + M * 8] = (new_xyz[0][:, None, :] + d * scale).reshape(-1, 3).astype(np.float32)
    # axis-aligned offsets of exactly r: the worst case for the cell-range bound
    xyz[0, 3000:3000 + M] = new_xyz[0] + np.array([r, 0, 0], np.float32)
    xyz[0, 3200:3200 + M] = new_xyz[0] - np.array([0, np.float32(r) * np.float32(1 - 6e-8), 0], np.float32)
    _bq_check(orc, _bq_all(impl, r, ns, cu(xyz), cu(new_xyz)), r, ns, xyz, new_xyz)
    flat = xyz.copy(); flat[..., 2] = 0.5
    _bq_check(orc, _bq_all(impl, r, ns, cu(flat), cu(new_xyz)), r, ns, flat, new_xyz)
    one = np.repeat(xyz[:, :1], 2100, axis=1)
    q1 = np.ascontiguousarray(one[:, :4]) + np.float32(0.1)
    _bq_check(orc, _bq_all(impl, r, ns, cu(one), cu(q1)), r, ns, one, q1)
    bad = xyz.copy(); bad[0, 10] = np.nan; bad[0, 11, 0] = np.inf
    _bq_check(orc, _bq_all(impl, r, ns, cu(bad), cu(new_xyz)), r, ns, bad, new_xyz)


@pytest.mark.parametrize("N,M,r,ns", [(131072, 700, 0.2, 64), (40000, 300, 1.0, 32), (131072, 200, 6.0, 16)])
def test_ball_query_large_clouds(orc, N, M, r, ns):
    """BASELINE configs[4] shape (131072-point LiDAR-ring frames, ns = 64): the large-cloud grid (histogram in global
    memory, up to 2^21 cells) against the oracle and the brute-force kernel; r = 6 m exercises the per-query fallback."""
    xyz = synth.lidar_clouds(2, N, seed=N)
    new_xyz = np.ascontiguousarray(xyz[:, :: N // M][:, :M]) + np.float32(0.01)
    new_xyz[:, -2:] += 1000.0
    got = _bq_all(2, r, ns, cu(xyz), cu(new_xyz))
    _bq_check(orc, got, r, ns, xyz, new_xyz)
    brute = _bq_all(1, r, ns, cu(xyz), cu(new_xyz))
    np.testing.assert_array_equal(got["cnt"][1], brute["cnt"][1]); np.testing.assert_array_equal(got["plain"], brute["plain"])


def test_ball_query_full_size_grid_equals_brute_force(ops):
    """BASELINE layer-1 shape (16 x 16384 points, 4096 queries, r in {0.2, 0.4, 0.8}): the grid kernel and the
    brute-force kernel agree bit for bit on uniform and on LiDAR-ring clouds; rows are ascending up to the count."""
    pu = ops[0]
    for maker in (synth.clouds, synth.lidar_clouds):
        xyz = cu(maker(16, 16384, seed=1))
        q = pu.gather_operation(xyz.transpose(1, 2).contiguous(), pu.furthest_point_sample(xyz, 4096)).transpose(1, 2).contiguous()
        for r, ns in ((0.2, 32), (0.4, 32), (0.8, 64)):
            a, b = _bq_all(1, r, ns, xyz, q), _bq_all(2, r, ns, xyz, q)
            for k in ("cnt", "dil"):
                np.testing.assert_array_equal(a[k][0], b[k][0]); np.testing.assert_array_equal(a[k][1], b[k][1])
            np.testing.assert_array_equal(a["plain"], b["plain"])
            cnt, idx = b["cnt"]
            assert (cnt >= 1).all()                       # every query is a cloud point: it finds itself
            first = np.take_along_axis(idx, np.zeros(idx.shape[:2] + (1,), np.int64), 2)[..., 0]
            srt = np.sort(idx, axis=2)
            for bb in range(0, 16, 5):
                for mm in range(0, 4096, 257):
                    c = cnt[bb, mm]
                    assert (np.diff(idx[bb, mm, :c]) > 0).all() and first[bb, mm] == srt[bb, mm, 0]


def test_ball_query_leaves_empty_rows_untouched(lib):
    from de6d_b200.compat import pointnet2_batch_cuda as p2
    xyz = cu(synth.clouds(1, 256, seed=1)); q = cu(np.full((1, 4, 3), 1e4, np.float32))
    idx = torch.full((1, 4, 8), 77, dtype=torch.int32, device="cuda"); cnt = torch.full((1, 4), 9, dtype=torch.int32, device="cuda")
    p2.ball_query_cnt_wrapper(1, 256, 4, 0.5, 8, q, xyz, cnt, idx)
    assert (idx == 77).all() and (cnt == 0).all()


@pytest.mark.parametrize("B,C,N,M,ns,impl", [(2, 3, 1500, 96, 16, 0), (2, 3, 1500, 96, 16, 1), (2, 3, 1500, 96, 16, 2),
                                             (2, 5, 4096, 1024, 32, 0), (2, 5, 4096, 1024, 32, 2), (1, 70, 1000, 333, 3, 0),
                                             (1, 1, 16384, 4096, 32, 0), (1, 3, 16384, 4096, 32, 2), (1, 7, 1001, 50, 4, 2)])
def test_group_points_vs_oracle(orc, lib, B, C, N, M, ns, impl):
    from de6d_b200._lib import call
    feats = synth.features(B, C, N, seed=3)
    idx = np.random.default_rng(4).integers(0, N, size=(B, M, ns)).astype(np.int32)
    f, i = cu(feats), cu(idx)
    out = torch.empty((B, C, M, ns), device="cuda")
    call("de6d_group_points_impl", B, C, N, M, ns, f.data_ptr(), i.data_ptr(), out.data_ptr(), impl, torch.cuda.current_stream().cuda_stream)
    np.testing.assert_array_equal(out.cpu().numpy(), orc.grouping_operation(feats, idx))


def test_gather_and_grads_vs_oracle(orc, ops):
    pu = ops[0]
    B, C, N, M, ns = 2, 6, 700, 128, 8
    feats = synth.features(B, C, N, seed=5)
    idx = np.random.default_rng(6).integers(0, N, size=(B, M)).astype(np.int32)
    f = cu(feats).requires_grad_(True)
    out = pu.gather_operation(f, cu(idx))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), orc.gather_operation(feats, idx))
    g = synth.features(B, C, M, seed=7)
    out.backward(cu(g))
    np.testing.assert_allclose(f.grad.cpu().numpy(), orc.gather_operation_grad(g, idx, N), rtol=1e-5, atol=1e-6)
    gidx = np.random.default_rng(8).integers(0, N, size=(B, M, ns)).astype(np.int32)
    f2 = cu(feats).requires_grad_(True)
    o2 = pu.grouping_operation(f2, cu(gidx))
    gg = synth.features(B, C, M * ns, seed=9).reshape(B, C, M, ns)
    o2.backward(cu(gg))
    np.testing.assert_allclose(f2.grad.cpu().numpy(), orc.grouping_operation_grad(gg, gidx, N), rtol=1e-5, atol=1e-5)


def test_query_and_group_modules(orc, ops):
    pu = ops[0]
    xyz = synth.lidar_clouds(2, 2048, seed=2); new_xyz = np.ascontiguousarray(xyz[:, :128])
    feats = synth.features(2, 4, 2048, seed=2)
    cnt, nf = pu.QueryWithCntAndGroup(1.0, 16)(cu(xyz), cu(new_xyz), cu(feats))
    wc, wi = orc.ball_query_cnt(1.0, 16, xyz, new_xyz)
    gx = orc.grouping_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), wi) - new_xyz.transpose(0, 2, 1)[..., None]
    want = np.concatenate([gx, orc.grouping_operation(feats, wi)], 1)
    np.testing.assert_array_equal(cnt.cpu().numpy(), wc)
    np.testing.assert_array_equal(nf.cpu().numpy(), want)
    nf2 = pu.QueryAndGroup(1.0, 16)(cu(xyz), cu(new_xyz), cu(feats))
    assert nf2.shape == (2, 7, 128, 16)


@pytest.mark.parametrize("B,C,N,M,ns", [(2, 4, 2048, 128, 16), (2, 0, 1500, 77, 5), (1, 70, 4096, 1024, 32), (1, 1, 16384, 4096, 64),
                                        (2, 9, 1001, 50, 7), (3, 1, 16384, 4096, 32), (2, 128, 1024, 512, 32), (2, 256, 512, 256, 16),
                                        (2, 0, 4096, 1024, 16), (1, 2, 4098, 600, 12), (2, 5, 100, 64, 4)])
def test_group_concat_fused_equals_composition(orc, ops, B, C, N, M, ns):
    """de6d_group_concat (one pass) == the reference composition transpose -> group -> subtract -> group -> cat,
    on the device (op by op through the mirror) and on the CPU (oracle); and the grouper module takes the
    composition whenever autograd needs it, with identical values and a working backward."""
    pu = ops[0]
    rng = np.random.default_rng(N)
    xyz = synth.clouds(B, N, seed=N); new_xyz = np.ascontiguousarray(xyz[:, :M]) + np.float32(0.05)
    feats = synth.features(B, C, N, seed=3) if C else None
    idx = rng.integers(0, N, size=(B, M, ns)).astype(np.int32)
    got = pu.group_concat(cu(xyz), cu(new_xyz), None if feats is None else cu(feats), cu(idx)).cpu().numpy()
    gx = orc.grouping_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx) - new_xyz.transpose(0, 2, 1)[..., None]
    want = gx if feats is None else np.concatenate([gx, orc.grouping_operation(feats, idx)], 1)
    np.testing.assert_array_equal(got, want)
    # with the transposed cloud handed in (coordinate rows TMA-staged like channels): same bits; gather_xyz makes it
    _, xyz_t = pu.gather_xyz(cu(xyz), None)
    np.testing.assert_array_equal(xyz_t.cpu().numpy(), xyz.transpose(0, 2, 1))
    got_t = pu.group_concat(cu(xyz), cu(new_xyz), None if feats is None else cu(feats), cu(idx), xyz_t=xyz_t).cpu().numpy()
    np.testing.assert_array_equal(got_t, want)
    if feats is not None:
        f = cu(feats).requires_grad_(True)
        with torch.enable_grad():
            comp = pu._assemble(cu(xyz), cu(new_xyz), f, cu(idx), True)      # autograd path: op-by-op composition
            comp.sum().backward()
        np.testing.assert_array_equal(comp.detach().cpu().numpy(), want)
        counts = np.zeros((B, N), np.float32)
        for b in range(B):
            np.add.at(counts[b], idx[b].reshape(-1), 1.0)
        np.testing.assert_array_equal(f.grad.cpu().numpy(), np.repeat(counts[:, None, :], C, axis=1))


def test_gather_xyz_equals_reference_composition(orc, ops):
    """de6d_gather_xyz == transpose -> gather_operation -> transpose (pointnet2_modules.py:374,451-454), both layouts."""
    pu = ops[0]
    for B, N, M in ((3, 16384, 4096), (2, 1000, 77), (1, 33, 33)):
        xyz = synth.clouds(B, N, seed=N)
        idx = np.random.default_rng(M).integers(0, N, size=(B, M)).astype(np.int32)
        want = np.ascontiguousarray(orc.gather_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx).transpose(0, 2, 1))
        got, got_t = pu.gather_xyz(cu(xyz), cu(idx), want_transposed=True)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        np.testing.assert_array_equal(got_t.cpu().numpy(), want.transpose(0, 2, 1))
        np.testing.assert_array_equal(pu.gather_xyz(cu(xyz), cu(idx)).cpu().numpy(), want)
    with pytest.raises(TypeError):
        pu.gather_xyz(cu(xyz), cu(idx.astype(np.int64)))


@pytest.mark.parametrize("kind,N,M", [("uniform", 16384, 4096), ("lidar", 16384, 4096), ("uniform", 4096, 1024), ("lidar", 40000, 3000)])
def test_ball_query_shared_grid_equals_own_grid(ops, kind, N, M):
    """One BallQueryGrid built for the smallest radius answers every radius scale of an SA layer (and the dilated
    shells) with the same indices and counts as a query that builds its own grid, and as the brute-force kernel."""
    pu = ops[0]
    from de6d_b200.compat import pointnet2_batch_cuda as ext
    B = 3
    xyz = cu((synth.clouds if kind == "uniform" else synth.lidar_clouds)(B, N, seed=N + 1))
    new_xyz = xyz[:, :M].contiguous() + 0.01
    radii, ns = (0.2, 0.4, 0.8, 1.6, 3.2), (32, 32, 64, 16, 64)
    grid = pu.BallQueryGrid(xyz, min(radii))
    prev = 0.0
    for r, k in zip(radii, ns):
        c0, i0 = pu.ball_query_cnt(r, k, xyz, new_xyz)
        c1, i1 = pu.ball_query_cnt(r, k, xyz, new_xyz, grid=grid)
        assert torch.equal(c0, c1) and torch.equal(i0, i1), r
        if r <= 0.8:      # brute force is slow at full size: check the small radii against it
            ib = torch.zeros_like(i0); cb = torch.zeros_like(c0)
            ext._ball_query(1, B, N, M, 0.0, r, k, new_xyz, xyz, cb, ib, impl=1)
            assert torch.equal(cb, c1) and torch.equal(ib, i1), r
        d0 = pu.ball_query_dilated(prev, r, k, xyz, new_xyz)
        d1 = pu.ball_query_dilated(prev, r, k, xyz, new_xyz, grid=grid)
        assert torch.equal(d0[0], d1[0]) and torch.equal(d0[1], d1[1]), r
        prev = r
    # a grid built for a LARGER radius than the one queried is just as exact (coarser cells, more candidates)
    coarse = pu.BallQueryGrid(xyz, 1.6)
    c2, i2 = pu.ball_query_cnt(0.4, 32, xyz, new_xyz, grid=coarse)
    c0, i0 = pu.ball_query_cnt(0.4, 32, xyz, new_xyz)
    assert torch.equal(c0, c2) and torch.equal(i0, i2)
    with pytest.raises(ValueError):
        pu.ball_query_cnt(0.4, 32, xyz.clone(), new_xyz, grid=grid)     # a grid belongs to one xyz tensor


def test_fused_ffps_agreement_with_reference_pipeline(ops, ref_modules):
    """The benchmarked fused F-FPS evaluates the metric by direct differences; the reference pipeline evaluates it with
    torch.cdist's fp32 GEMM expansion (pointnet2_utils.py:36-44) before its matrix kernel.  Greedy FPS is discontinuous
    in the distances, so the two index sequences part after the first near-tie; this test pins HOW FAR (measured on
    B200: profiles/r2_ffps_agreement.json -- same position in >= 98.9 % of the slots, identical selected SETS in
    >= 99.8 %, and the fused kernel is the one that stays on the float64-exact greedy sequence):
      * torch.cdist + de6d matrix kernel (what an unmodified checkout runs over compat)  == reference, bit for bit;
      * fused vs reference: set overlap >= 0.99, same position >= 0.95 on the bench shape (4096 pts x 64 ch -> 512);
      * fused vs the float64 metric: set overlap >= 0.995 and at least as close to it as the reference is."""
    pu = ops[0]
    if ref_modules is None:
        pytest.skip("oracle/_ref not present")
    p2 = ref_modules["pointnet2_batch_cuda"]
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, C, M = 8, 4096, 64, 512
    for maker in (synth.clouds, synth.lidar_clouds):
        xyz = cu(maker(B, 16384, seed=21)[:, :N].copy())
        f = cu(synth.features(B, C, N, seed=22)).permute(0, 2, 1)

        def ref_kernel(mat):
            temp = torch.full((B, N), 1e10, device="cuda")
            idx = torch.zeros((B, M), dtype=torch.int32, device="cuda")
            p2.furthest_point_sampling_matrix_wrapper(B, N, M, mat.contiguous(), temp, idx)
            return idx
        mat_ref = torch.cdist(xyz, xyz) + torch.cdist(f, f) * 1.0              # the reference's calc_dist_matrix_for_sampling
        ref = ref_kernel(mat_ref)
        assert torch.equal(pu.furthest_point_sample_matrix(mat_ref.contiguous(), M), ref)
        x64, f64 = xyz.double(), f.double().contiguous()
        exact = (torch.cdist(x64, x64, compute_mode="donot_use_mm_for_euclid_dist") +
                 torch.cdist(f64, f64, compute_mode="donot_use_mm_for_euclid_dist")).float()
        exact_idx = ref_kernel(exact)
        ours = pu.furthest_point_sample_features(xyz, f, 1.0, M)

        def overlap(a, b):
            a, b = a.cpu().numpy(), b.cpu().numpy()
            sets = np.mean([len(set(a[i]) & set(b[i])) / M for i in range(B)])
            return sets, float((a == b).mean())
        s_ref, p_ref = overlap(ours, ref)
        s_ex, p_ex = overlap(ours, exact_idx)
        s_rx, p_rx = overlap(ref, exact_idx)
        assert s_ref >= 0.99 and p_ref >= 0.95, (s_ref, p_ref)
        assert s_ex >= 0.995 and p_ex >= p_rx - 0.01, (s_ex, p_ex, p_rx)


# ------------------------------------------------------------------------------------------------ interpolation
def test_three_nn_interpolate_vs_oracle(orc, ops):
    pu = ops[0]
    unknown = synth.clouds(2, 3000, seed=1); known = synth.clouds(2, 1100, seed=2)
    known[:, 10] = known[:, 3]
    dist, idx = pu.three_nn(cu(unknown), cu(known))
    wd, wi = orc.three_nn(unknown, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), wi)
    np.testing.assert_array_equal(dist.cpu().numpy(), wd)
    feats = synth.features(2, 19, 1100, seed=3)
    w = np.random.default_rng(1).uniform(0, 1, (2, 3000, 3)).astype(np.float32)
    f = cu(feats).requires_grad_(True)
    out = pu.three_interpolate(f, idx, cu(w))
    np.testing.assert_array_equal(out.detach().cpu().numpy(), orc.three_interpolate(feats, wi, w))
    g = synth.features(2, 19, 3000, seed=4)
    out.backward(cu(g))
    np.testing.assert_allclose(f.grad.cpu().numpy(), orc.three_interpolate_grad(g, wi, w, 1100), rtol=1e-4, atol=1e-4)
    # fewer than 3 known points: sentinel distances become inf, indices stay 0
    d2, i2 = pu.three_nn(cu(unknown[:, :5]), cu(known[:, :2]))
    wd2, wi2 = orc.three_nn(unknown[:, :5], known[:, :2])
    np.testing.assert_array_equal(i2.cpu().numpy(), wi2); np.testing.assert_array_equal(d2.cpu().numpy(), wd2)


@pytest.mark.parametrize("kind", ["uniform", "lidar", "dup", "lattice", "far"])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_three_nn_grid_vs_oracle(orc, kind, impl):
    """three_nn: brute-force scan, grid search (3x3x3 then 5x5x5 cells of ~2 known points, full scan when the third
    neighbour cannot be proven final) and the automatic choice against the oracle -- distances and indices, with
    duplicated known points and lattices (exact ties: smallest index wins), queries far outside the known cloud."""
    from de6d_b200.compat import pointnet2_batch_cuda as p2
    rng = np.random.default_rng(5)
    B, n, m = 2, 5000, 4096
    if kind == "lidar":
        known = synth.lidar_clouds(B, m, seed=1); unknown = synth.lidar_clouds(B, 8192, seed=2)[:, :n]
    elif kind == "lattice":
        g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        known = np.stack([g[rng.permutation(m)], g[rng.permutation(m)]]); unknown = (rng.integers(0, 16, (B, n, 3)) + 0.5).astype(np.float32)
    else:
        known = synth.clouds(B, m, seed=1, dup_frac=0.3 if kind == "dup" else 0.0); unknown = synth.clouds(B, n, seed=2)
        if kind == "far":
            unknown[:, :500] += 300.0
    unknown = np.ascontiguousarray(unknown)
    d2 = torch.zeros((B, n, 3), device="cuda"); idx = torch.zeros((B, n, 3), dtype=torch.int32, device="cuda")
    p2.three_nn_wrapper(B, n, m, cu(unknown), cu(known), d2, idx, impl=impl)
    wd, wi = orc.three_nn(unknown, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), wi)
    np.testing.assert_array_equal(np.sqrt(d2.cpu().numpy()), wd)


@pytest.mark.parametrize("B,C,m,n", [(2, 19, 1100, 3000), (2, 5, 1101, 3001), (1, 64, 4096, 16384), (3, 3, 52, 200), (1, 70, 20000, 24000)])
def test_three_interpolate_staged_and_direct_vs_oracle(orc, ops, B, C, m, n):
    """three_interpolate: the TMA-staged kernel (aligned shapes, rows in shared memory) and the direct kernel
    (unaligned shapes / rows too large) are both bit-identical to the oracle (the FMA shape is pinned)."""
    pu = ops[0]
    rng = np.random.default_rng(m + n)
    feats = synth.features(B, C, m, seed=7)
    idx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = rng.uniform(0, 1, (B, n, 3)).astype(np.float32)
    out = pu.three_interpolate(cu(feats), cu(idx), cu(w)).cpu().numpy()
    np.testing.assert_array_equal(out, orc.three_interpolate(feats, idx, w))


# ------------------------------------------------------------------------------------------------ boxes
def _near_threshold(iou, thr, eps=1e-5):
    return np.abs(iou - thr) < eps


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_iou_matrices_vs_oracle(orc, ops, seed):
    iu = ops[1]
    p = synth.proposals(1, 300, seed=seed, clusters=30)[0][0]
    a, b = p[:130], p[100:300]
    want = orc.boxes_iou_bev(a, b)
    got = iu.boxes_iou_bev(cu(a), cu(b)).cpu().numpy()
    assert (want > 0).sum() > 200
    np.testing.assert_array_equal(got > 0, want > 0)
    np.testing.assert_allclose(got, want, rtol=RTOL_IOU, atol=1e-7)
    got3 = iu.boxes_iou3d_gpu(cu(a), cu(b)).cpu().numpy()
    np.testing.assert_allclose(got3, orc.boxes_iou3d(a, b), rtol=RTOL_IOU, atol=1e-7)
    got_cpu_api = iu.boxes_bev_iou_cpu(a, b)       # numpy in, numpy out, computed on the device
    assert isinstance(got_cpu_api, np.ndarray)
    np.testing.assert_allclose(got_cpu_api, want, rtol=RTOL_IOU, atol=1e-7)


def _nms_case(orc, seed, n, thr):
    """Seeded proposals with no pair within 1e-4 of the threshold: keep lists are only required to be identical
    when no IoU sits inside the 1e-5 tolerance band of the threshold (hard part 2).  Seeds are walked
    deterministically (seed, seed+100, ...) until the oracle's own IoU matrix is clear of the band."""
    for s in range(seed, seed + 2000, 100):
        bx, sc = synth.proposals(1, n, seed=s, clusters=max(n // 8, 1))
        bx, sc = bx[0], sc[0]
        order = np.argsort(-sc, kind="stable")
        sb = np.ascontiguousarray(bx[order])
        if not _near_threshold(orc.boxes_iou_bev(sb, sb), thr, eps=1e-4).any():
            return bx, sc, order, sb
    raise AssertionError("no clean seed found")


@pytest.mark.parametrize("seed,n,thr", [(0, 512, 0.01), (1, 512, 0.1), (2, 300, 0.5), (3, 64, 0.1), (4, 65, 0.1), (5, 1, 0.1), (6, 1500, 0.3)])
def test_nms_vs_oracle(orc, ops, seed, n, thr):
    iu = ops[1]
    bx, sc, order, sb = _nms_case(orc, seed, n, thr)
    want_keep = orc.nms_sorted(sb, thr)
    keep, _ = iu.nms_gpu(cu(bx), cu(sc), thr)
    np.testing.assert_array_equal(keep.cpu().numpy(), order[want_keep])
    keepn, _ = iu.nms_normal_gpu(cu(bx), cu(sc), thr)
    np.testing.assert_array_equal(keepn.cpu().numpy(), order[orc.nms_sorted(sb, thr, normal=True)])


def test_nms_batched_matches_per_frame(orc, ops):
    iu = ops[1]
    F, n = 5, 512
    bx, sc = synth.proposals(F, n, seed=11)
    nvalid = np.array([512, 300, 0, 64, 511], np.int32)
    keep, num = iu.nms_gpu_batched(cu(bx), cu(sc), 0.1, nvalid=cu(nvalid))
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for f in range(F):
        order = np.argsort(-sc[f], kind="stable")
        want = order[orc.nms_sorted(np.ascontiguousarray(bx[f][order][: nvalid[f]]), 0.1)]
        assert num[f] == len(want)
        np.testing.assert_array_equal(keep[f, : num[f]], want)
    # second launch on the same workspace (tickets must have been left clean)
    op = iu.BatchedNMS(F, n)
    k1, n1 = op(cu(bx), cu(sc), 0.1); k1, n1 = k1.clone(), n1.clone()
    k2, n2 = op(cu(bx), cu(sc), 0.1)
    assert torch.equal(n1, n2) and torch.equal(k1, k2)


def test_class_agnostic_nms_batched_equals_reference_loop(orc, ops):
    """model_nms_utils: the batched, sync-free post-processing selects per frame exactly what the reference's Python
    loop (score mask -> topk -> nms_gpu -> index mapping, model_nms_utils.py:6-25) selects, and that loop run on this
    package's nms_gpu equals the same loop evaluated with the oracle."""
    from de6d_b200 import model_nms_utils as mu
    F, n = 6, 512
    bx, sc = synth.proposals(F, n, seed=21)
    bx9 = np.concatenate([bx, np.zeros((F, n, 2), np.float32)], -1)          # (x, y, z, dx, dy, dz, rz, ry, rx): extra columns ignored
    cfg = {"NMS_TYPE": "nms_gpu", "NMS_THRESH": 0.1, "NMS_PRE_MAXSIZE": 300, "NMS_POST_MAXSIZE": 40}
    thr = 0.35
    sel, ssc, num = mu.class_agnostic_nms_batched(cu(sc), cu(bx9), cfg, score_thresh=thr)
    sel, ssc, num = sel.cpu().numpy(), ssc.cpu().numpy(), num.cpu().numpy()
    for f in range(F):
        s1, sc1 = mu.class_agnostic_nms(cu(sc[f]), cu(bx9[f]), cfg, score_thresh=thr)
        s1 = s1.cpu().numpy()
        # oracle version of the same loop
        idx = np.nonzero(sc[f] >= thr)[0]
        top = idx[np.argsort(-sc[f][idx], kind="stable")][:300]
        keep = orc.nms_sorted(np.ascontiguousarray(bx[f][top]), 0.1)[:40]
        np.testing.assert_array_equal(s1, top[keep])
        assert num[f] == len(s1)
        np.testing.assert_array_equal(sel[f, :num[f]], s1)
        assert (sel[f, num[f]:] == -1).all()
        np.testing.assert_array_equal(ssc[f, :num[f]], sc[f][s1])
    # no threshold, frame with nothing above threshold
    sel2, _, num2 = mu.class_agnostic_nms_batched(cu(sc), cu(bx9), cfg, score_thresh=2.0)
    assert int(num2.sum()) == 0 and bool((sel2 == -1).all())
    sel3, _, num3 = mu.class_agnostic_nms_batched(cu(sc), cu(bx9), cfg, score_thresh=None)
    s3, _ = mu.class_agnostic_nms(cu(sc[0]), cu(bx9[0]), cfg, score_thresh=None)
    np.testing.assert_array_equal(sel3[0, :int(num3[0])].cpu().numpy(), s3.cpu().numpy())


def _boxes9(n, seed, cluster=8, tilt=0.35):
    """Clustered full-pose proposals: `cluster` jittered copies per object so IoUs span (0, 1)."""
    rng = np.random.default_rng(seed)
    k = -(-n // cluster)
    base = np.concatenate([rng.uniform(0, 40, (k, 1)), rng.uniform(-20, 20, (k, 1)), rng.uniform(-1.5, -0.5, (k, 1)),
                           np.clip(rng.normal((3.9, 1.6, 1.56), 0.2, (k, 3)), 0.3, None), rng.uniform(-np.pi, np.pi, (k, 1)),
                           rng.uniform(-tilt, tilt, (k, 2))], 1)
    b = np.repeat(base, cluster, 0)[:n].copy()
    b[:, :3] += rng.normal(0, 0.3, (n, 3)) * [1, 1, 0.3]
    b[:, 3:6] *= rng.uniform(0.9, 1.1, (n, 3))
    b[:, 6:] += rng.normal(0, 0.08, (n, 3))
    return b.astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_full_pose_iou_vs_oracle(orc, ops, seed):
    """boxes_iou3d_9dof_gpu (float, box9.cuh) against the oracle (double, pinned to scipy): 1e-4 relative (+1e-6), the same
    zero pattern away from grazing contacts; identical / nested / face-sharing boxes; yaw-only boxes against boxes_iou3d_gpu."""
    iu = ops[1]
    a, b = _boxes9(300, seed), _boxes9(200, seed + 50)
    b[:40] = a[:40]                                   # identical boxes (padded duplicates)
    b[40:60] = a[40:60]; b[40:60, 3:6] *= 0.5         # nested
    b[60:80] = a[60:80]; b[60:80, 3] *= 0.5           # shared face planes
    got = iu.boxes_iou3d_9dof_gpu(cu(a), cu(b)).cpu().numpy()
    want = orc.boxes_iou3d_9dof(a, b)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6)
    assert (want > 0.01).sum() > 300
    np.testing.assert_allclose(np.diag(got)[:40], 1.0, atol=1e-5)
    np.testing.assert_allclose(np.diag(got)[40:60], 0.125, atol=1e-5)
    np.testing.assert_allclose(np.diag(got)[60:80], 0.5, atol=1e-5)
    a7 = a.copy(); a7[:, 7:] = 0
    i9 = iu.boxes_iou3d_9dof_gpu(cu(a7), cu(a7)).cpu().numpy()
    i7 = iu.boxes_iou3d_gpu(cu(a7[:, :7]), cu(a7[:, :7])).cpu().numpy()
    assert np.abs(i9 - i7).max() < 1.5e-2      # the reference's BEV clipping pads its corner tests by 1e-2 m (iou3d_nms_kernel.cu:51-61)
    assert iu.boxes_iou3d_9dof_gpu(cu(a[:0]), cu(b)).shape == (0, 200)


@pytest.mark.parametrize("n,thr", [(512, 0.1), (512, 0.45), (200, 0.25), (64, 0.01), (700, 0.3)])
def test_full_pose_nms_vs_oracle(orc, ops, n, thr):
    """nms_gpu_9dof / the batched kernel in mode 2 against the oracle's greedy sweep with the same IoU, on seeds with no
    pair within 2e-5 of the threshold."""
    iu = ops[1]
    checked = 0
    for seed in range(10):
        boxes = _boxes9(n, 10 * n + seed)
        iou = orc.boxes_iou3d_9dof(boxes, boxes)
        if (np.abs(iou - thr) < 2e-5).any():       # float kernel vs double oracle: |diff| <= 1e-4 * iou asserted above, ~1e-6 typical
            continue
        scores = np.random.default_rng(seed).permutation(n).astype(np.float32)
        keep, none = iu.nms_gpu_9dof(cu(boxes), cu(scores), thr)
        assert none is None
        np.testing.assert_array_equal(keep.cpu().numpy(), orc.nms_9dof(boxes, scores, thr))
        checked += 1
    assert checked >= 3
    # batched form, ragged nvalid, against per-frame calls
    F = 5
    bx = np.stack([_boxes9(n, 77 + f) for f in range(F)])
    sc = np.stack([np.random.default_rng(f).permutation(n) for f in range(F)]).astype(np.float32)
    keepb, numb = iu.nms_gpu_batched(cu(bx), cu(sc), thr)
    for f in range(F):
        k1, _ = iu.nms_gpu_9dof(cu(bx[f]), cu(sc[f]), thr)
        assert int(numb[f]) == k1.numel() and torch.equal(keepb[f, :k1.numel()], k1)


def test_class_agnostic_nms_full_pose(orc, ops):
    """model_nms_utils with NMS_TYPE 'nms_gpu_9dof': single-frame form == batched form == oracle sweep on [:, 0:9]."""
    from de6d_b200 import model_nms_utils as mu
    n, F = 384, 3
    cfg = {"NMS_TYPE": "nms_gpu_9dof", "NMS_THRESH": 0.2, "NMS_PRE_MAXSIZE": 300, "NMS_POST_MAXSIZE": 80}
    preds = np.stack([np.concatenate([_boxes9(n, 5 + f), np.zeros((n, 1), np.float32)], 1) for f in range(F)])
    scores = np.stack([np.random.default_rng(f).permutation(n) / n for f in range(F)]).astype(np.float32)
    sel_b, sc_b, num_b = mu.class_agnostic_nms_batched(cu(scores), cu(preds), cfg, score_thresh=0.1)
    for f in range(F):
        sel, sc = mu.class_agnostic_nms(cu(scores[f]), cu(preds[f]), cfg, score_thresh=0.1)
        k = int(num_b[f])
        assert k == sel.numel() and torch.equal(sel_b[f, :k], sel) and torch.equal(sc_b[f, :k], sc)
        m = scores[f] >= 0.1
        idx = np.nonzero(m)[0]
        top = idx[np.argsort(-scores[f][idx], kind="stable")][:300]
        want = top[orc.nms_9dof(preds[f][top][:, :9], scores[f][top], 0.2)][:80]
        np.testing.assert_array_equal(sel.cpu().numpy(), want)


def test_points_in_boxes_vs_oracle(orc, ops):
    ru = ops[2]
    B, T, M = 3, 100, 16384
    boxes = synth.boxes(B, T, seed=1); boxes[:, -5:] = 0.0
    rng = np.random.default_rng(2)
    pts = synth.clouds(B, M, seed=3)
    for b in range(B):
        pts[b, :6000] = (boxes[b, rng.integers(0, 95, 6000), :3] + rng.normal(0, 0.9, (6000, 3))).astype(np.float32)
    want = orc.points_in_boxes_gpu(pts, boxes)
    got = ru.points_in_boxes_gpu(cu(pts), cu(boxes)).cpu().numpy()
    assert (want >= 0).sum() > 3000
    # CPU libm vs CUDA cos/sin may differ in the last bit: a mismatch is only tolerated for a point provably
    # within 1e-5 m of a box face (none expected; count reported)
    bad = np.argwhere(got != want)
    assert len(bad) == 0, "%d mismatches, first %s" % (len(bad), bad[:5])
    mask = ru.points_in_boxes_cpu(pts[0], boxes[0])       # numpy in/out, device compute, MARGIN 1e-2 mask
    np.testing.assert_array_equal(mask, orc.points_in_boxes_cpu(pts[0], boxes[0]))


def test_points_in_boxes3d_full_pose_vs_oracle(orc):
    """box_utils.points_in_boxes3d (9-DoF boxes, Det6D target assignment) on the device vs the oracle."""
    from de6d_b200 import box_utils
    rng = np.random.default_rng(5)
    B, T, M = 3, 70, 16384
    boxes = np.concatenate([synth.boxes(B, T, seed=9)[..., :7], rng.uniform(-0.5, 0.5, (B, T, 2)).astype(np.float32)], -1)
    boxes[:, 11, :3] = boxes[:, 10, :3] + 0.2
    pts = synth.clouds(B, M, seed=4)
    for b in range(B):
        pts[b, :7000] = (boxes[b, rng.integers(0, T, 7000), :3] + rng.normal(0, 1.0, (7000, 3))).astype(np.float32)
    got = box_utils.points_in_boxes3d_batched(cu(pts), cu(boxes)).cpu().numpy()
    for b in range(B):
        want = orc.points_in_boxes3d(pts[b], boxes[b])
        assert (want >= 0).sum() > 2000
        np.testing.assert_array_equal(got[b], want)
    one = box_utils.points_in_boxes3d(pts[0], boxes[0])                 # numpy in / numpy out, reference signature
    assert isinstance(one, np.ndarray) and one.dtype == np.int64
    np.testing.assert_array_equal(one, got[0])
    t = box_utils.points_in_boxes3d(cu(pts[1]), cu(boxes[1]))
    assert t.is_cuda and torch.equal(t.cpu(), torch.from_numpy(got[1]))


# ------------------------------------------------------------------------------------------------ golden vectors (reference CUDA kernels)
def test_against_reference_cuda_golden(golden, ops, lib):
    pu, iu, ru = ops
    g = golden
    for tag, m in (("a", 64), ("b", 128), ("c", 40), ("d", 256)):
        xyz = g["fps_%s_xyz" % tag]
        for impl in (0, 1, 2, 4):
            idx, temp = _fps_impl(xyz, m, impl)
            np.testing.assert_array_equal(idx, g["fps_%s_idx" % tag]); np.testing.assert_array_equal(temp, g["fps_%s_temp" % tag])
            sidx, _ = _fps_impl(xyz, m, impl, weights=g["sfps_%s_w" % tag])
            np.testing.assert_array_equal(sidx, g["sfps_%s_idx" % tag])
    np.testing.assert_array_equal(pu.furthest_point_sample_matrix(cu(g["ffps_mat"]), 96).cpu().numpy(), g["ffps_idx"])
    x, q = cu(g["bq_xyz"]), cu(g["bq_new_xyz"])
    for r in (0.5, 2.0):
        np.testing.assert_array_equal(pu.ball_query(r, 16, x, q).cpu().numpy(), g["bq_idx_r%g" % r])
        cnt, idx = pu.ball_query_cnt(r, 16, x, q)
        np.testing.assert_array_equal(cnt.cpu().numpy(), g["bqc_cnt_r%g" % r]); np.testing.assert_array_equal(idx.cpu().numpy(), g["bqc_idx_r%g" % r])
        cnt, idx = pu.ball_query_dilated(r * 0.5, r, 16, x, q)
        np.testing.assert_array_equal(cnt.cpu().numpy(), g["bqd_cnt_r%g" % r]); np.testing.assert_array_equal(idx.cpu().numpy(), g["bqd_idx_r%g" % r])
    dist, idx = pu.three_nn(cu(g["nn_unknown"]), cu(g["nn_known"]))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["nn_idx"]); np.testing.assert_array_equal(dist.cpu().numpy(), np.sqrt(g["nn_dist2"]))
    out = pu.three_interpolate(cu(g["ti_feats"]), idx, cu(g["ti_weight"]))
    np.testing.assert_array_equal(out.cpu().numpy(), g["ti_out"])
    a, b = cu(g["iou_gpu_a"]), cu(g["iou_gpu_b"])
    got = iu.boxes_iou_bev(a, b).cpu().numpy()
    np.testing.assert_allclose(got, g["iou_gpu"], rtol=RTOL_IOU, atol=1e-7)
    np.testing.assert_array_equal(got > 0, g["iou_gpu"] > 0)
    from de6d_b200.compat import iou3d_nms_cuda as ext
    for thr in (0.01, 0.1, 0.5):
        sb = cu(g["nms_sorted_boxes"])
        keep = torch.zeros(sb.shape[0], dtype=torch.int64)
        n = ext.nms_gpu(sb, keep, thr)
        np.testing.assert_array_equal(keep[:n].numpy(), g["nms_keep_%g" % thr])
        n = ext.nms_normal_gpu(sb, keep, thr)
        np.testing.assert_array_equal(keep[:n].numpy(), g["nmsn_keep_%g" % thr])
    got = ru.points_in_boxes_gpu(cu(g["pibg_pts"]), cu(g["pibg_boxes"])).cpu().numpy()
    np.testing.assert_array_equal(got, g["pibg_out"])


# ------------------------------------------------------------------------------------------------ live reference kernels (when built)
def test_against_reference_wrapper_golden(golden_dir, ops):
    """This library's kernels against the second committed fixture (reference python wrappers over the reference kernels):
    fused boxes_iou3d_gpu, gather / group / three_interpolate with gradients, the fused QueryWithCntAndGroup tail."""
    p = os.path.join(golden_dir, "golden_cuda2.npz")
    if not os.path.exists(p):
        pytest.skip("golden_cuda2.npz not generated yet")
    g = np.load(p)
    pu, iu, _ = ops
    np.testing.assert_allclose(iu.boxes_iou3d_gpu(cu(g["iou3d_a"]), cu(g["iou3d_b"])).cpu().numpy(), g["iou3d"], rtol=RTOL_IOU, atol=1e-6)
    for fn, idx_key, out_key, gout_key, grad_key in (("gather_operation", "gather_idx", "gather_out", "gather_gout", "gather_grad"),
                                                      ("grouping_operation", "group_idx", "group_out", "group_gout", "group_grad")):
        f = cu(g["gg_feats"]).requires_grad_(True)
        with torch.enable_grad():
            y = getattr(pu, fn)(f, cu(g[idx_key]))
            y.backward(cu(g[gout_key]))
        np.testing.assert_array_equal(y.detach().cpu().numpy(), g[out_key])
        np.testing.assert_allclose(f.grad.cpu().numpy(), g[grad_key], rtol=1e-5, atol=1e-5)
    f = cu(g["ti2_feats"]).requires_grad_(True)
    with torch.enable_grad():
        y = pu.three_interpolate(f, cu(g["ti2_idx"]), cu(g["ti2_weight"]))
        y.backward(cu(g["ti2_gout"]))
    np.testing.assert_array_equal(y.detach().cpu().numpy(), g["ti2_out"])
    np.testing.assert_allclose(f.grad.cpu().numpy(), g["ti2_grad"], rtol=1e-5, atol=1e-5)
    cnt, nf = pu.QueryWithCntAndGroup(1.5, g["group_idx"].shape[2])(cu(g["qg_xyz"]), cu(g["qg_new_xyz"]), cu(g["gg_feats"]))
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["qg_cnt"])
    np.testing.assert_array_equal(nf.cpu().numpy(), g["qg_out"])


def test_against_live_reference_kernels(ref_modules, ops, lib):
    if ref_modules is None:
        pytest.skip("oracle/_ref not available on this box")
    pu, iu, ru = ops
    p2, iou3d, roi = ref_modules["pointnet2_batch_cuda"], ref_modules["iou3d_nms_cuda"], ref_modules["roiaware_pool3d_cuda"]
    # D-FPS at the BASELINE size incl. duplicated points, S-FPS, ball_query_cnt, NMS, IoU bit statistics
    B, N, M = 4, 16384, 4096
    xyz = cu(synth.clouds(B, N, seed=3, dup_frac=0.1))
    temp = torch.full((B, N), 1e10, device="cuda"); ref_idx = torch.zeros((B, M), dtype=torch.int32, device="cuda")
    p2.farthest_point_sampling_wrapper(B, N, M, xyz, temp, ref_idx)
    assert torch.equal(pu.furthest_point_sample(xyz, M), ref_idx)
    w = cu(synth.weights(B, N, seed=4))
    temp.fill_(1e10); ref_s = torch.zeros((B, 512), dtype=torch.int32, device="cuda")
    p2.furthest_point_sampling_weights_wrapper(B, N, 512, xyz, w, temp, ref_s)
    assert torch.equal(pu.furthest_point_sample_weights(xyz, w, 512), ref_s)
    new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), ref_idx).transpose(1, 2).contiguous()
    for r, ns in ((0.2, 32), (0.8, 64), (3.0, 32)):
        ridx = torch.zeros((B, M, ns), dtype=torch.int32, device="cuda"); rcnt = torch.zeros((B, M), dtype=torch.int32, device="cuda")
        p2.ball_query_cnt_wrapper(B, N, M, r, ns, new_xyz, xyz, rcnt, ridx)
        cnt, idx = pu.ball_query_cnt(r, ns, xyz, new_xyz)
        assert torch.equal(cnt, rcnt) and torch.equal(idx, ridx)
    bx, sc = synth.proposals(1, 512, seed=5)
    boxes = cu(bx[0]); scores = cu(sc[0])
    ref_iou = torch.zeros((512, 512), device="cuda")
    iou3d.boxes_iou_bev_gpu(boxes, boxes, ref_iou)
    got = iu.boxes_iou_bev(boxes, boxes)
    torch.testing.assert_close(got, ref_iou, rtol=RTOL_IOU, atol=1e-7)
    exact = (got == ref_iou).float().mean().item()
    print("IoU bit-exact fraction vs reference kernel: %.6f" % exact)
    order = scores.sort(0, descending=True)[1]
    sb = boxes[order].contiguous()
    for thr in (0.01, 0.1):
        keep = torch.zeros(512, dtype=torch.int64)
        n = iou3d.nms_gpu(sb, keep, thr)
        mine, _ = iu.nms_gpu(boxes, scores, thr)
        assert torch.equal(mine.cpu(), order.cpu()[keep[:n]])
    T = 100
    bxs = cu(synth.boxes(B, T, seed=6))
    pts = xyz.clone()
    pts[:, :4000] = bxs[:, torch.randint(0, T, (4000,), device="cuda"), :3] + torch.randn(B, 4000, 3, device="cuda")
    ref_o = torch.full((B, N), -1, dtype=torch.int32, device="cuda")
    roi.points_in_boxes_gpu(bxs, pts, ref_o)
    assert torch.equal(ru.points_in_boxes_gpu(pts, bxs), ref_o)


def test_degenerate_sizes_do_not_launch_or_crash(ops):
    """Empty batches / clouds / query sets / channel counts through the Python mirror: shapes come back right, nothing
    faults (the reference would launch zero-sized grids and exit(-1) on the launch error)."""
    pu, iu, ru = ops
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device="cuda")  # noqa: E731
    assert pu.furthest_point_sample(z(0, 10, 3), 4).shape == (0, 4)
    assert pu.furthest_point_sample(z(2, 10, 3), 0).shape == (2, 0)
    assert pu.ball_query(0.5, 4, z(2, 10, 3), z(2, 0, 3)).shape == (2, 0, 4)
    cnt, idx = pu.ball_query_cnt(0.5, 4, z(2, 0, 3), z(2, 5, 3))
    assert cnt.shape == (2, 5) and int(cnt.sum()) == 0 and int(idx.sum()) == 0
    assert pu.grouping_operation(z(2, 0, 7), z(2, 3, 4, dt=torch.int32)).shape == (2, 0, 3, 4)
    assert pu.gather_operation(z(2, 3, 7), z(2, 0, dt=torch.int32)).shape == (2, 3, 0)
    assert pu.group_concat(z(1, 8, 3), z(1, 2, 3), None, z(1, 2, 4, dt=torch.int32)).shape == (1, 3, 2, 4)
    d, i = pu.three_nn(z(1, 4, 3), z(1, 2, 3))
    assert bool(torch.isinf(d[..., 2]).all()) and int(i[..., 2].sum()) == 0     # two known points: third slot stays (inf, 0)
    assert iu.boxes_iou_bev(z(0, 7), z(5, 7)).shape == (0, 5)
    keep, _ = iu.nms_gpu(z(0, 7), z(0), 0.1)
    assert keep.numel() == 0
    assert ru.points_in_boxes_gpu(z(1, 0, 3), z(1, 4, 7)).shape == (1, 0)
    assert int((ru.points_in_boxes_gpu(z(1, 6, 3), z(1, 0, 7)) + 1).sum()) == 0
    assert pu.calc_dist_matrix_for_sampling(z(1, 1, 3), None, 1.0).shape == (1, 1, 1)
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ the chain
def test_chain_graph_equals_eager_and_shards_concatenate(lib):
    from de6d_b200 import chain
    cfg = chain.small_config()
    host = chain.make_inputs(cfg, batch=4, seed=1)
    full = chain.OpChain(cfg, 4, use_graph=True)
    out_g = {k: v.clone() for k, v in full.step_host(host).items()}
    eager = chain.OpChain(cfg, 4, use_graph=False)
    out_e = eager.step_host(host)
    for k in out_g:
        assert torch.equal(out_g[k], out_e[k]), k
    comp = chain.OpChain(cfg, 4, use_graph=False, fused_group=False, fused_ffps=False)   # the reference's op-by-op call sequence
    comp.step_host(host)
    torch.cuda.synchronize()
    for k, v in eager.outputs.items():
        if isinstance(v, torch.Tensor) and not k.endswith("_xyz_t"):     # the transposed clouds exist only on the fused route
            assert torch.equal(v, comp.outputs[k]), k
    # batch sharding: two half-batches reproduce the full batch bit for bit (ops never mix frames)
    halves = []
    for lo in (0, 2):
        sub = {k: v[lo:lo + 2].contiguous().pin_memory() for k, v in host.items()}
        c = chain.OpChain(cfg, 2, use_graph=True)
        halves.append({k: v.clone() for k, v in c.step_host(sub).items()})
    for k in out_g:
        assert torch.equal(torch.cat([halves[0][k], halves[1][k]]), out_g[k]), k


def test_chain_vs_oracle_chain(orc, lib):
    """Whole op chain (small config) against the same chain executed with the oracle on the CPU."""
    from de6d_b200 import chain
    from oracle import chain_ref
    cfg = chain.small_config()
    host = chain.make_inputs(cfg, batch=2, seed=3)
    c = chain.OpChain(cfg, 2, use_graph=False, keep_matrices=True)
    c.step_host(host)
    torch.cuda.synchronize()
    want = chain_ref.run_chain(cfg, host, keep_groups=True)   # F-FPS matrix from the oracle: bit-identical to the kernel's
    for k, v in want.items():
        if v is None:
            continue
        got = c.outputs[k].cpu().numpy()
        if k == "nms_keep":
            for f in range(2):
                n = want["nms_num"][f]
                np.testing.assert_array_equal(got[f, :n], v[f, :n], err_msg=k)
        else:
            np.testing.assert_array_equal(got, v, err_msg=k)

"""Input staging: host mirror of DataProcessor.sample_points against selections produced by the
unmodified reference (tests/golden/golden_staging.npz, made by make_golden_staging.py), and the device staging kernel
against the numpy restatement of break_up_pc / points[choice].  Copies: bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden_staging as mgs  # noqa: E402  (seeded frame generator + case table; reads no reference file on import)


def test_sample_points_choice_matches_reference_selection(golden_dir):
    from de6d_b200.staging import sample_points_choice
    gold = np.load(os.path.join(golden_dir, "golden_staging.npz"))
    for name, n, num, far, seed in mgs.CASES:
        pts = mgs.frame(n, far, seed)
        np.random.seed(seed)
        choice = sample_points_choice(pts, num)
        assert np.array_equal(choice, gold[name]), name
        assert len(choice) == num and choice.min() >= 0 and choice.max() < n
        if num <= n:
            assert len(np.unique(choice)) == num                      # drawn without replacement
        if name == "near_fill":                                       # every far point survives (data_processor.py:160-163)
            far_rows = np.where(np.linalg.norm(pts[:, :3], axis=1) >= 40.0)[0]
            assert np.isin(far_rows, choice).all()
        if name in ("pad_once", "pad_replace"):
            assert np.isin(np.arange(n), choice).all()                # short frames keep every point
    assert np.array_equal(sample_points_choice(mgs.frame(10, 0.0, 9), -1), np.arange(10))


def test_collate_choice_offsets():
    from de6d_b200.staging import collate_choice
    frames = [mgs.frame(n, 0.2, s)[:, :4] for s, n in enumerate((700, 300, 512))]
    rows, choice = collate_choice(frames, 512, np.random.RandomState(0))
    assert rows.shape == (1512, 4) and choice.shape == (3, 512) and choice.dtype == np.int32
    lo = np.array([0, 700, 1000])[:, None]
    hi = np.array([700, 1000, 1512])[:, None]
    assert ((choice >= lo) & (choice < hi)).all()


def test_oracle_break_up_pc_small(orc):
    B, N, C = 3, 5, 2
    rng = np.random.default_rng(0)
    pc = rng.random((B * N, 4 + C)).astype(np.float32)
    pc[:, 0] = np.repeat(np.arange(B), N)
    bidx, xyz, feats = orc.break_up_pc(pc, B)
    assert xyz.shape == (B, N, 3) and feats.shape == (B, C, N)
    assert xyz[1, 2, 1] == pc[N + 2, 2] and feats[2, 1, 3] == pc[2 * N + 3, 5] and bidx[2, 0] == 2.0
    pc[0, 0] = 1
    with pytest.raises(AssertionError):
        orc.break_up_pc(pc, B)


# ---------------------------------------------------------------------------------------------------------------
def _collated(B, N, C, seed):
    rng = np.random.default_rng(seed)
    pc = rng.normal(0, 10, (B * N, 4 + C)).astype(np.float32)
    pc[:, 0] = np.repeat(np.arange(B), N)
    return pc


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,C", [(1, 1, 0), (2, 255, 1), (3, 257, 1), (2, 1000, 4), (4, 16384, 1), (2, 4096, 64), (1, 300, 0)])
def test_break_up_pc_vs_oracle(orc, lib, B, N, C):
    from de6d_b200.staging import break_up_pc
    pc = _collated(B, N, C, B * 1000 + N + C)
    e_b, e_x, e_f = orc.break_up_pc(pc, B)
    bidx, xyz, feats = break_up_pc(torch.from_numpy(pc).cuda(), B)
    assert np.array_equal(bidx.cpu().numpy(), e_b) and np.array_equal(xyz.cpu().numpy(), e_x)
    if C == 0:
        assert feats is None and e_f is None
    else:
        assert feats.is_contiguous() and np.array_equal(feats.cpu().numpy(), e_f)


@pytest.mark.gpu
def test_break_up_pc_rejects_unequal_frames(orc, lib):
    from de6d_b200._lib import De6dError
    from de6d_b200.staging import break_up_pc
    pc = _collated(4, 500, 1, 0)
    pc[500 * 2 - 1, 0] = 2                      # frame 1 one row short, frame 2 one row long: counts differ
    with pytest.raises(AssertionError):
        orc.break_up_pc(pc, 4)
    with pytest.raises(AssertionError):
        break_up_pc(torch.from_numpy(pc).cuda(), 4)
    break_up_pc(torch.from_numpy(pc).cuda(), 4, check=False)          # sync-free form does not look
    with pytest.raises(AssertionError):
        break_up_pc(torch.from_numpy(pc[:-1]).cuda(), 4)
    with pytest.raises(De6dError):
        from de6d_b200._lib import call
        x = torch.zeros(12, device="cuda")
        call("de6d_stage_points", 2, 2, 0, 0, 5, x.data_ptr(), None, x.data_ptr(), None, None, None, 0)


@pytest.mark.gpu
def test_stage_frames_equals_reference_pipeline(orc, lib):
    """host points[choice] + collate + break_up_pc (numpy)  ==  one staging kernel on the raw rows."""
    from de6d_b200.staging import collate_choice, stage_frames
    N = 2048
    sizes = (3000, 1500, 600, 2048, 5000)
    frames = [mgs.frame(n, 0.25, 40 + s)[:, :4] for s, n in enumerate(sizes)]
    rows, choice = collate_choice(frames, N, np.random.RandomState(7))
    rs = np.random.RandomState(7)                                       # the reference pipeline, frame by frame
    from de6d_b200.staging import sample_points_choice
    coll = np.concatenate([np.pad(f[sample_points_choice(f, N, rs)], ((0, 0), (1, 0)), constant_values=k)
                           for k, f in enumerate(frames)]).astype(np.float32)
    _, e_x, e_f = orc.break_up_pc(coll, len(frames))
    xyz, feats = stage_frames(torch.from_numpy(rows).cuda(), torch.from_numpy(choice).cuda())
    assert np.array_equal(xyz.cpu().numpy(), e_x) and np.array_equal(feats.cpu().numpy(), e_f)
    bad = choice.copy()
    bad[1, 5] = rows.shape[0]
    bad[2, 9] = -1
    with pytest.raises(AssertionError, match="2 sample indices"):
        stage_frames(torch.from_numpy(rows).cuda(), torch.from_numpy(bad).cuda())
    xyz2, _ = stage_frames(torch.from_numpy(rows).cuda(), torch.from_numpy(bad).cuda(), check=False)
    assert float(xyz2[1, 5].abs().sum()) == 0.0 and float(xyz2[2, 9].abs().sum()) == 0.0


@pytest.mark.gpu
def test_staged_frames_feed_the_sampling_path(orc, lib):
    """The staged tensors are what the first SA layer reads: D-FPS on them equals the oracle on the host-built ones."""
    from de6d_b200 import pointnet2_utils as pu
    from de6d_b200.staging import break_up_pc
    pc = _collated(2, 4096, 1, 5)
    pc[:, 1:4] = np.random.default_rng(5).uniform(-40, 40, (2 * 4096, 3)).astype(np.float32)
    _, xyz, _ = break_up_pc(torch.from_numpy(pc).cuda(), 2)
    got = pu.furthest_point_sample(xyz, 512).cpu().numpy()
    _, e_x, _ = orc.break_up_pc(pc, 2)
    assert np.array_equal(got, orc.furthest_point_sample(e_x, 512))

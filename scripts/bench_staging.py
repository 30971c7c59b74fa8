"""Times the input-staging kernel against the torch composition the reference runs
(pointnet2_backbone.py:193-222: slice copies, B `.sum()` syncs, view, permute().contiguous()) on the same collated
array.  Development aid; prints a small table (copied to profiles/)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from de6d_b200.staging import break_up_pc, stage_frames  # noqa: E402


def timeit(fn, reps=20, warm=5):
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


def reference_composition(pc, B):
    bidx = pc[:, 0]
    xyz = pc[:, 1:4].contiguous()
    feats = pc[:, 4:].contiguous() if pc.size(-1) > 4 else None
    cnt = xyz.new_zeros(B).int()
    for b in range(B):
        cnt[b] = (bidx == b).sum()
    assert cnt.min() == cnt.max()
    xyz = xyz.view(B, -1, 3).contiguous()
    if feats is not None:
        feats = feats.view(B, -1, feats.shape[-1]).permute(0, 2, 1).contiguous()
    return bidx.view(B, -1).float(), xyz, feats


peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6551.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6551.0
rows = []
for B, N, C in ((64, 16384, 1), (256, 16384, 1), (64, 16384, 4), (16, 131072, 1)):
    rng = np.random.default_rng(0)
    pc = rng.normal(0, 10, (B * N, 4 + C)).astype(np.float32)
    pc[:, 0] = np.repeat(np.arange(B), N)
    d = torch.from_numpy(pc).cuda()
    r = reference_composition(d, B)
    g = break_up_pc(d, B)
    assert all(torch.equal(a, b) for a, b in zip(r, g))
    t_ref = timeit(lambda: reference_composition(d, B), reps=5, warm=2)
    t_chk = timeit(lambda: break_up_pc(d, B))
    t_free = timeit(lambda: break_up_pc(d, B, check=False))
    raw = torch.from_numpy(np.ascontiguousarray(pc[:, 1:])).cuda()
    choice = torch.from_numpy(np.stack([rng.permutation(N) + b * N for b in range(B)]).astype(np.int32)).cuda()
    t_gather = timeit(lambda: stage_frames(raw, choice, check=False))
    alg = B * N * 4 * ((4 + C) + (3 + C) + 1)          # rows read + xyz/features/batch_idx written
    rows.append((B, N, C, t_ref, t_chk, t_free, t_gather, alg / t_free / 1e6, alg / t_free / 1e6 / peak))
print("| B | N | C | torch composition (reference) ms | stage kernel + 1 sync ms | stage kernel sync-free ms | "
      "with sample_points gather ms | GB/s (alg. bytes, sync-free) | of HBM peak |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print("| %d | %d | %d | %.3f | %.3f | %.3f | %.3f | %.0f | %.2f |" % r)

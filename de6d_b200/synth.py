"""Seeded synthetic KITTI-shape inputs.  numpy PCG64 so the
same seed gives the same bytes in this container and on the GPU box.

The reference ships no point clouds and no configs; these generators are the
workload definition shared by tests/, bench.py and tests/golden/make_golden.py.
"""
import numpy as np

# KITTI crop used by the reference (range literal at pointnet2_modules.py:391)
X_RANGE = (0.0, 70.4)
Y_RANGE = (-40.0, 40.0)
Z_RANGE = (-3.0, 1.0)


def clouds(batch, n, seed=0, dup_frac=0.0):
    """(batch, n, 3) float32 uniform clouds.  dup_frac > 0 overwrites that fraction of
    points with copies of other points -- mimics DataProcessor.sample_points padding
    (datasets/processor/data_processor.py:170-176) and exercises the FPS tie rule."""
    rng = np.random.default_rng(seed)
    p = np.empty((batch, n, 3), np.float32)
    p[..., 0] = rng.uniform(*X_RANGE, size=(batch, n))
    p[..., 1] = rng.uniform(*Y_RANGE, size=(batch, n))
    p[..., 2] = rng.uniform(*Z_RANGE, size=(batch, n))
    if dup_frac > 0:
        k = int(n * dup_frac)
        for b in range(batch):
            dst = rng.choice(n, size=k, replace=False)
            src = rng.integers(0, n, size=k)
            p[b, dst] = p[b, src]
    return p


def lidar_clouds(batch, n, seed=0, beams=64):
    """Ring-structured variant: `beams` elevation rings x n/beams azimuths hitting a ground plane
    at z=-1.7 (clipped to 80 m), plus range noise.  Realistic density fall-off for ball query."""
    rng = np.random.default_rng(seed)
    per = n // beams
    elev = np.deg2rad(np.linspace(-24.8, 2.0, beams, dtype=np.float64))
    out = np.empty((batch, n, 3), np.float32)
    for b in range(batch):
        az = rng.uniform(-np.pi / 4, np.pi / 4, size=(beams, per))
        rng_ground = np.where(elev[:, None] < -0.01, 1.7 / np.maximum(np.tan(-elev[:, None]), 1e-3), 80.0)
        r = np.minimum(rng_ground, 80.0) * rng.uniform(0.6, 1.0, size=(beams, per))
        x = r * np.cos(elev[:, None]) * np.cos(az)
        y = r * np.cos(elev[:, None]) * np.sin(az)
        z = r * np.sin(elev[:, None])
        pts = np.stack([x, y, z], -1).reshape(-1, 3)
        if pts.shape[0] < n:  # pad by duplication, like the reference data pipeline
            extra = pts[rng.integers(0, pts.shape[0], size=n - pts.shape[0])]
            pts = np.concatenate([pts, extra], 0)
        out[b] = pts[rng.permutation(n)].astype(np.float32)
    return out


def boxes(batch, t, seed=0):
    """(batch, t, 7) car-sized boxes [x, y, z, dx, dy, dz, heading] (prior: experiments/results/Det6D.npy)."""
    rng = np.random.default_rng(seed + 1000)
    bx = np.empty((batch, t, 7), np.float32)
    bx[..., 0] = rng.uniform(*X_RANGE, size=(batch, t))
    bx[..., 1] = rng.uniform(*Y_RANGE, size=(batch, t))
    bx[..., 2] = rng.uniform(-1.5, -0.5, size=(batch, t))
    dims = rng.normal((3.9, 1.6, 1.56), 0.2, size=(batch, t, 3))
    bx[..., 3:6] = np.maximum(dims, 0.1)
    bx[..., 6] = rng.uniform(-np.pi, np.pi, size=(batch, t))
    return bx


def proposals(batch, n=512, seed=0, clusters=64, sigma_xy=0.3, sigma_yaw=0.1):
    """NMS input: `clusters` objects x n/clusters jittered copies so NMS actually suppresses.
    Returns boxes (batch, n, 7) and scores (batch, n) with distinct values per frame."""
    rng = np.random.default_rng(seed + 2000)
    per = n // clusters
    base = boxes(batch, clusters, seed=seed + 7)
    bx = np.repeat(base, per, axis=1)
    if bx.shape[1] < n:
        bx = np.concatenate([bx, boxes(batch, n - bx.shape[1], seed=seed + 11)], axis=1)
    bx = bx.copy()
    bx[..., 0:2] += rng.normal(0, sigma_xy, size=(batch, n, 2)).astype(np.float32)
    bx[..., 6] += rng.normal(0, sigma_yaw, size=(batch, n)).astype(np.float32)
    bx[..., 3:6] *= rng.uniform(0.95, 1.05, size=(batch, n, 3)).astype(np.float32)
    scores = np.stack([rng.permutation(n) for _ in range(batch)]).astype(np.float32)
    scores = (scores + 0.5) / n  # distinct, so every sort order is unambiguous
    return bx.astype(np.float32), scores.astype(np.float32)


def weights(batch, n, seed=0):
    """S-FPS weights: sigmoid(N(0,1)) (caller side is sigmoid(score)**gamma, pointnet2_modules.py:419)."""
    rng = np.random.default_rng(seed + 3000)
    return (1.0 / (1.0 + np.exp(-rng.normal(size=(batch, n))))).astype(np.float32)


def features(batch, c, n, seed=0):
    rng = np.random.default_rng(seed + 4000)
    return rng.normal(size=(batch, c, n)).astype(np.float32)


def dist_matrix(xyz, feats=None, gamma=1.0):
    """F-FPS input, calc_dist_matrix_for_sampling (pointnet2_utils.py:36-44) in float64 -> float32.
    The matrix is an INPUT of the kernel under test, so how it is produced does not affect parity."""
    x = xyz.astype(np.float64)
    d = np.sqrt(np.maximum(((x[:, :, None, :] - x[:, None, :, :]) ** 2).sum(-1), 0.0))
    if feats is not None:
        f = np.transpose(feats.astype(np.float64), (0, 2, 1))
        d = d + gamma * np.sqrt(np.maximum(((f[:, :, None, :] - f[:, None, :, :]) ** 2).sum(-1), 0.0))
    return d.astype(np.float32)


class AttrDict(dict):
    """Minimal EasyDict stand-in (attribute access + dict methods) for the reference's `model_cfg` objects."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    @staticmethod
    def wrap(x):
        if isinstance(x, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in x.items()})
        return x


def sasa_backbone_cfg(n_points=16384, scale=1):
    """MODEL.BACKBONE_3D of the SASA / 3DSSD-style PointNet2FSMSG backbone Det6D uses (pointnet2_backbone.py:97-191 reads these
    keys).  The reference ships no YAML: layer shapes are the ones BASELINE.json configs[1] names
    (16384 -> 4096 -> 1024 -> 512), radii / nsample / MLP widths as upstream SASA's 3dssd_sasa_car.yaml.  `scale` divides
    the point counts for small test clouds."""
    n1, n2, n3 = 4096 // scale, 512 // scale, 256 // scale
    return AttrDict.wrap({
        "NAME": "PointNet2FSMSG",
        "SA_CONFIG": {
            "NPOINT_LIST": [[n1], [n2, n2], [n3, n3]],
            "SAMPLE_RANGE_LIST": [[[0, n_points]], [[0, n1], [0, n1]], [[0, n2], [n2, 2 * n2]]],
            "SAMPLE_METHOD_LIST": [["d-fps"], ["f-fps", "d-fps"], ["s-fps", "d-fps"]],
            "RADIUS": [[0.2, 0.4, 0.8], [0.4, 0.8, 1.6], [1.6, 3.2, 4.8]],
            "NSAMPLE": [[32, 32, 64], [32, 32, 64], [32, 32, 32]],
            "MLPS": [[[16, 16, 32], [16, 16, 32], [32, 32, 64]],
                     [[64, 64, 128], [64, 64, 128], [64, 96, 128]],
                     [[128, 128, 256], [128, 192, 256], [128, 256, 256]]],
            "AGGREGATION_MLPS": [[64], [128], [256]],
            "CONFIDENCE_MLPS": [[32], [64], []],
            "WEIGHT_GAMMA": 1.0,
        },
    })

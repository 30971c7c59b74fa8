#!/usr/bin/env python
"""Generates tests/golden/golden_staging.npz by running the UNMODIFIED reference
`DataProcessor.sample_points` (pcdet/datasets/processor/data_processor.py:145-177) on seeded frames.

    python tests/golden/make_golden_staging.py        # needs /root/reference (this container only)

The reference module is loaded from where it lies with its unrelated imports (skimage, pcdet.utils.*) stubbed; the
frames carry their own row number in the last column so the selection (`choice`) can be read back from the output.
The fixtures store only seeds, sizes and the selected row numbers: tests rebuild the frames with the same generator.
"""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("DE6D_REFERENCE", "/root/reference/core")

# (name, points in the frame, NUM_POINTS, fraction beyond 40 m, seed): one case per branch of the reference
CASES = [
    ("near_fill", 3000, 2048, 0.2, 0),       # far points kept, near points drawn
    ("no_far", 3000, 2048, 0.0, 1),          # no far points at all
    ("far_overflow", 3000, 1024, 0.6, 2),    # more far points than the budget: uniform draw
    ("pad_once", 1500, 2048, 0.2, 3),        # short frame, pad without replacement
    ("pad_replace", 600, 2048, 0.2, 4),      # very short frame, pad with replacement
    ("exact", 2048, 2048, 0.2, 5),           # equal: shuffle only
]


def frame(n, far_frac, seed):
    """Seeded (n, 5) frame [x, y, z, intensity, row]; `far_frac` of the points lie beyond 40 m."""
    rng = np.random.default_rng(1000 + seed)
    r = np.where(rng.random(n) < far_frac, rng.uniform(41.0, 70.0, n), rng.uniform(2.0, 39.0, n))
    a = rng.uniform(-0.7, 0.7, n)
    pts = np.stack([r * np.cos(a), r * np.sin(a), rng.uniform(-3.0, 1.0, n), rng.random(n),
                    np.arange(n, dtype=np.float64)], axis=1)
    return pts.astype(np.float32)


def load_reference():
    for name in ("pcdet", "pcdet.utils", "pcdet.utils.box_utils", "pcdet.utils.common_utils", "pcdet.datasets",
                 "pcdet.datasets.processor", "skimage", "skimage.transform"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    sys.modules["pcdet.utils"].box_utils = sys.modules["pcdet.utils.box_utils"]
    sys.modules["pcdet.utils"].common_utils = sys.modules["pcdet.utils.common_utils"]
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    path = os.path.join(REF, "pcdet", "datasets", "processor", "data_processor.py")
    spec = importlib.util.spec_from_file_location("pcdet.datasets.processor.data_processor", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.DataProcessor


def main():
    DP = load_reference()
    out = {}
    for name, n, num, far, seed in CASES:
        pts = frame(n, far, seed)
        np.random.seed(seed)
        me = SimpleNamespace(mode="train")
        res = DP.sample_points(me, {"points": pts}, SimpleNamespace(NUM_POINTS={"train": num}))
        out[name] = res["points"][:, 4].astype(np.int32)
        assert len(out[name]) == num
    dst = os.path.join(ROOT, "tests", "golden", "golden_staging.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()

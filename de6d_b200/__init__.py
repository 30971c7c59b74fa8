"""de6d_b200 -- B200-native (sm_100a) point-set-abstraction and box ops for Det6D / SASA / 3DSSD detectors.

Importing the op modules loads de6d_b200/lib/libde6d_b200.so and raises if it is missing: there is no CPU
or PyTorch fallback.  Layout:
    csrc/                     hand-written CUDA kernels + the C ABI (include/de6d_b200.h)
    compat/                   drop-ins for the reference's pybind11 modules (same names / positional args)
    pointnet2_utils.py        mirrors of the reference's Python op wrappers
    iou3d_nms_utils.py
    roiaware_pool3d_utils.py
    chain.py                  the SA + NMS op chain of BASELINE.json (streams + CUDA graph), batch sharding
    synth.py                  seeded synthetic KITTI-shape inputs
"""
__version__ = "0.1.0"

"""de6d_b200 -- B200-native (sm_100a) point-set-abstraction and box ops for Det6D / SASA / 3DSSD detectors.

Importing the op modules loads de6d_b200/lib/libde6d_b200.so and raises if it is missing: there is no CPU
or PyTorch fallback.  Layout:
    csrc/                     hand-written CUDA kernels + the C ABI (include/de6d_b200.h)
    compat/                   drop-ins for the reference's pybind11 modules (same names / positional args)
    pointnet2_utils.py        mirrors of the reference's Python op wrappers (+ fused group_concat,
    iou3d_nms_utils.py          calc_dist_matrix_for_sampling / furthest_point_sample_features, BatchedNMS)
    roiaware_pool3d_utils.py
    model_nms_utils.py        class_agnostic_nms and its batched, host-sync-free form (post-processing)
    box_utils.py              full-pose (9-DoF) points_in_boxes3d on the device
    staging.py                sample_points selection + break_up_pc: raw frames -> (B,N,3) / (B,C,N) in one kernel
    chain.py                  the SA + NMS op chain of BASELINE.json (streams + CUDA graph)
    dist.py                   frame sharding across ranks, NCCL gather of detections
    synth.py                  seeded synthetic KITTI-shape inputs
    build.py                  nvcc build of the library (python -m de6d_b200.build)
"""
__version__ = "0.1.0"

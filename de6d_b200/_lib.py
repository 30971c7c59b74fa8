"""ctypes binding of libde6d_b200.so (the C ABI declared in include/de6d_b200.h).

There is no fallback of any kind: if the shared library is missing or fails to load, importing the ops
raises.  Build it with `python -m de6d_b200.build` (or __graft_entry__.build()).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DE6D_LIB") or os.path.join(_HERE, "lib", "libde6d_b200.so")   # DE6D_LIB: development builds

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_sz = C.c_size_t

# name -> argtypes (all return int status unless listed in _RESTYPES)
PROTOTYPES = {
    "de6d_furthest_point_sampling": [_i, _i, _i, _p, _p, _p, _p],
    "de6d_furthest_point_sampling_matrix": [_i, _i, _i, _p, _p, _p, _p],
    "de6d_furthest_point_sampling_weights": [_i, _i, _i, _p, _p, _p, _p, _p],
    "de6d_furthest_point_sampling_impl": [_i, _i, _i, _p, _p, _p, _i, _p],
    "de6d_furthest_point_sampling_weights_impl": [_i, _i, _i, _p, _p, _p, _p, _i, _p],
    "de6d_dist_matrix": [_i, _i, _i, _p, _p, C.c_longlong, C.c_longlong, C.c_longlong, _f, _p, _p],
    "de6d_furthest_point_sampling_features_fits": [_i, _i],
    "de6d_furthest_point_sampling_features": [_i, _i, _i, _i, _p, _p, C.c_longlong, C.c_longlong, C.c_longlong, _f, _p, _p, _p],
    "de6d_furthest_point_sampling_features_impl": [_i, _i, _i, _i, _p, _p, C.c_longlong, C.c_longlong, C.c_longlong, _f, _p, _p, _i, _i, _p],
    "de6d_gather_points": [_i, _i, _i, _i, _p, _p, _p, _p],
    "de6d_gather_points_grad": [_i, _i, _i, _i, _p, _p, _p, _p],
    "de6d_ball_query": [_i, _i, _i, _f, _i, _p, _p, _p, _p],
    "de6d_ball_query_cnt": [_i, _i, _i, _f, _i, _p, _p, _p, _p, _p],
    "de6d_ball_query_dilated": [_i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _p],
    "de6d_ball_query_workspace_bytes": [_i, _i],
    "de6d_ball_query_grid_bytes": [_i, _i],
    "de6d_ball_query_grid_build": [_i, _i, _f, _p, _p, _sz, _p],
    "de6d_ball_query_ex": [_i, _i, _i, _i, _i, _f, _f, _i, _p, _p, _p, _p, _p, _sz, _p],
    "de6d_group_points": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "de6d_group_points_impl": [_i, _i, _i, _i, _i, _p, _p, _p, _i, _p],
    "de6d_group_concat": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "de6d_group_concat_t": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "de6d_gather_xyz": [_i, _i, _i, _p, _p, _p, _p, _p],
    "de6d_group_points_grad": [_i, _i, _i, _i, _i, _p, _p, _p, _p],
    "de6d_three_nn": [_i, _i, _i, _p, _p, _p, _p, _p],
    "de6d_three_nn_ex": [_i, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p],
    "de6d_three_interpolate": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "de6d_three_interpolate_grad": [_i, _i, _i, _i, _p, _p, _p, _p, _p],
    "de6d_boxes_overlap_bev": [_i, _p, _i, _p, _p, _p],
    "de6d_boxes_iou_bev": [_i, _p, _i, _p, _p, _p],
    "de6d_boxes_iou3d": [_i, _p, _i, _p, _p, _p],
    "de6d_boxes_iou3d9": [_i, _p, _i, _p, _p, _p],
    "de6d_nms_workspace_bytes": [_i, _i],
    "de6d_nms_workspace_init": [_i, _p, _p],
    "de6d_nms_batched": [_i, _i, _p, _p, _f, _i, _p, _p, _p, _sz, _p],
    "de6d_points_in_boxes": [_i, _i, _i, _p, _p, _p, _p],
    "de6d_points_in_boxes9": [_i, _i, _i, _p, _p, _p, _p],
    "de6d_points_in_boxes_mask": [_i, _i, _p, _p, _p, _p],
    "de6d_points_in_boxes_mask_host": [_i, _i, _p, _p, _p, _i],
    "de6d_boxes_iou_bev_host": [_i, _p, _i, _p, _p, _i],
    "de6d_sa_mlp_fits": [_i, _p, _i],
    "de6d_sa_mlp_packed_floats": [_i, _p],
    "de6d_sa_mlp_pack": [_i, _p, _p, _p, _p],
    "de6d_sa_mlp_fused": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p],
    "de6d_sa_mlp_fused_slice": [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _p, _p],
    "de6d_stage_points": [_i, _i, _i, _i, C.c_longlong, _p, _p, _p, _p, _p, _p, _p],
    "de6d_last_error_string": [],
    "de6d_version": [],
    "de6d_build_info": [],
    "de6d_launch_count": [],
}
_RESTYPES = {
    "de6d_nms_workspace_bytes": _sz,
    "de6d_ball_query_workspace_bytes": _sz,
    "de6d_ball_query_grid_bytes": _sz,
    "de6d_sa_mlp_packed_floats": _sz,
    "de6d_last_error_string": C.c_char_p,
    "de6d_build_info": C.c_char_p,
    "de6d_launch_count": C.c_longlong,
}
_NO_STATUS = set(_RESTYPES) | {"de6d_version", "de6d_furthest_point_sampling_features_fits", "de6d_sa_mlp_fits"}

_lib = None


class De6dError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "de6d_b200: %s not found -- the CUDA library is required (no CPU or PyTorch fallback exists). "
            "Build it with `python -m de6d_b200.build`." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch, which must be loud
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib


_trace = None   # None, or a list of (entry point, args, start event, end event) filled by call()


def trace_begin():
    """Start bracketing every entry-point call with CUDA events on the stream it is launched on (bench.py's
    per-kernel timing pass; not for use inside CUDA-graph capture)."""
    global _trace
    _trace = []


def trace_end():
    """Stop tracing and return [(name, args, milliseconds)]; synchronises the device."""
    global _trace
    import torch
    torch.cuda.synchronize()
    out = [(name, args, e0.elapsed_time(e1)) for name, args, e0, e1 in (_trace or [])]
    _trace = None
    return out


_fn_cache = {}


def call(name, *args):
    """Invoke a status-returning entry point; raise De6dError with the library's message on failure."""
    lib = _lib or load()
    if _trace is None:
        fn = _fn_cache.get(name)
        if fn is None:
            fn = _fn_cache[name] = getattr(lib, name)
        rc = fn(*args)
        if rc != 0:
            raise De6dError("%s failed (code %d): %s" % (name, rc, lib.de6d_last_error_string().decode()))
        return
    if _trace is not None:
        import torch
        s = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        rc = getattr(lib, name)(*args)
        e1.record(s)
        _trace.append((name, args, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise De6dError("%s failed (code %d): %s" % (name, rc, lib.de6d_last_error_string().decode()))


def launch_count():
    return int(load().de6d_launch_count())

"""The SA + NMS op chain executed with the REFERENCE's own CUDA kernels through the reference's own Python
(oracle/_ref/*.so under oracle/_ref/py/pcdet/...), op by op, exactly as the reference model issues them.

TEST INFRASTRUCTURE / yardstick only: used by tests/ (chain-level parity on the GPU) and by bench.py's `reference_cuda`
extra (rank 0, outside every timed region of the product arm).  Never imported by de6d_b200.

Call sequence per SA layer = _PointnetSAModuleFSBase.forward (pointnet2_modules.py:374-463): transpose, per-method FPS
(D-FPS; F-FPS = calc_dist_matrix_for_sampling (2 x torch.cdist) + furthest_point_sample_matrix; S-FPS), cat, gather_operation,
transpose, then one QueryWithCntAndGroup module per radius scale; head = the same grouper over the vote centres;
post-processing = the per-frame Python loop of Detector3DTemplate.post_processing (detector3d_template.py:199-261)
calling iou3d_nms_utils.nms_gpu (cudaMalloc + kernel + blocking D2H + host sweep per frame).
Everything runs on the legacy default stream like the reference.
"""
import time

import torch


def run(cfg, inp, tree, timed=False, keep_groups=False):
    """inp: dict of CUDA tensors as de6d_b200.chain.OpChain.inputs.  tree: oracle.ref_py.load_tree(..., extensions=
    build_ref.load()).  Returns (outs, per-op milliseconds {name: ms} when timed)."""
    pu, iu = tree.pointnet2_utils, tree.iou3d_nms_utils
    outs, ops = {}, {}

    def op(name, fn):
        if not timed:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        e1.synchronize()
        ops[name] = ops.get(name, 0.0) + e0.elapsed_time(e1)
        return r

    xyz = inp["xyz"]
    with torch.no_grad():
        for li, layer in enumerate(cfg.layers):
            feats, scores = inp["feats%d" % li], inp["scores%d" % li]
            xyz_flipped = op("transpose", lambda: xyz.transpose(1, 2).contiguous())
            parts = []
            for method, npnt, (lo, hi) in zip(layer.methods, layer.npoints, layer.ranges):
                xs = xyz[:, lo:hi, :].contiguous()
                if method == "d-fps":
                    sidx = op("d-fps %d->%d" % (hi - lo, npnt), lambda: pu.furthest_point_sample(xs, npnt))
                elif method == "f-fps":
                    fs = feats[:, :, lo:hi]
                    mat = op("f-fps cdist %d" % (hi - lo), lambda: pu.calc_dist_matrix_for_sampling(xs, fs.permute(0, 2, 1), cfg.ffps_gamma))
                    sidx = op("f-fps matrix kernel %d->%d" % (hi - lo, npnt), lambda: pu.furthest_point_sample_matrix(mat, npnt))
                    del mat
                elif method == "s-fps":
                    w = scores[:, lo:hi].contiguous()
                    sidx = op("s-fps %d->%d" % (hi - lo, npnt), lambda: pu.furthest_point_sample_weights(xs, w, npnt))
                else:
                    raise NotImplementedError(method)
                parts.append(sidx + lo)
            sample_idx = torch.cat(parts, dim=-1)
            new_xyz = op("gather", lambda: pu.gather_operation(xyz_flipped, sample_idx).transpose(1, 2).contiguous())
            outs["l%d_idx" % li] = sample_idx
            outs["l%d_new_xyz" % li] = new_xyz
            for si, (r, ns) in enumerate(zip(layer.radii, layer.nsamples)):
                grouper = pu.QueryWithCntAndGroup(r, ns, use_xyz=True)
                cnt, nf = op("query+group l%d" % li, lambda: grouper(xyz, new_xyz, feats))
                outs["l%d_s%d_cnt" % (li, si)] = cnt
                if keep_groups:
                    outs["l%d_s%d" % (li, si)] = nf
                del nf
            xyz = new_xyz
        if cfg.n_votes > 0:
            votes = (xyz[:, :cfg.n_votes, :] + inp["vote_offsets"]).contiguous()
            for si, (r, ns) in enumerate(zip(cfg.vote_radii, cfg.vote_nsamples)):
                grouper = pu.QueryWithCntAndGroup(r, ns, use_xyz=True)
                cnt, nf = op("query+group head", lambda: grouper(xyz, votes, inp["vote_feats"]))
                outs["head_s%d_cnt" % si] = cnt
                if keep_groups:
                    outs["head_s%d" % si] = nf
                del nf
        if cfg.n_proposals > 0:
            def nms_loop():
                kept = []
                for f in range(inp["boxes"].shape[0]):          # detector3d_template.py:199: one call per frame
                    sel, _ = iu.nms_gpu(inp["boxes"][f], inp["box_scores"][f], cfg.nms_thresh)
                    kept.append(sel)
                return kept
            outs["nms_keep_list"] = op("nms per-frame loop", nms_loop)
    return outs, ops


def time_chain(cfg, inp, tree, steps=2, warmup=1):
    """Wall-clock (host-synchronised, like the reference's own timing habit) milliseconds per chain pass + per-op ms."""
    for _ in range(warmup):
        run(cfg, inp, tree)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        run(cfg, inp, tree)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    _, ops = run(cfg, inp, tree, timed=True)
    return ms, ops

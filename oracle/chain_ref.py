"""CPU execution of the SA + NMS op chain with the oracle (TEST INFRASTRUCTURE: used by tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs only).

Same call sequence as de6d_b200.chain.OpChain._forward, same synthetic inputs, numpy in / numpy out.  The
F-FPS distance matrix is produced with torch.cdist on the host exactly as the reference's
calc_dist_matrix_for_sampling would (pointnet2_utils.py:36-44) -- it is an input of the kernel under test.
"""
import numpy as np
import torch

from . import oracle as orc


def _group(xyz, new_xyz, feats, r, ns):
    cnt, idx = orc.ball_query_cnt(r, ns, xyz, new_xyz)
    g_xyz = orc.grouping_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    g_xyz = g_xyz - new_xyz.transpose(0, 2, 1)[..., None]
    g_feat = orc.grouping_operation(feats, idx)
    return cnt, np.concatenate([g_xyz, g_feat], axis=1)


def run_chain(cfg, host, frames=None, keep_groups=False, matrices=None, ffps_matrix="oracle"):
    """host: dict of numpy arrays / CPU tensors from de6d_b200.chain.make_inputs.  frames: optional slice.
    matrices: optional {layer index: (B, n, n) array} used instead of recomputing the F-FPS distance matrix
    (the matrix is an input of the F-FPS op under test).  ffps_matrix: "oracle" = orc.calc_dist_matrix_for_sampling
    (direct differences, bit-identical to de6d_b200's dist-matrix kernel), "torch" = torch.cdist exactly as the
    reference's calc_dist_matrix_for_sampling (used by the CPU timing legs of bench.py)."""
    h = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    if frames is not None:
        h = {k: np.ascontiguousarray(v[frames]) for k, v in h.items()}
    out = {}
    xyz = h["xyz"]
    for li, layer in enumerate(cfg.layers):
        feats, scores = h["feats%d" % li], h["scores%d" % li]
        parts = []
        for method, npnt, (lo, hi) in zip(layer.methods, layer.npoints, layer.ranges):
            sl = np.ascontiguousarray(xyz[:, lo:hi])
            if method == "d-fps":
                idx = orc.furthest_point_sample(sl, npnt)
            elif method == "f-fps":
                x = torch.from_numpy(sl)
                f = torch.from_numpy(np.ascontiguousarray(feats[:, :, lo:hi])).permute(0, 2, 1)
                if matrices is not None and li in matrices:
                    mat = torch.from_numpy(np.ascontiguousarray(matrices[li]))
                elif ffps_matrix == "torch":
                    mat = torch.cdist(x, x) + torch.cdist(f, f) * cfg.ffps_gamma
                else:
                    mat = torch.from_numpy(orc.calc_dist_matrix_for_sampling(sl, np.ascontiguousarray(f.numpy()), cfg.ffps_gamma))
                idx = orc.furthest_point_sample_matrix(mat.numpy(), npnt)
            elif method == "s-fps":
                idx = orc.furthest_point_sample_weights(sl, np.ascontiguousarray(scores[:, lo:hi]), npnt)
            else:
                raise NotImplementedError(method)
            parts.append(idx + lo)
        sample_idx = np.concatenate(parts, axis=-1).astype(np.int32)
        new_xyz = np.ascontiguousarray(orc.gather_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), sample_idx).transpose(0, 2, 1))
        out["l%d_idx" % li] = sample_idx
        for si, (r, ns) in enumerate(zip(layer.radii, layer.nsamples)):
            cnt, nf = _group(xyz, new_xyz, feats, r, ns)
            out["l%d_s%d_cnt" % (li, si)] = cnt
            if keep_groups:
                out["l%d_s%d" % (li, si)] = nf
        xyz = new_xyz
    votes = np.ascontiguousarray(xyz[:, :cfg.n_votes] + h["vote_offsets"]) if cfg.n_votes > 0 else None
    for si, (r, ns) in enumerate(zip(cfg.vote_radii, cfg.vote_nsamples) if cfg.n_votes > 0 else ()):
        cnt, nf = _group(xyz, votes, h["vote_feats"], r, ns)
        out["head_s%d_cnt" % si] = cnt
        if keep_groups:
            out["head_s%d" % si] = nf
    if cfg.n_proposals <= 0:
        return out
    F = h["boxes"].shape[0]
    keep = np.zeros((F, cfg.n_proposals), np.int64)
    num = np.zeros(F, np.int32)
    for f in range(F):
        k = orc.nms_gpu(h["boxes"][f], h["box_scores"][f], cfg.nms_thresh)
        keep[f, :len(k)] = k
        num[f] = len(k)
    out["nms_keep"], out["nms_num"] = keep, num
    return out

"""Set abstraction with the shared MLP fused behind the grouper (second half).

The reference runs, per radius scale (pcdet/ops/pointnet2/pointnet2_batch/pointnet2_modules.py:461-478):

    idx_cnt, new_features = self.groupers[i](xyz, new_xyz, features)     # (B, 3+C, npoint, nsample) written to HBM
    new_features = self.mlps[i](new_features)                            # [Conv2d 1x1, BatchNorm2d, ReLU] x L: L more round trips
    new_features *= (idx_cnt > 0)
    pooled = F.max_pool2d(new_features, kernel_size=[1, nsample])        # (B, C_out, npoint)

`FusedSAScale` does the same from (xyz, new_xyz, features) to `pooled` with ball_query_cnt + ONE kernel
(csrc/sa_mlp.cu: gather -> tcgen05 tf32 MMAs with activations in tensor memory -> mask -> max-pool): the grouped tensor and
the hidden activations never reach HBM.  Inference only (eval-mode BatchNorm is folded into the weights; no autograd).

`fuse_sa_module(module)` takes an UNMODIFIED reference SA module (PointnetSAModuleFSMSG, or anything with its attributes)
and returns a callable with the signature of its forward that runs sampling with this package's FPS kernels and every
eligible scale fused; scales the kernel cannot take (weights that do not fit in shared memory, odd widths) run the
reference composition.
"""
import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils as pu
from ._lib import call, load


def fold_mlp(mlp: nn.Sequential):
    """[Conv2d(k=1) (, BatchNorm2d) (, ReLU)] x L in eval mode -> [(W' (c_out, c_in), b' (c_out))]: y = relu(W' x + b')."""
    layers, mods, i = [], list(mlp), 0
    while i < len(mods):
        conv = mods[i]
        if not (isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.groups == 1):
            raise ValueError("unsupported layer in shared MLP: %r" % (conv,))
        w = conv.weight.detach().double().flatten(1)
        b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(w.size(0), dtype=torch.float64, device=w.device)
        i += 1
        if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
            bn = mods[i]
            if bn.training or not bn.track_running_stats:
                raise ValueError("BatchNorm must be in eval mode with running statistics to be folded")
            scale = (bn.weight.detach().double() if bn.affine else 1.0) / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            shift = (bn.bias.detach().double() if bn.affine else 0.0) - bn.running_mean.detach().double() * scale
            w, b = w * scale[:, None], b * scale + shift
            i += 1
        if not (i < len(mods) and isinstance(mods[i], nn.ReLU)):
            raise ValueError("every layer of the shared MLP must end in ReLU")
        i += 1
        layers.append((w.float().contiguous(), b.float().contiguous()))
    return layers


class FusedSAScale:
    """One radius scale of an SA layer: ball_query_cnt + fused group/MLP/mask/max-pool.

        scale = FusedSAScale(radius, nsample, mlp)            # mlp: the nn.Sequential of the reference module, eval mode
        pooled = scale(xyz, new_xyz, features)                # (B, N, 3), (B, M, 3), (B, C, N) -> (B, C_out, M)
    """

    @staticmethod
    def plan(widths, nsample: int):
        """(mode, parts): mode 1 = one CTA holds all weights, 2 = cta_group::2 pairs, 0 = cannot run fused; parts = how many
        row blocks the LAST layer is split into (one launch each, every launch recomputing the earlier layers) -- 1 unless the
        whole MLP exceeds the shared memory of an SM pair (SASA SA3's 131 -> 128 -> 256 -> 256: two blocks of 128 rows)."""
        lib = load()
        for parts in (1, 2, 4):
            if widths[-1] % parts:
                break
            w = list(widths[:-1]) + [widths[-1] // parts]
            mode = int(lib.de6d_sa_mlp_fits(len(w) - 1, (C.c_int * len(w))(*w), int(nsample)))
            if mode:
                return mode, parts
        return 0, 0

    @staticmethod
    def supported(mlp: nn.Sequential, nsample: int) -> bool:
        try:
            layers = fold_mlp(mlp)
        except ValueError:
            return False
        widths = [layers[0][0].size(1)] + [w.size(0) for w, _ in layers]
        return bool(FusedSAScale.plan(widths, nsample)[0])

    @staticmethod
    def single_cta(scale: "FusedSAScale") -> bool:
        """True when one CTA holds all weights; False when the kernel runs as cta_group::2 pairs (half of every layer per SM)."""
        return scale.mode == 1

    def __init__(self, radius: float, nsample: int, mlp: nn.Sequential, radius_in: Optional[float] = None):
        self.radius, self.nsample, self.radius_in = float(radius), int(nsample), radius_in
        layers = fold_mlp(mlp)
        self.widths = [layers[0][0].size(1)] + [w.size(0) for w, _ in layers]
        self.c_feat = self.widths[0] - 3
        lib = load()
        self.mode, self.parts = self.plan(self.widths, self.nsample)      # mode 1: one CTA, 2: CTA pairs; parts: launches over the last layer
        if not self.mode:
            raise ValueError("shared MLP %s with nsample %d cannot run fused" % (self.widths, self.nsample))
        dev = layers[0][0].device
        rows = self.widths[-1] // self.parts
        self._w = (C.c_int * len(self.widths))(*(self.widths[:-1] + [rows]))
        self.slices = []      # (packed weights, biases, first output channel) per launch
        for h in range(self.parts):
            last_w, last_b = layers[-1][0][h * rows:(h + 1) * rows], layers[-1][1][h * rows:(h + 1) * rows]
            cat = torch.cat([w.flatten() for w, _ in layers[:-1]] + [last_w.flatten()]).contiguous()
            bias = torch.cat([b for _, b in layers[:-1]] + [last_b]).contiguous()
            packed = torch.empty(int(lib.de6d_sa_mlp_packed_floats(len(layers), self._w)), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                call("de6d_sa_mlp_pack", len(layers), self._w, cat.data_ptr(), packed.data_ptr(), torch.cuda.current_stream().cuda_stream)
            self.slices.append((packed, bias, h * rows))
        self.packed, self.bias = self.slices[0][0], self.slices[0][1]
        self.n_layers = len(layers)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)

    @torch.no_grad()
    def pooled(self, xyz, new_xyz, feats_pm, idx, idx_cnt):
        """feats_pm: POINT-major features (B, N, C) (one transposition per SA layer, shared by its scales)."""
        B, N, _ = xyz.shape
        M = new_xyz.size(1)
        out = torch.empty((B, self.widths[-1], M), dtype=torch.float32, device=xyz.device)
        chk = pu._chk
        px, pn = chk(xyz, "xyz", torch.float32, (B, N, 3)), chk(new_xyz, "new_xyz", torch.float32, (B, M, 3))
        pf = None if self.c_feat == 0 else chk(feats_pm, "feats_pm", torch.float32, (B, N, self.c_feat))
        pi = chk(idx, "idx", torch.int32, (B, M, self.nsample))
        pc = None if idx_cnt is None else chk(idx_cnt, "idx_cnt", torch.int32, (B, M))
        for packed, bias, c_off in self.slices:
            call("de6d_sa_mlp_fused_slice", B, N, M, self.nsample, self.c_feat, px, pn, pf, pi, pc, self.n_layers, self._w,
                 packed.data_ptr(), bias.data_ptr(), out.data_ptr(), self.widths[-1], c_off, self.status.data_ptr(),
                 torch.cuda.current_stream().cuda_stream)
        return out

    @torch.no_grad()
    def __call__(self, xyz, new_xyz, features, grid=None, feats_pm=None):
        if self.radius_in is None:
            idx_cnt, idx = pu.ball_query_cnt(self.radius, self.nsample, xyz, new_xyz, grid=grid)
        else:
            idx_cnt, idx = pu.ball_query_dilated(self.radius_in, self.radius, self.nsample, xyz, new_xyz, grid=grid)
        if feats_pm is None and features is not None:
            feats_pm = features.transpose(1, 2).contiguous()
        return self.pooled(xyz, new_xyz, feats_pm, idx, idx_cnt)


def fuse_sa_module(module: nn.Module, ffps: str = "cdist"):
    """Returns forward(xyz, features=None, new_xyz=None, scores=None) -> (new_xyz, new_features, new_scores) equivalent to
    _PointnetSAModuleFSBase.forward (pointnet2_modules.py:358-494) of `module` in eval mode, with this package's sampling
    kernels, one ball-query grid per layer and every eligible scale fused.  `forward.fused` lists which scales are.
    ffps: "cdist" = the reference's F-FPS pipeline (torch.cdist matrix + matrix kernel: the reference's own indices),
    "fused" = de6d_furthest_point_sampling_features (no (B, N, N) matrix; direct-difference metric)."""
    if module.training:
        raise ValueError("fuse_sa_module needs module.eval(): BatchNorm statistics are folded into the weights")
    if module.pool_method != "max_pool":
        raise ValueError("only max_pool is fused")
    scales: List[Optional[FusedSAScale]] = []
    former = 0.0
    for grouper, mlp in zip(module.groupers, module.mlps):
        r_in = former if module.dilated_radius_group else None
        radius = getattr(grouper, "radius", None) if not module.dilated_radius_group else grouper.radius_out
        ok = getattr(grouper, "use_xyz", True) and FusedSAScale.supported(mlp, grouper.nsample)
        scales.append(FusedSAScale(radius, grouper.nsample, mlp, radius_in=r_in) if ok else None)
        former = radius

    @torch.no_grad()
    def forward(xyz, features=None, new_xyz=None, scores=None):
        if new_xyz is None:
            parts = []
            for (lo, hi), method, npnt in zip(module.sample_range_list, module.sample_method_list, module.npoint_list):
                xs = xyz[:, lo:hi, :].contiguous()
                if method == "d-fps":
                    sidx = pu.furthest_point_sample(xs, npnt)
                elif method == "f-fps":
                    fs = features[:, :, lo:hi].permute(0, 2, 1)
                    if ffps == "fused":
                        sidx = pu.furthest_point_sample_features(xs, fs, module.weight_gamma, npnt)
                    else:
                        mat = torch.cdist(xs, xs)
                        mat += torch.cdist(fs, fs) * module.weight_gamma
                        sidx = pu.furthest_point_sample_matrix(mat, npnt)
                elif method == "s-fps":
                    w = scores[:, lo:hi].contiguous().sigmoid() ** module.weight_gamma
                    sidx = pu.furthest_point_sample_weights(xs, w.contiguous(), npnt)
                else:
                    raise NotImplementedError("sampling method %r" % method)
                parts.append(sidx + lo)
            sample_idx = torch.cat(parts, dim=-1)
            new_xyz = pu.gather_xyz(xyz, sample_idx)
            old_features = pu.gather_operation(features, sample_idx) if (module.skip_connection and features is not None) else None
        else:
            old_features = None
        grid = pu.BallQueryGrid(xyz, min(s.radius if s else g.radius for s, g in zip(scales, module.groupers))) \
            if (pu.BallQueryGrid.wanted(xyz.size(1)) and not module.dilated_radius_group) else None
        feats_pm = features.transpose(1, 2).contiguous() if (features is not None and any(scales)) else None
        outs = []
        for scale, grouper, mlp in zip(scales, module.groupers, module.mlps):
            if scale is not None:
                outs.append(scale(xyz, new_xyz, features, grid=grid, feats_pm=feats_pm))
            else:       # the reference composition for this scale
                idx_cnt, nf = grouper(xyz, new_xyz, features)
                nf = mlp(nf) * (idx_cnt > 0).float().unsqueeze(1).unsqueeze(-1)
                outs.append(F.max_pool2d(nf, kernel_size=[1, nf.size(3)]).squeeze(-1))
        if module.skip_connection and old_features is not None:
            outs.append(old_features)
        new_features = torch.cat(outs, dim=1)
        if module.aggregation_mlp is not None:
            new_features = module.aggregation_mlp(new_features)
        if module.confidence_mlp is not None:
            return new_xyz, new_features, module.confidence_mlp(new_features).squeeze(1)
        return new_xyz, new_features, None

    forward.fused = [s is not None for s in scales]
    forward.scales = scales
    return forward


def fuse_backbone(backbone: nn.Module, ffps: str = "cdist"):
    """Returns forward(batch_dict) -> batch_dict equivalent to PointNet2FSMSG.forward (pcdet/models/backbones_3d/
    pointnet2_backbone.py:199-263) of an UNMODIFIED reference backbone in eval mode: input staging with one kernel
    (de6d_b200.staging.break_up_pc instead of slicing copies + B host-synchronising .sum() calls), every SA module through
    fuse_sa_module, FP modules (three_nn / three_interpolate + their MLPs) as they are.  Same keys as the reference."""
    from . import staging
    if backbone.training:
        raise ValueError("fuse_backbone needs backbone.eval()")
    sa = [fuse_sa_module(m, ffps=ffps) for m in backbone.SA_modules]

    @torch.no_grad()
    def forward(batch_dict):
        batch_size = batch_dict["batch_size"]
        batch_idx, xyz, features = staging.break_up_pc(batch_dict["points"], batch_size)
        l_xyz, l_features, l_scores = [xyz], [features], [None]
        for i, fwd in enumerate(sa):
            li_xyz, li_features, li_scores = fwd(l_xyz[i], l_features[i], scores=l_scores[i])
            l_xyz.append(li_xyz); l_features.append(li_features); l_scores.append(li_scores)
        batch_dict["point_coords_list"] = [torch.cat([batch_idx[:, :x.size(1)].reshape(-1, 1), x.reshape(-1, 3)], dim=1) for x in l_xyz[1:]]
        batch_dict["point_scores_list"] = [None if sc is None else sc.reshape(-1, 1) for sc in l_scores[1:]]
        i = 0
        if backbone.FP_modules is not None:
            for i in range(-1, -(len(backbone.FP_modules) + 1), -1):
                l_features[i - 1] = backbone.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
        point_features = l_features[i - 1].permute(0, 2, 1).contiguous()
        batch_dict["point_features"] = point_features.view(-1, point_features.shape[-1])
        batch_dict["point_coords"] = torch.cat((batch_idx[:, :l_xyz[i - 1].size(1)].reshape(-1, 1).float(), l_xyz[i - 1].view(-1, 3)), dim=1)
        batch_dict["point_scores"] = l_scores[-1]
        return batch_dict

    forward.fused = [f.fused for f in sa]
    return forward

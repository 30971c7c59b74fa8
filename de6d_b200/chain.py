"""The SA + NMS op chain of BASELINE.json, built only from this package's ops.

It is the sequence of hot-path calls one Det6D / SASA inference forward makes on a batch of frames
(SURVEY.md section 3.1): per SA layer  FPS (D / F / S) -> gather_operation -> per radius scale
ball_query_cnt -> grouping_operation(xyz) -> centre subtraction -> grouping_operation(features) -> concat,
then the head's vote-centre grouping and the rotated NMS over the frame's proposals.  The learned parts in
between (shared MLPs, vote / box regression) are out of this path's scope; their outputs -- per-layer
features, S-FPS scores, proposals -- are synthetic inputs of the right shape (de6d_b200.synth).

Data dependencies follow the real network: F-FPS / S-FPS of layer k+1 wait for the groupings of layer k (they
consume features / scores computed from them), D-FPS of layer k+1 only needs layer k's sampled coordinates and
overlaps with layer k's grouping on a side stream.  The whole step is captured once into a CUDA graph.

Layer shapes: the reference ships no YAML; the defaults below are the SASA/3DSSD shapes named
in BASELINE.json configs[1] (16384 -> 4096 -> 1024 -> 512, 256 vote centres) and are plain parameters.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import pointnet2_utils as pu
from .iou3d_nms_utils import BatchedNMS


@dataclass
class SALayer:
    npoints: Tuple[int, ...]                  # points sampled by each method
    methods: Tuple[str, ...]                  # 'd-fps' | 'f-fps' | 's-fps'
    ranges: Tuple[Tuple[int, int], ...]       # slice of the layer input each method samples from
    radii: Tuple[float, ...]
    nsamples: Tuple[int, ...]
    c_in: int                                 # feature channels entering the layer


@dataclass
class ChainConfig:
    n_points: int = 16384
    layers: List[SALayer] = field(default_factory=lambda: [
        SALayer((4096,), ('d-fps',), ((0, 16384),), (0.2, 0.4, 0.8), (32, 32, 64), 1),
        SALayer((512, 512), ('f-fps', 'd-fps'), ((0, 4096), (0, 4096)), (0.4, 0.8, 1.6), (32, 32, 64), 64),
        SALayer((256, 256), ('s-fps', 'd-fps'), ((0, 512), (512, 1024)), (1.6, 3.2, 4.8), (32, 32, 32), 128),
    ])
    n_votes: int = 256
    vote_radii: Tuple[float, ...] = (4.8, 6.4)
    vote_nsamples: Tuple[int, ...] = (16, 32)
    vote_c_in: int = 256
    n_proposals: int = 512
    nms_thresh: float = 0.01
    ffps_gamma: float = 1.0
    cloud: str = "uniform"                    # synthetic generator: "uniform" (KITTI crop) | "lidar" (64-beam rings)

    def name(self):
        return "sasa-3dssd-chain-%d" % self.n_points


def stress_config():
    """BASELINE.json configs[4]: 131072-point 64-beam LiDAR frames, D-FPS to 16384, ball_query ns = 64 + grouping
    (one SA layer, C = 1; no head, no NMS)."""
    return ChainConfig(n_points=131072, cloud="lidar",
                       layers=[SALayer((16384,), ('d-fps',), ((0, 131072),), (0.2,), (64,), 1)],
                       n_votes=0, vote_radii=(), vote_nsamples=(), n_proposals=0)


def small_config():
    """A few-thousand-point version for tests and smoke()."""
    return ChainConfig(
        n_points=2048,
        layers=[
            SALayer((512,), ('d-fps',), ((0, 2048),), (0.8, 1.6), (16, 32), 1),
            SALayer((128, 128), ('f-fps', 'd-fps'), ((0, 512), (0, 512)), (1.6, 3.2), (16, 32), 16),
            SALayer((64, 64), ('s-fps', 'd-fps'), ((0, 128), (128, 256)), (3.2, 4.8), (16, 16), 32),
        ],
        n_votes=64, vote_radii=(4.8,), vote_nsamples=(16,), vote_c_in=32, n_proposals=128)


def make_inputs(cfg: ChainConfig, batch: int, seed: int = 0, pinned: bool = True):
    """Host-side inputs of one step (numpy-seeded, identical on every machine)."""
    from . import synth
    import numpy as np
    xyz = (synth.lidar_clouds if cfg.cloud == "lidar" else synth.clouds)(batch, cfg.n_points, seed)
    feats = [synth.features(batch, cfg.layers[0].c_in, cfg.n_points, seed)]
    n_in = cfg.n_points
    scores = []
    for li, layer in enumerate(cfg.layers):
        n_out = sum(layer.npoints)
        if li + 1 < len(cfg.layers):
            feats.append(synth.features(batch, cfg.layers[li + 1].c_in, n_out, seed + 10 * (li + 1)))
        scores.append(synth.weights(batch, n_in, seed + 77 + li))   # S-FPS weights = sigmoid(score) ** gamma
        n_in = n_out
    host = {"xyz": xyz}
    if cfg.n_votes > 0:
        host["vote_feats"] = synth.features(batch, cfg.vote_c_in, n_in, seed + 99)
        host["vote_offsets"] = np.random.default_rng(seed + 5).normal(0, 0.5, size=(batch, cfg.n_votes, 3)).astype(np.float32)
    if cfg.n_proposals > 0:
        host["boxes"], host["box_scores"] = synth.proposals(batch, cfg.n_proposals, seed)
    for i, f in enumerate(feats):
        host["feats%d" % i] = f
    for i, s in enumerate(scores):
        host["scores%d" % i] = s
    out = {}
    for k, v in host.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        out[k] = t.pin_memory() if pinned and torch.cuda.is_available() else t
    return out


class OpChain:
    """One process, one GPU, `batch` frames per step.

        chain = OpChain(cfg, batch)                 # allocates static device buffers
        out = chain.step_host(host_inputs)          # H2D -> graph replay -> D2H, returns host tensors
        chain.load(host_inputs); chain.step()       # device-resident step (inputs already in HBM)
    """

    N_SIDE = 4

    def __init__(self, cfg: ChainConfig, batch: int, device: Optional[torch.device] = None, use_graph: bool = True,
                 keep_matrices: bool = False, serial: bool = False, fused_group: bool = True,
                 fused_ffps: bool = True, ffps: Optional[str] = None, shared_grid: bool = True, ffps_cluster: int = 0):
        self.cfg, self.batch = cfg, batch
        self.keep_matrices = keep_matrices   # tests: expose the F-FPS distance matrices fed to the kernel
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.use_graph = use_graph
        self.fused_group = fused_group
        # F-FPS route: "fused" (no (B,N,N) matrix; direct-difference metric), "matrix" (this library's matrix kernel + FPS
        # kernel, same metric, same indices as "fused"), "cdist" (torch.cdist + FPS kernel = the reference pipeline: its
        # fp32 GEMM-expansion metric, bit-identical indices to the reference on the same GPU)
        self.ffps = ffps or ("fused" if fused_ffps else "matrix")
        self.shared_grid = shared_grid
        # thread-block cluster form of the fused F-FPS kernel: 0 = the launcher's choice (least time for ONE launch of this batch);
        # 44 = 4-CTA clusters regardless (least SM-time per cloud: what pays when several chains are in flight and the SMs, not
        # one launch's latency, are the limit -- e.g. 16-frame batches); identical indices either way
        self.ffps_cluster = ffps_cluster
        self._grids = {}
        self.main = torch.cuda.Stream(self.device)
        if serial:   # everything on one stream (per-kernel timing pass of bench.py, ncu launch lists)
            self.side = [self.main] * self.N_SIDE
            self.samp = [self.main] * 2
        else:
            self.side = [torch.cuda.Stream(self.device) for _ in range(self.N_SIDE)]   # grouping, one per scale
            self.samp = [torch.cuda.Stream(self.device) for _ in range(2)]             # extra sampling methods
        self.inputs = None          # static device copies
        self.outputs = None         # static device outputs of the last captured step
        self.graph = None
        self.host_out = None
        self.kernels_per_step = None
        self.nms = None

    # ---- buffers ------------------------------------------------------------------------------------
    def _alloc_like(self, host):
        with torch.cuda.device(self.device):
            self.inputs = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            if self.cfg.n_proposals > 0:
                with torch.cuda.stream(self.main):     # the workspace is zeroed on the stream the kernel will run on
                    self.nms = BatchedNMS(self.batch, self.cfg.n_proposals, device=self.device)

    SENSOR_KEYS = ("xyz", "feats0")   # what a deployment copies per frame: points + intensity

    def load(self, host, stream=None, keys=None):
        if self.inputs is None:
            self._alloc_like(host)
        s = stream or self.main
        with torch.cuda.stream(s):
            for k, v in host.items():
                if keys is None or k in keys:
                    self.inputs[k].copy_(v, non_blocking=True)

    def h2d_bytes(self, keys=None):
        return sum(v.numel() * v.element_size() for k, v in self.inputs.items() if keys is None or k in keys)

    # ---- the chain --------------------------------------------------------------------------------
    def _group_scales(self, xyz, xyz_t, new_xyz, feats, radii, nsamples, after, outs, tag):
        """One QueryWithCntAndGroup per radius scale, scales spread over the side streams.  The scales share one
        ball-query grid of `xyz` (built on the main stream before `after`) and the transposed cloud `xyz_t`."""
        done = []
        grid = self._grids.get(tag)
        for si, (r, ns) in enumerate(zip(radii, nsamples)):
            st = self.side[si % self.N_SIDE]
            st.wait_event(after)
            with torch.cuda.stream(st):
                idx_cnt, idx = pu.ball_query_cnt(r, ns, xyz, new_xyz, grid=grid)
                if self.fused_group:      # QueryWithCntAndGroup's tail as one pass (what pu.QueryWithCntAndGroup runs)
                    new_features = pu.group_concat(xyz, new_xyz, feats, idx, xyz_t=xyz_t)
                else:                     # the reference's op-by-op composition (pointnet2_utils.py:410-417)
                    xyz_tr = xyz.transpose(1, 2).contiguous()
                    g_xyz = pu.grouping_operation(xyz_tr, idx)
                    g_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
                    g_feat = pu.grouping_operation(feats, idx)
                    new_features = torch.cat([g_xyz, g_feat], dim=1)
                outs["%s_s%d_cnt" % (tag, si)] = idx_cnt
                outs["%s_s%d" % (tag, si)] = new_features
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
        return done

    def _sampled_xyz(self, xyz, sample_idx):
        """new_xyz (B, M, 3) and its transposed copy (B, 3, M) for the next layer's grouping."""
        if self.fused_group:
            return pu.gather_xyz(xyz, sample_idx, want_transposed=True)
        xyz_flipped = xyz.transpose(1, 2).contiguous()                 # the reference's sequence (pointnet2_modules.py:374,451-454)
        new_t = pu.gather_operation(xyz_flipped, sample_idx)
        return new_t.transpose(1, 2).contiguous(), new_t

    def _build_grid(self, tag, xyz, radii):
        """Called on the main stream right before the `after` event of _group_scales."""
        self._grids[tag] = pu.BallQueryGrid(xyz, min(radii)) if (self.shared_grid and pu.BallQueryGrid.wanted(xyz.size(1))) else None

    def _forward(self):
        cfg, inp = self.cfg, self.inputs
        outs = {}
        main = self.main
        xyz = inp["xyz"]
        self._grids = {}
        grouped_done: List[torch.cuda.Event] = []
        pending: List[torch.cuda.Event] = []   # every side-stream event main has not joined yet
        with torch.cuda.stream(main):
            xyz_t = pu.gather_xyz(xyz, None)[1] if self.fused_group else None   # (B, 3, N) once per step
            for li, layer in enumerate(cfg.layers):
                feats = inp["feats%d" % li]
                scores = inp["scores%d" % li]
                start = torch.cuda.Event()
                start.record(main)
                idx_parts = [None] * len(layer.methods)
                joins = []
                for mi, (method, npnt, (lo, hi)) in enumerate(zip(layer.methods, layer.npoints, layer.ranges)):
                    st = main if mi == 0 else self.samp[(mi - 1) % len(self.samp)]
                    if st is not main:
                        st.wait_event(start)
                    needs_prev_features = method in ("f-fps", "s-fps")
                    with torch.cuda.stream(st):
                        if needs_prev_features:
                            for ev in grouped_done:
                                st.wait_event(ev)
                        xyz_slice = xyz[:, lo:hi, :].contiguous()
                        if method == "d-fps":
                            sidx = pu.furthest_point_sample(xyz_slice, npnt)
                        elif method == "f-fps":
                            f_slice = feats[:, :, lo:hi].permute(0, 2, 1)
                            if self.ffps == "fused" and not self.keep_matrices:
                                cl = self.ffps_cluster if (self.ffps_cluster and f_slice.size(2) == 64 and 3072 < xyz_slice.size(1) <= 4096) else 0
                                sidx = pu.furthest_point_sample_features(xyz_slice, f_slice, cfg.ffps_gamma, npnt, cluster_size=cl)
                            elif self.ffps == "cdist":   # exactly the reference's call pair (pointnet2_modules.py:383-388)
                                mat = torch.cdist(xyz_slice, xyz_slice)
                                mat += torch.cdist(f_slice, f_slice) * cfg.ffps_gamma
                                sidx = pu.furthest_point_sample_matrix(mat, npnt)
                            else:   # the call pair with this library's direct-difference matrix kernel
                                mat = pu.calc_dist_matrix_for_sampling(xyz_slice, f_slice, cfg.ffps_gamma)
                                sidx = pu.furthest_point_sample_matrix(mat, npnt)
                                if self.keep_matrices:
                                    outs["l%d_ffps_matrix" % li] = mat
                        elif method == "s-fps":
                            w = scores[:, lo:hi].contiguous()   # already sigmoid(score) ** gamma (caller side)
                            sidx = pu.furthest_point_sample_weights(xyz_slice, w, npnt)
                        else:
                            raise NotImplementedError(method)
                        idx_parts[mi] = sidx + lo
                        idx_parts[mi].record_stream(main)
                        if st is not main:
                            ev = torch.cuda.Event()
                            ev.record(st)
                            joins.append(ev)
                for ev in joins:
                    main.wait_event(ev)
                sample_idx = torch.cat(idx_parts, dim=-1)
                new_xyz, new_xyz_t = self._sampled_xyz(xyz, sample_idx)
                outs["l%d_idx" % li] = sample_idx
                outs["l%d_new_xyz" % li] = new_xyz   # also keeps the block alive while side streams read it
                if layer.radii:
                    self._build_grid("l%d" % li, xyz, layer.radii)
                sampled = torch.cuda.Event()
                sampled.record(main)
                grouped_done = self._group_scales(xyz, xyz_t, new_xyz, feats, layer.radii, layer.nsamples, sampled, outs, "l%d" % li)
                pending.extend(grouped_done)
                outs["l%d_xyz_t" % li] = xyz_t       # keeps the block alive while side streams read it
                xyz, xyz_t = new_xyz, (new_xyz_t if self.fused_group else None)
            # head: vote centres grouped over the last layer's points (after its features exist)
            for ev in pending:
                main.wait_event(ev)
            if cfg.n_votes > 0:
                votes = (xyz[:, :cfg.n_votes, :] + inp["vote_offsets"]).contiguous()
                outs["votes"] = votes
                self._build_grid("head", xyz, cfg.vote_radii)
                ready = torch.cuda.Event()
                ready.record(main)
                head_done = self._group_scales(xyz, xyz_t, votes, inp["vote_feats"], cfg.vote_radii, cfg.vote_nsamples, ready, outs, "head")
                outs["head_xyz_t"] = xyz_t
                for ev in head_done:
                    main.wait_event(ev)
            if cfg.n_proposals > 0:
                keep, num = self.nms(inp["boxes"], inp["box_scores"], cfg.nms_thresh)
                outs["nms_keep"], outs["nms_num"] = keep, num
        return outs

    def capture(self):
        """Warm up eagerly once (sets kernel attributes, fills allocator pools), then capture the graph."""
        assert self.inputs is not None, "call load() first"
        from ._lib import launch_count
        with torch.cuda.device(self.device):
            c0 = launch_count()
            self.outputs = self._forward()
            self.kernels_per_step = launch_count() - c0
            torch.cuda.synchronize(self.device)
            if self.use_graph:
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.main):
                    self.outputs = self._forward()
                torch.cuda.synchronize(self.device)
            res = self.result_tensors()
            self.host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in res.items()}

    def step(self):
        """One pass over the batch with inputs already resident in HBM (enqueue only, no sync)."""
        if self.graph is not None:
            with torch.cuda.stream(self.main):
                self.graph.replay()
        else:
            self.outputs = self._forward()

    def result_tensors(self):
        """What a caller reads back per step: sampled indices of the last layer and the detections kept."""
        last = len(self.cfg.layers) - 1
        res = {"sample_idx": self.outputs["l%d_idx" % last]}
        if self.cfg.n_proposals > 0:
            res["nms_keep"], res["nms_num"] = self.outputs["nms_keep"], self.outputs["nms_num"]
        return res

    def d2h_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.result_tensors().values())

    def step_host(self, host, sync=True, keys=None):
        """Public end-to-end call: pinned host inputs -> device, run, results -> pinned host.  keys: copy only these
        inputs (e.g. SENSOR_KEYS: in the real detector the per-layer features, S-FPS scores and proposals are products
        of the MLPs / head on the device; here they are synthetic stand-ins that `load` made resident)."""
        self.load(host, keys=keys)
        if self.graph is None and self.outputs is None:
            self.capture()
        self.step()
        with torch.cuda.stream(self.main):
            for k, v in self.result_tensors().items():
                self.host_out[k].copy_(v, non_blocking=True)
        if sync:
            self.main.synchronize()
        return self.host_out

#!/usr/bin/env python
"""Benchmark of the Det6D SA + NMS op chain (BASELINE.json metric: frames/s for the 16384-point chain).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]          # this package on N B200s
    python bench.py --impl reference [...]                                    # the CPU path on the host cores

A step = one pass of the op chain (de6d_b200.chain.OpChain: 3 SA layers with D-/F-/S-FPS, gather, ball_query_cnt
+ grouping per radius scale, the head's vote grouping, batched rotated NMS on 512 proposals/frame) over one
batch of synthetic KITTI-shape frames.  Per GPU the batch is BASELINE.json configs[2] (64 frames); frames are
sharded across ranks with no data-path collective (weak scaling: 8 GPUs x 64 = configs[3]'s 512 frames).

One JSON line on stdout (rank 0):
  value     frames/s, inputs resident in HBM, CUDA-graph replay, CUDA events on the launching stream, max over ranks
  e2e       frames/s through OpChain.step_host: pinned host sensor inputs (points + intensity) -> H2D -> chain -> D2H of
            the results, every step; the stand-ins for tensors the detector's MLPs produce ON the device stay resident
            (`e2e_all_inputs` copies those too, every step)
  roofline  the HBM-bound kernel that moves most of the step's bytes against the measured HBM peak, and the kernel that
            takes most of the step's time (issue/latency-bound sampling) as fraction of the GPU's issue slots;
            `kernels` lists every entry point
  cpu_baseline   the oracle (CPU restatement of the reference kernels) on a bounded sample of the same workload
  reference_cuda the same chain issued op by op with the reference's OWN CUDA kernels (oracle/_ref, recompiled for sm_100)
            through the reference's own python, on this GPU -- the yardstick SURVEY.md names; outside every timed region
  backbone  the reference's own PointNet2FSMSG backbone (SA layers with their MLPs), 16 frames: reference python over the
            reference kernels / over compat / with fused SA scales;  sa_mlp_fused: SA layers fused vs unfused, 64 frames
  other_configs  BASELINE.json configs[1] (16 frames) and configs[4] (131072-point stress frames), short runs
oracle/ is used here only for cpu_baseline / reference_cuda / the --impl reference arm, never on the measured GPU path.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames_per_second_16384pt_SA_NMS_op_chain"
UNIT = "frames/s"
L2_BYTES = 126 * 1024 * 1024


# ------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(entry, shape):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel behind a C-ABI entry point,
    from the committed `ncu --set full` captures (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py from the
    .ncu-rep of the same workload); None when that shape was not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    try:
        with open(p) as f:
            tab = json.load(f)
    except Exception:
        return None
    rec = tab.get(entry + ":" + ",".join(str(x) for x in shape))
    return None if rec is None else rec.get("dram_bytes")


def algorithmic_bytes(name, a):
    """Compulsory HBM bytes of one launch (DESIGN.md 'Algorithmic bytes'), from the C-ABI arguments."""
    if name in ("de6d_furthest_point_sampling", "de6d_furthest_point_sampling_impl"):
        b, n, m = a[0], a[1], a[2]
        return b * (12 * n + 4 * m)
    if name in ("de6d_furthest_point_sampling_weights", "de6d_furthest_point_sampling_weights_impl"):
        b, n, m = a[0], a[1], a[2]
        return b * (16 * n + 4 * m)
    if name == "de6d_furthest_point_sampling_matrix":
        b, n, m = a[0], a[1], a[2]
        return b * (4 * n * m + 4 * m)
    if name == "de6d_furthest_point_sampling_features":
        b, n, c, m = a[0], a[1], a[2], a[3]
        return b * (12 * n + 4 * n * c + 4 * m)
    if name == "de6d_gather_points":
        b, c, n, npnt = a[0], a[1], a[2], a[3]
        return b * (4 * npnt + 8 * c * npnt)
    if name == "de6d_ball_query_ex":
        mode, b, n, m, ns = a[0], a[2], a[3], a[4], a[7]
        return b * (12 * n + 12 * m + 4 * m * ns + (4 * m if mode else 0))
    if name == "de6d_dist_matrix":
        b, n, c = a[0], a[1], a[2]
        return b * (12 * n + 4 * n * c + 4 * n * n)
    if name in ("de6d_ball_query", "de6d_ball_query_cnt"):
        b, n, m, ns = a[0], a[1], a[2], a[4]
        return b * (12 * n + 12 * m + 4 * m * ns + (4 * m if name.endswith("cnt") else 0))
    if name == "de6d_ball_query_dilated":
        b, n, m, ns = a[0], a[1], a[2], a[5]
        return b * (12 * n + 12 * m + 4 * m * ns + 4 * m)
    if name in ("de6d_group_points", "de6d_group_points_impl"):
        b, c, n, npnt, ns = a[0], a[1], a[2], a[3], a[4]
        return b * (4 * npnt * ns + 4 * c * min(n, npnt * ns) + 4 * c * npnt * ns)
    if name in ("de6d_group_concat", "de6d_group_concat_t"):
        b, c, n, npnt, ns = a[0], a[1], a[2], a[3], a[4]
        touched = min(n, npnt * ns)
        return b * (4 * npnt * ns + 12 * touched + 12 * npnt + 4 * c * touched + 4 * (3 + c) * npnt * ns)
    if name == "de6d_gather_xyz":
        b, n, m = a[0], a[1], a[2]
        return b * (4 * m + 12 * m + 12 * m + 12 * m)      # idx + gathered points read, both layouts written
    if name == "de6d_ball_query_grid_build":
        b, n = a[0], a[1]
        return b * (12 * n + 16 * n)                        # cloud read, sorted records written
    if name == "de6d_nms_batched":
        frames, n = a[0], a[1]
        return frames * 36 * n
    return 0


class ClockSampler:
    """Samples SM clock and clock-event (throttle) reasons of one GPU while the timed regions run."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x100: "display_clocks_setting",
               0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_one_frame(args):
    """Worker: the oracle chain on one frame (single thread)."""
    batch_seed, frame = args
    import torch
    torch.set_num_threads(1)
    from de6d_b200 import chain as ch
    from oracle import chain_ref
    host = _cpu_one_frame.cache.get(batch_seed)
    if host is None:
        host = ch.make_inputs(ch.ChainConfig(), _cpu_one_frame.batch, seed=batch_seed, pinned=False)
        _cpu_one_frame.cache = {batch_seed: host}
    t0 = time.perf_counter()
    chain_ref.run_chain(ch.ChainConfig(), host, frames=slice(frame, frame + 1), ffps_matrix="torch")
    return time.perf_counter() - t0


_cpu_one_frame.cache = {}
_cpu_one_frame.batch = 16


def cpu_baseline_single_core(budget_s=12.0, max_frames=16):
    """Oracle port, one core, frame after frame until ~budget_s of CPU work is done."""
    from oracle import oracle
    oracle.build()
    _cpu_one_frame((0, 0))  # warm-up: builds inputs, pages the library in
    t, frames = 0.0, 0
    while t < budget_s and frames < max_frames:
        t += _cpu_one_frame((0, frames % _cpu_one_frame.batch))
        frames += 1
    return {"value": frames / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d frames of the same chain config, oracle (C restatement of the reference kernels, -O2) on one host core, %.1f s" % (frames, t)}


def run_reference_arm(args):
    """--impl reference: the CPU path on all host cores.  The reference ships no CPU implementation of FPS, ball
    query, grouping or NMS (CUDA only), so the CPU arm is the oracle port run frame-parallel over a process pool."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    from oracle import oracle
    oracle.build()
    cores = len(os.sched_getaffinity(0))
    per_step = cores
    ctx = mp.get_context("fork")
    with ProcessPoolExecutor(max_workers=cores, mp_context=ctx) as ex:
        def step(nframes):
            t0 = time.perf_counter()
            list(ex.map(_cpu_one_frame, [(0, i % _cpu_one_frame.batch) for i in range(nframes)]))
            return time.perf_counter() - t0
        step(cores)                                     # pool start-up + input generation in every worker
        t_probe = step(per_step)
        budget = 240.0
        total_steps = args.steps + max(args.warmup - 1, 0)
        if t_probe * total_steps > budget:              # keep the whole run within a few minutes
            per_step = max(1, int(per_step * budget / (t_probe * total_steps)))
        for _ in range(max(args.warmup - 1, 0)):
            step(per_step)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(per_step)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, config_of(args)[1], max(args.gpus, 1),
                                  note="CPU arm: each step is a bounded sample of %d frames of this workload" % per_step),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d frames per step, one oracle process per host core (%d), %d steps" % (per_step, cores, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.emit(json.dumps(line))
    return 0


CONFIGS = {
    # name: (BASELINE.json configs index, default frames per GPU)
    "chain64": (2, 64),          # configs[2]; x8 GPUs = configs[3] (512 frames)
    "chain16": (1, 16),          # configs[1]
    "stress131072": (4, 4),      # configs[4]
}


def config_of(args):
    from de6d_b200 import chain as ch
    name = getattr(args, "config", "chain64")
    cfg = ch.stress_config() if name == "stress131072" else ch.ChainConfig()
    batch = args.batch if getattr(args, "batch", None) else CONFIGS[name][1]
    return cfg, batch


def workload_config(args, batch, world, note=None):
    cfg, _ = config_of(args)
    name = getattr(args, "config", "chain64")
    if name == "stress131072":
        what = ("large-cloud stress chain (BASELINE configs[4]): 131072-point 64-beam LiDAR frames, D-FPS 131072->16384 (8-CTA "
                "cluster per cloud) + ball_query_cnt r=0.2 ns=64 + grouping; %d frames per GPU" % batch)
    else:
        what = ("Det6D/SASA SA chain 16384->4096->1024->512 (D-FPS+F-FPS+S-FPS, ball_query_cnt+group per scale) "
                "+ vote grouping + rotated NMS on %d proposals/frame; %d frames per GPU (BASELINE configs[%d]%s)"
                % (cfg.n_proposals, batch, CONFIGS[name][0], "; x8 GPUs = configs[3]" if name == "chain64" else ""))
    c = {
        "workload": what, "name": name,
        "points_per_frame": cfg.n_points, "frames_per_gpu": batch, "global_frames": batch * world,
        "sharding": "frames x%d, no data-path collective" % world,
        "layers": [{"npoints": l.npoints, "methods": l.methods, "radii": l.radii, "nsamples": l.nsamples, "c_in": l.c_in}
                   for l in cfg.layers],
        "nms_thresh": cfg.nms_thresh,
    }
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------ GPU arm
def ncu_issue(entry, shape):
    """Issue-slot statistics of the kernel behind a latency/issue-bound entry point from the committed ncu captures
    (profiles/ncu_issue.json: issue-active fraction of the SMs that ran the kernel, SMs occupied, cycles per sample)."""
    p = os.path.join(ROOT, "profiles", "ncu_issue.json")
    try:
        with open(p) as f:
            tab = json.load(f)
    except Exception:
        return None
    return tab.get(entry + ":" + ",".join(str(x) for x in shape))


class ChainRunner:
    """P independent chains (own static buffers, CUDA graph and streams) fed round-robin: step k+1 starts its latency-bound
    sampling while step k is in its bandwidth-bound grouping.  All timing is CUDA events on a stream that fences every
    chain's stream, between barriers, max over ranks."""

    def __init__(self, cfg, batch, dev, P, rank, world, **chain_kw):
        import torch
        from de6d_b200 import chain as ch
        self.torch, self.cfg, self.batch, self.dev, self.P, self.world = torch, cfg, batch, dev, P, world
        self.hosts = [ch.make_inputs(cfg, batch, seed=rank * P + i) for i in range(P)]
        self.chains = []
        for i in range(P):
            c = ch.OpChain(cfg, batch, device=dev, **chain_kw)
            c.load(self.hosts[i])
            c.capture()
            self.chains.append(c)
        self.tstream = torch.cuda.Stream(dev)
        torch.cuda.synchronize()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, step_fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.tstream)
        for c in self.chains:
            c.main.wait_event(e0)
        for k in range(steps):
            step_fn(k)
        for c in self.chains:
            ev = torch.cuda.Event()
            ev.record(c.main)
            self.tstream.wait_event(ev)
        e1.record(self.tstream)
        self.barrier()
        return e0.elapsed_time(e1)

    def resident(self, steps, warmup):
        from de6d_b200 import dist as ddist
        self.timed(lambda k: self.chains[k % self.P].step(), warmup)
        return ddist.max_over_ranks(self.timed(lambda k: self.chains[k % self.P].step(), steps), self.dev)

    def e2e(self, steps, keys):
        from de6d_b200 import dist as ddist

        def step(k):
            c = self.chains[k % self.P]
            c.main.synchronize()            # this chain's previous step is complete: its host results are readable
            c.step_host(self.hosts[k % self.P], sync=False, keys=keys)
        self.timed(step, 3)
        return ddist.max_over_ranks(self.timed(step, steps), self.dev)


def per_kernel_pass(cfg, batch, dev, host, reps=5):
    """The same chain, one stream, eager, CUDA events around every C-ABI call.  Returns (kernels[], total ms per step)."""
    import torch
    from de6d_b200 import _lib, chain as ch
    prof = ch.OpChain(cfg, batch, device=dev, use_graph=False, serial=True)
    prof.load(host)
    prof.capture()                      # eager warm-up (use_graph=False: nothing is captured)
    torch.cuda.synchronize()
    _lib.trace_begin()
    for _ in range(reps):
        prof.step()
    trace = _lib.trace_end()
    agg = {}
    for name, a, t in trace:
        shape = []
        for x in a:                      # leading sizes / scalars of the C-ABI call, up to the first pointer
            if x is None or (isinstance(x, int) and abs(x) >= (1 << 31)):
                break
            shape.append(round(x, 4) if isinstance(x, float) else x)
        d = agg.setdefault((name, tuple(shape)), {"ms": 0.0, "launches": 0, "bytes": algorithmic_bytes(name, a)})
        d["ms"] += t
        d["launches"] += 1
    peak, _ = measured_peaks()
    kernels = []
    for (name, shape), d in agg.items():
        avg = d["ms"] / d["launches"]
        gbs = d["bytes"] / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
        kernels.append({"entry": name, "shape": list(shape), "ms_per_launch": round(avg, 5),
                        "launches_per_step": d["launches"] // reps, "alg_bytes": d["bytes"],
                        "gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)})
    kernels.sort(key=lambda k: -k["ms_per_launch"] * k["launches_per_step"])
    total_ms = sum(k["ms_per_launch"] * k["launches_per_step"] for k in kernels)
    del prof
    return kernels, total_ms


HBM_BOUND = ("de6d_group_points", "de6d_group_concat", "de6d_group_concat_t", "de6d_gather_points", "de6d_gather_xyz",
             "de6d_furthest_point_sampling_matrix", "de6d_three_interpolate")


def build_roofline(kernels, total_ms, step_ms):
    """Top level: the HBM-bound kernel with the most algorithmic bytes per step (the roofline north_star targets at
    >= 60 %), and -- as *_by_time -- the kernel with the most time per step, which is an issue/latency-bound sampling
    kernel whose HBM fraction is meaningless (it reads its 12.6 MB once and then iterates on chip)."""
    peak, peak_src = measured_peaks()
    by_time = kernels[0]
    hbm = [k for k in kernels if k["entry"] in HBM_BOUND]
    by_bytes = max(hbm, key=lambda k: k["alg_bytes"] * k["launches_per_step"]) if hbm else by_time
    step_bytes = sum(k["alg_bytes"] * k["launches_per_step"] for k in kernels)
    r = {"kernel": by_bytes["entry"], "shape": by_bytes["shape"], "bound": "hbm", "achieved": by_bytes["gbs"], "peak": peak,
         "unit": "GB/s", "frac": by_bytes["frac_hbm"], "traffic": ncu_traffic(by_bytes["entry"], by_bytes["shape"]),
         "alg_bytes": by_bytes["alg_bytes"], "ms_per_launch": by_bytes["ms_per_launch"], "peak_source": peak_src,
         "share_of_step_bytes": round(by_bytes["alg_bytes"] * by_bytes["launches_per_step"] / step_bytes, 4) if step_bytes else None,
         "note": "dominant kernel by bytes; timed alone (single stream, eager) with CUDA events on its stream; "
                 "achieved = algorithmic bytes / duration"}
    hb = sum(k["alg_bytes"] * k["launches_per_step"] for k in hbm)
    ht = sum(k["ms_per_launch"] * k["launches_per_step"] for k in hbm)
    r["hbm_bound_kernels_all"] = {"alg_bytes_per_step": hb, "ms_per_step": round(ht, 4),
                                  "gbs": round(hb / (ht * 1e-3) / 1e9, 1) if ht else None,
                                  "frac": round(hb / (ht * 1e-3) / 1e9 / peak, 4) if ht else None,
                                  "min_frac_over_group_rows": min([k["frac_hbm"] for k in hbm if k["entry"].startswith("de6d_group")] or [None])}
    issue = ncu_issue(by_time["entry"], by_time["shape"]) or {}
    r.update({"kernel_by_time": by_time["entry"], "shape_by_time": by_time["shape"], "bound_by_time": "issue",
              "ms_per_launch_by_time": by_time["ms_per_launch"],
              "share_of_step_time": round(by_time["ms_per_launch"] * by_time["launches_per_step"] / total_ms, 4) if total_ms else None,
              "issue_active": issue.get("issue_active"), "sms_occupied": issue.get("sms_occupied"),
              "frac_by_time": (round(issue["issue_active"] * issue["sms_occupied"] / 148.0, 4)
                               if issue.get("issue_active") is not None and issue.get("sms_occupied") else None),
              "cycles_per_sample": issue.get("cycles_per_sample"), "smem_wavefront_frac": issue.get("smem_wavefront_frac"),
              "stalls": issue.get("stalls"), "issue_source": issue.get("source"),
              "hbm_frac_by_time": by_time["frac_hbm"], "traffic_by_time": ncu_traffic(by_time["entry"], by_time["shape"]),
              "note_by_time": "dominant kernel by time: latency/issue-bound (one CTA or cluster per cloud iterating on chip); "
                              "frac_by_time = issue-active fraction x SMs occupied / 148 from the committed ncu capture"})
    r["whole_step"] = {"alg_bytes": step_bytes, "ms_per_step_pipelined": round(step_ms, 4),
                       "gbs": round(step_bytes / (step_ms * 1e-3) / 1e9, 1), "frac_hbm": round(step_bytes / (step_ms * 1e-3) / 1e9 / peak, 4)}
    return r


def reference_cuda_leg(cfg, batch, dev, chain):
    """The yardstick SURVEY.md names: the reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100)
    driven by the reference's own python, op by op, on this GPU and these inputs.  Not part of any timed product region."""
    try:
        import warnings
        warnings.filterwarnings("ignore")
        from oracle import build_ref, ref_py, chain_ref_cuda
        if not (build_ref.available() and ref_py.available()):
            return {"unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
        import torch
        tree = ref_py.load_tree("pcdet_ref", build_ref.load())
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.cuda.device(dev):
            ms, ops = chain_ref_cuda.time_chain(cfg, chain.inputs, tree, steps=2, warmup=1)
            ref, _ = chain_ref_cuda.run(cfg, chain.inputs, tree)
            torch.cuda.synchronize()
            same = {}
            for k in ("l0_idx", "l1_idx", "l2_idx"):
                if k in ref and k in chain.outputs:
                    same[k] = round(float((ref[k] == chain.outputs[k]).float().mean()), 5)
        return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": round(ms, 3),
                "per_op_ms": {k: round(v, 3) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])},
                "kind": "reference CUDA kernels recompiled for sm_100 (nvcc -O2), reference python wrappers, legacy default stream, "
                        "host-synchronised wall clock over 2 passes",
                "same_index_fraction_vs_product_chain": same}
    except Exception as e:  # noqa: BLE001 -- a yardstick that cannot run must not take the bench line down
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def sa_mlp_leg(batch, dev):
    """SA layers WITH their shared MLPs -- the reference composition (ball query -> grouped tensor in HBM
    -> cuDNN 1x1 convs + BN + ReLU -> mask -> max-pool; TF32 convolutions, torch's and the reference's default) against the
    fused kernel (ball query -> csrc/sa_mlp.cu: gather + tcgen05 tf32 MLP in tensor memory + mask + max-pool).  Layer shapes:
    SASA / 3DSSD SA1, SA2, SA3 (SA3's first two scales run as cta_group::2 pairs, its third -- 131->128->256->256, 464 KB of
    tf32 weights -- keeps the composition); random weights, eval-mode BatchNorm."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from de6d_b200 import pointnet2_utils as pu, sa_fused, synth
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf32_peak = float(json.load(f)["bf16_tflops"]) / 2.0
        peak_src = "half the measured dense bf16 peak (MEASURED_PEAKS.json): tf32 runs at half the bf16 rate"
    except Exception:
        tf32_peak, peak_src = 1590.0 / 2.0, "half the fallback bf16 peak (B200_PROFILING.md)"
    layers = {
        "SA1 16384->4096, C=1": (16384, 4096, 1, [(0.2, 32, [16, 16, 32]), (0.4, 32, [16, 16, 32]), (0.8, 64, [32, 32, 64])]),
        "SA2 4096->1024, C=64": (4096, 1024, 64, [(0.4, 32, [64, 64, 128]), (0.8, 32, [64, 64, 128]), (1.6, 64, [64, 96, 128])]),
        "SA3 1024->512, C=128": (1024, 512, 128, [(1.6, 32, [128, 128, 256]), (3.2, 32, [128, 192, 256]), (4.8, 32, [128, 256, 256])]),
    }
    out = {}
    with torch.cuda.device(dev), torch.no_grad():
        for name, (n, m, c, scales) in layers.items():
            xyz = torch.from_numpy(synth.clouds(batch, n, seed=1)).to(dev)
            new_xyz = xyz[:, :m].contiguous()
            feats = torch.randn(batch, c, n, device=dev)
            mods, fused, flops, grouped_bytes = [], [], 0.0, 0
            for r, ns, mlp in scales:
                widths = [c + 3] + mlp
                seq = []
                for a, b in zip(widths[:-1], widths[1:]):
                    seq += [nn.Conv2d(a, b, 1, bias=False), nn.BatchNorm2d(b), nn.ReLU()]
                    flops += 2.0 * a * b * batch * m * ns
                seq = nn.Sequential(*seq).to(dev).eval()
                mods.append((r, ns, seq))
                fused.append(sa_fused.FusedSAScale(r, ns, seq) if sa_fused.FusedSAScale.supported(seq, ns) else None)
                grouped_bytes += 4 * (c + 3) * batch * m * ns * (fused[-1] is not None)
            grid = pu.BallQueryGrid(xyz, min(r for r, _, _ in scales)) if pu.BallQueryGrid.wanted(n) else None
            _, xyz_t = pu.gather_xyz(xyz, None)

            def one_unfused(r, ns, seq):
                cnt, idx = pu.ball_query_cnt(r, ns, xyz, new_xyz, grid=grid)
                g = pu.group_concat(xyz, new_xyz, feats, idx, xyz_t=xyz_t)
                y = seq(g) * (cnt > 0).float().unsqueeze(1).unsqueeze(-1)
                return F.max_pool2d(y, kernel_size=[1, ns]).squeeze(-1)

            def unfused():
                return [one_unfused(r, ns, seq) for r, ns, seq in mods]

            def fused_run():       # scales the kernel cannot take (weights beyond two SMs' shared memory) keep the composition
                fpm = feats.transpose(1, 2).contiguous()
                return [fs(xyz, new_xyz, feats, grid=grid, feats_pm=fpm) if fs is not None else one_unfused(*md)
                        for fs, md in zip(fused, mods)]

            def timeit(fn, reps=5):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize(dev)
                return e0.elapsed_time(e1) / reps
            a, b2 = unfused(), fused_run()
            err = max(float((x - y).abs().max() / x.abs().max()) for x, y in zip(a, b2))
            del a, b2
            t_un, t_fu = timeit(unfused), timeit(fused_run)
            out[name] = {"unfused_ms": round(t_un, 3), "fused_ms": round(t_fu, 3), "speedup": round(t_un / t_fu, 2),
                         "mlp_gflop": round(flops / 1e9, 1), "fused_tflops_incl_ball_query": round(flops / (t_fu * 1e-3) / 1e12, 1),
                         "grouped_tensor_bytes_not_written": grouped_bytes, "max_rel_diff_vs_cudnn_tf32": round(err, 5),
                         "scales_fused": [(("pair" if not sa_fused.FusedSAScale.single_cta(fs) else "yes") +
                                           (" x%d launches over row blocks of the last layer" % fs.parts if fs.parts > 1 else ""))
                                          if fs is not None else "no (weights exceed two SMs' shared memory)" for fs in fused]}
            torch.cuda.empty_cache()
    out["tensor_peak_tflops"] = tf32_peak
    out["tensor_peak_source"] = peak_src
    out["note"] = ("batch %d; unfused = ball_query_cnt (shared grid) + de6d_group_concat_t + torch Conv2d/BN/ReLU (cuDNN, allow_tf32) + mask + "
                   "max_pool2d; fused = ball_query_cnt + de6d_sa_mlp_fused (+ one feature transposition per layer)" % batch)
    return out


def backbone_leg(batch, dev):
    """The caller one level above the op chain: the reference's own PointNet2FSMSG backbone (pointnet2_backbone.py:97-263: the
    three SA layers WITH their shared / aggregation / confidence MLPs) on `batch` 16384-point frames, eval mode:
      reference_kernels   unmodified reference python over the reference's CUDA extension (oracle/_ref)
      compat              the same unmodified python over de6d_b200.compat (drop-in: bit-identical outputs)
      fused               de6d_b200.sa_fused.fuse_backbone (input staging kernel, shared grids, fused SA scales on tcgen05)
    cuDNN / cuBLAS TF32 allowed as by torch default.  Outside every timed region of the op-chain arm."""
    import copy
    import warnings
    import numpy as np
    import torch
    warnings.filterwarnings("ignore")
    from de6d_b200 import compat, sa_fused, synth
    from oracle import build_ref, ref_py
    if not (build_ref.available() and ref_py.available()):
        return {"unavailable": "oracle/_ref not built"}
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    ours, theirs = ref_py.load_pair()
    xyz = synth.clouds(batch, 16384, seed=2)
    inten = np.random.default_rng(2).random((batch, 16384, 1), dtype=np.float32)
    pts = np.concatenate([np.repeat(np.arange(batch, dtype=np.float32), 16384)[:, None], np.concatenate([xyz, inten], -1).reshape(-1, 4)], 1)
    out = {}
    with torch.cuda.device(dev), torch.no_grad():
        points = torch.from_numpy(pts).to(dev)
        mods = {}
        for name, tree in (("reference_kernels", theirs), ("compat", ours)):
            torch.manual_seed(0)
            mods[name] = tree.pointnet2_backbone.PointNet2FSMSG(copy.deepcopy(synth.sasa_backbone_cfg()), input_channels=4).to(dev).eval()
        fused = sa_fused.fuse_backbone(mods["compat"], ffps="fused")
        runs = {"reference_kernels": lambda: mods["reference_kernels"]({"batch_size": batch, "points": points}),
                "compat": lambda: mods["compat"]({"batch_size": batch, "points": points}),
                "fused": lambda: fused({"batch_size": batch, "points": points})}
        for name, fn in runs.items():
            fn()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            reps = 2 if name == "reference_kernels" else 5
            for _ in range(reps):
                r = fn()
            torch.cuda.synchronize(dev)
            ms = 1e3 * (time.perf_counter() - t0) / reps
            out[name] = {"ms_per_batch": round(ms, 2), "frames_per_s": round(batch / (ms * 1e-3), 1)}
        a, b = runs["compat"](), runs["reference_kernels"]()
        out["compat_bit_identical_to_reference"] = bool(torch.equal(a["point_features"], b["point_features"]) and torch.equal(a["point_coords"], b["point_coords"]))
    out["note"] = ("batch %d, host-synchronised wall clock; SASA / 3DSSD PointNet2FSMSG 16384 -> 4096 -> 1024 -> 512 with d-fps / f-fps+d-fps / "
                   "s-fps+d-fps sampling, three radius scales per layer; random weights, eval mode" % batch)
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from de6d_b200 import build as de6d_build
    from de6d_b200 import dist as ddist

    rank, world, local = ddist.init_from_env()
    if rank == 0:
        de6d_build.build()           # no-op when lib/libde6d_b200.so is newer than csrc/ (it travels with the repo)
    if world > 1:
        dist.barrier()               # the other ranks only load the library
    from de6d_b200 import chain as ch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the op chain has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg, batch = config_of(args)
    P = max(1, args.pipeline)
    run = ChainRunner(cfg, batch, dev, P, rank, world)
    chain = run.chains[0]
    out_bytes = sum(v.numel() * v.element_size() for k, v in chain.outputs.items() if isinstance(v, torch.Tensor))
    clocks = ClockSampler(local)

    # ---- device-resident timed region, then end to end (H2D of the sensor inputs + chain + D2H), every step ---------
    W = max(args.warmup, 3)
    run.timed(lambda k: run.chains[k % P].step(), W)
    clocks.start()
    ms = ddist.max_over_ranks(run.timed(lambda k: run.chains[k % P].step(), args.steps), dev)
    ms_e2e = run.e2e(args.steps, ch.OpChain.SENSOR_KEYS)
    clocks.stop()
    ms_e2e_all = run.e2e(args.steps, None)
    frames_global = batch * world
    # the one collective of the deployment: all ranks' padded keep lists gathered over NCCL (outside the timed regions)
    gather_ms = None
    if world > 1 and cfg.n_proposals > 0:
        keep, num = chain.outputs["nms_keep"], chain.outputs["nms_num"]
        ddist.gather_detections(keep, num)
        run.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        keep_all, num_all = ddist.gather_detections(keep, num)
        g1.record()
        torch.cuda.synchronize()
        assert keep_all.shape[0] == world * batch and num_all.shape[0] == world * batch
        gather_ms = g0.elapsed_time(g1)

    line = None
    if rank == 0:
        kernels, total_ms = per_kernel_pass(cfg, batch, dev, run.hosts[0])
        roofline = build_roofline(kernels, total_ms, ms / args.steps)
        fps_us = None
        for k in kernels:
            if k["entry"] == "de6d_furthest_point_sampling" and k["shape"][1] == cfg.n_points:
                fps_us = 1e3 * k["ms_per_launch"] / batch
        line = {
            "metric": METRIC, "value": frames_global * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, batch, world,
                                      note="L2: per-step working set %.0f MB of inputs+outputs >> 126 MB L2, no explicit flush; "
                                           "%d chains in flight (round-robin), every step a full pass over its own batch. "
                                           "F-FPS: the fused kernel evaluates the metric by direct differences, the reference "
                                           "pipeline by torch.cdist's fp32 GEMM expansion -- same selected sets in >= 99.8 %%, same "
                                           "position in >= 98.9 %% of the slots, and closer to the float64-exact greedy sequence than "
                                           "the reference itself (profiles/r2_ffps_agreement.json); `value_ffps_reference_route` is "
                                           "the chain with torch.cdist + matrix kernel (bit-identical to the reference)" % (
                                          (out_bytes + chain.h2d_bytes()) / 1e6, P)),
            "e2e": {"value": frames_global * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": chain.h2d_bytes(ch.OpChain.SENSOR_KEYS), "d2h_bytes_per_step": chain.d2h_bytes(),
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "every step: H2D of the sensor inputs a deployment receives per frame (points + intensity) from pinned "
                            "host memory, the chain, D2H of sampled indices + kept detections.  The per-layer features, S-FPS "
                            "scores, vote features and proposals are outputs of the detector's MLPs / head ON the device "
                            "(pointnet2_backbone.py:199-263); their synthetic stand-ins stay resident.  e2e_all_inputs copies them too"},
            "e2e_all_inputs": {"value": frames_global * args.steps / (ms_e2e_all * 1e-3), "unit": UNIT,
                               "h2d_bytes_per_step": chain.h2d_bytes(), "d2h_bytes_per_step": chain.d2h_bytes(),
                               "ms_per_step": ms_e2e_all / args.steps,
                               "note": "also copies the stand-ins for on-device MLP outputs every step (PCIe-bound)"},
            "gpu_launches": int(chain.kernels_per_step) * args.steps,
            "gpu_launches_note": "%d launches of this library's kernels per step (counted by de6d_launch_count on the eager "
                                 "warm-up pass) replayed from one CUDA graph per step" % chain.kernels_per_step,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "fps_us_per_frame": fps_us,
            "kernels": kernels,
            "serialized_kernel_ms_per_step": round(total_ms, 4),
            "detections_all_gather_ms": gather_ms,
        }
    # ---- extras (rank 0 of a single-GPU run; each outside the timed regions above) ------------------------------------
    if rank == 0 and world == 1 and not args.no_extras:
        extras_t0 = time.perf_counter()
        has_ffps = any("f-fps" in l.methods for l in cfg.layers)
        if has_ffps:       # the chain with the reference-identical F-FPS route: the cost of the deviation made visible
            del run.chains[1:]
            alt = ChainRunner(cfg, batch, dev, min(P, 3), rank, world, ffps="cdist")
            ms_alt = alt.resident(10, 3)
            line["value_ffps_reference_route"] = {"value": batch * 10 / (ms_alt * 1e-3), "unit": UNIT, "ms_per_step": ms_alt / 10,
                                                  "note": "F-FPS as torch.cdist (cuBLAS) + de6d_furthest_point_sampling_matrix: "
                                                          "bit-identical indices to the reference pipeline on this GPU"}
            del alt
            torch.cuda.empty_cache()
        line["reference_cuda"] = reference_cuda_leg(cfg, batch, dev, chain)
        try:
            line["backbone"] = backbone_leg(16, dev)
        except Exception as e:  # noqa: BLE001
            line["backbone"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
        try:
            line["sa_mlp_fused"] = sa_mlp_leg(batch, dev)
        except Exception as e:  # noqa: BLE001
            line["sa_mlp_fused"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        others = {}
        for name in ("chain16", "stress131072"):
            if name == args.config:
                continue
            a2 = argparse.Namespace(**vars(args))
            a2.config, a2.batch = name, None
            cfg2, b2 = config_of(a2)
            try:
                p2 = 8 if name == "chain16" else 4      # enough independent chains in flight to fill the SMs the sampling kernels leave idle
                # several chains in flight: the SM-time per cloud, not one launch's latency, is what counts -> 4-CTA F-FPS clusters
                r2 = ChainRunner(cfg2, b2, dev, p2, rank, world, ffps_cluster=44)
                ms2 = r2.resident(10, 3)
                ms2_e2e = r2.e2e(10, ch.OpChain.SENSOR_KEYS)
                r1 = ChainRunner(cfg2, b2, dev, 1, rank, world)
                ms1 = r1.resident(10, 3)
                k2, t2 = per_kernel_pass(cfg2, b2, dev, r2.hosts[0], reps=3)
                others[name] = {"workload": workload_config(a2, b2, 1)["workload"], "frames_per_step": b2,
                                "value": b2 * 10 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / 10, "chains_in_flight": p2,
                                "e2e_value": b2 * 10 / (ms2_e2e * 1e-3),
                                "serial_value": b2 * 10 / (ms1 * 1e-3), "serial_ms_per_step": ms1 / 10,
                                "kernels": [{kk: k[kk] for kk in ("entry", "shape", "ms_per_launch", "launches_per_step", "gbs", "frac_hbm")} for k in k2[:8]]}
                del r2, r1
            except Exception as e:  # noqa: BLE001
                others[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
            torch.cuda.empty_cache()
        line["other_configs"] = others
        line["extras_seconds"] = round(time.perf_counter() - extras_t0, 1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single_core(args.cpu_seconds)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        args.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Route everything that libraries print on fd 1 (e.g. NCCL's version banner) to stderr; returns a writer for
    the one JSON line that must be the only thing on stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(saved, "w")

    def emit(line):
        out.write(line + "\n")
        out.flush()
    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="chain64", choices=sorted(CONFIGS), help="BASELINE.json workload (default: configs[2])")
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step (default: the config's)")
    ap.add_argument("--no-extras", action="store_true", help="skip reference_cuda / other_configs / reference-route legs")
    ap.add_argument("--pipeline", type=int, default=8, help="independent chains in flight (1 = strictly serial steps)")
    ap.add_argument("--impl", default="de6d_b200", choices=["de6d_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.emit = _claim_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())

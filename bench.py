#!/usr/bin/env python
"""bench.py -- pillar-encoder throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode B|A]

Workload (BASELINE.json configs[2]/[3]): paired radar + LiDAR pillar encoding, forward + backward,
8 frames per GPU with the radar_distill_train.yaml grid; synthetic nuScenes-shaped clouds
(radardistill_b200/synth.py), random-init weights.  A "step" is one pass of both encoders over the batch.
  mode B (default, the headline "fwd+bwd"): both encoders train-mode BN, forward + parameter backward
          (teacher pre-training, tools/cfgs/nuscenes_models/pillarnet.yaml);
  mode A (reference-faithful distillation step, radar_distill_train.yaml:68): LiDAR encoder frozen
          (eval BN, forward only), radar encoder train-mode forward + backward.  Reported under "mode_a".
value  = points/s over all ranks with inputs resident in HBM (CUDA events per step, L2 flushed between steps); the
         backward is driven by a dense upstream gradient (what SparseEnc hands back), resident like the inputs
e2e    = the same through the module API from pinned HOST buffers: every step uploads its points (H2D) and reads the
         step's result -- the parameter gradients -- back (D2H); "e2e_host_outputs" additionally downloads the
         pillar features / coords (what a host-side consumer of rdp_encode_host would receive)
--impl reference: the CPU oracle port (oracle/pillar_oracle.c, all host threads) on the same workload.
Multi-GPU: frames shard over ranks (weak scaling, 8 frames per GPU); the only collective is DDP's gradient
all-reduce of the PFN parameters.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_GPU = 8
S2D_CFG = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])


class Cfg(dict):
    __getattr__ = dict.__getitem__


def make_clouds(rank: int, frames: int):
    from radardistill_b200 import synth
    lidar = synth.collate([synth.lidar_frame(rank * frames + b) for b in range(frames)])
    radar = synth.collate([synth.radar_frame(rank * frames + b) for b in range(frames)])
    return lidar, radar


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_pair(seed=0):
    from oracle import oracle as orc
    from radardistill_b200 import synth
    rng = np.random.default_rng(seed)
    out = {}
    for kind, c in (("lidar", 5), ("radar", 6)):
        cfg = orc.OracleConfig(num_point_features=c, voxel_size=tuple(synth.VOXEL_SIZE), grid_size=tuple(synth.grid_size_of()),
                               point_cloud_range=tuple(synth.PC_RANGE))
        w = (rng.standard_normal((32, cfg.c_in)) * 0.2).astype(np.float32)
        out[kind] = orc.PillarOracle(cfg, w, rng.uniform(0.5, 1.5, 32), rng.normal(0, 0.2, 32), rng.normal(0, 1, 32),
                                     rng.uniform(0.5, 4, 32))
    return out


def cpu_step(oracles, lidar, radar, mode):
    """One step of the workload on the CPU oracle (same work as the GPU arm)."""
    for kind, pts in (("lidar", lidar), ("radar", radar)):
        o = oracles[kind]
        train = (mode == "B") or kind == "radar"
        r = o.forward(pts, training=train, keep_intermediates=train)
        if train:
            g = np.ones_like(r["features"])
            o.backward(r, g)


def time_cpu(lidar, radar, mode, steps, warmup, threads):
    from oracle import oracle as orc
    orc.set_threads(threads)
    oracles = oracle_pair()
    for _ in range(warmup):
        cpu_step(oracles, lidar, radar, mode)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(oracles, lidar, radar, mode)
    dt = (time.perf_counter() - t0) / steps
    return (len(lidar) + len(radar)) / dt, dt


def run_reference(args):
    """--impl reference: the CPU port of the reference path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = 2  # bounded sample: 2 of the 8 paired frames per step (same generators, same per-frame shapes)
    lidar, radar = make_clouds(0, frames)
    value, dt = time_cpu(lidar, radar, args.mode, max(args.steps, 1), min(args.warmup, 1), cores)
    line = {"impl": "reference", "metric": "pillar-encoder points/s (fwd+bwd)", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.mode, frames, len(lidar), len(radar)),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": f"{frames} paired frames/step ({len(lidar)} LiDAR + {len(radar)} radar rows), C oracle, "
                                       f"{cores} threads in the per-point loops"},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(mode, frames, n_lidar, n_radar):
    return {"workload": "paired radar+LiDAR pillar encoding fwd+bwd (BASELINE.json configs[2]; radar_distill_train.yaml grid 1440x1440, "
                        "0.075 m pillars)", "mode": "B: both encoders train-BN fwd+bwd" if mode == "B" else
            "A: LiDAR frozen eval fwd, radar train fwd+bwd", "frames_per_gpu": frames, "lidar_rows_per_gpu": int(n_lidar),
            "radar_rows_per_gpu": int(n_radar), "encoders": "DynamicPillarVFESimple2D 14->32 + Radar_DynamicPillarVFESimple2D 15->32",
            "l2": "256 MiB scratch written between timed steps (L2 flushed)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def build_modules(device, mode, ddp):
    import torch
    from radardistill_b200 import synth, vfe
    from radardistill_b200.vfe import forward_pair
    torch.manual_seed(1234)
    lid = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D_CFG), num_point_features=5, voxel_size=synth.VOXEL_SIZE,
                                       grid_size=synth.grid_size_of(), point_cloud_range=synth.PC_RANGE).to(device)
    rad = vfe.Radar_DynamicPillarVFESimple2D(model_cfg=Cfg(S2D_CFG), num_point_features=6, voxel_size=synth.VOXEL_SIZE,
                                             grid_size=synth.grid_size_of(), point_cloud_range=synth.PC_RANGE).to(device)
    for m in (lid, rad):
        n = m.pfn_layers[0].norm
        with torch.no_grad():
            n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
    if mode == "A":
        lid.eval()
        for p in lid.parameters():
            p.requires_grad_(False)
    else:
        lid.train()
    rad.train()
    class PairedEncoders(torch.nn.Module):
        """`vfe` then `radar_vfe`, as PillarNet's module list runs them (pillarnet.py:28-33); frozen modules run under
        no_grad in eval mode like FREEZE_PIPELINE does (pillarnet.py:17-25,31-32)."""

        def __init__(self, vfe_mod, radar_vfe_mod, lidar_no_grad):
            super().__init__()
            self.vfe, self.radar_vfe, self.lidar_no_grad = vfe_mod, radar_vfe_mod, lidar_no_grad

        def forward(self, batch_dict):
            # the two encoders are independent: radar on a side stream, LiDAR on the current one
            return forward_pair(self.vfe, self.radar_vfe, batch_dict, first_no_grad=self.lidar_no_grad)

    pair = PairedEncoders(lid, rad, lidar_no_grad=(mode == "A"))
    call = pair
    if ddp == "ddp":
        from torch.nn.parallel import DistributedDataParallel as DDP
        call = DDP(pair, device_ids=[device.index], gradient_as_bucket_view=True)   # one reducer / one bucket for all PFN parameters
    return lid, rad, call




def gpu_step(call, lidar_dev, radar_dev, mode, frames, upstream):
    """One step: both encoders over the batch as PillarNet.forward runs them (pillarnet.py:28-33), then the backward
    from a dense upstream gradient d(loss)/d(pillar_features) -- what the SparseEnc backbone's backward hands to the
    encoder (spconv_backbone_2d.py:262) -- into the PFN parameters."""
    import torch
    for p in upstream["params"]:   # optimizer.zero_grad(set_to_none=True) of the training loop (train_utils.py:55)
        p.grad = None
    bd = call({"points": lidar_dev, "radar_points": radar_dev, "batch_size": frames})
    outs = [bd["radar_pillar_features"]]
    grads = [upstream["radar"][:outs[0].shape[0]]]
    if mode == "B":
        outs.append(bd["pillar_features"])
        grads.append(upstream["lidar"][:outs[1].shape[0]])
    torch.autograd.backward(outs, grads)
    if upstream.get("reducer") is not None:   # the step's one collective: gradient all-reduce of the PFN parameters
        upstream["reducer"].reduce()
    return bd


def make_upstream(device, n_lidar, n_radar, params=(), c_out=32):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(99)
    return {"lidar": torch.randn((n_lidar, c_out), device=device, generator=g),
            "radar": torch.randn((n_radar, c_out), device=device, generator=g), "params": list(params)}


def grads_vector(mods):
    """The step's result: every PFN parameter gradient, flattened (1 056 floats for the two shipped encoders)."""
    import torch
    return torch.cat([p.grad.reshape(-1) for m in mods for p in m.parameters() if p.grad is not None])


def pin_cpus(local: int, local_world: int):
    """Gives every rank its own physical cores (all hyperthreads of each): the ranks' launch threads then never share a core.
    Returns the CPU set, or None if the topology cannot be read."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        cores = {}
        for c in allowed:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                sib = f.read().strip()
            cores.setdefault(sib, []).append(c)
        groups = sorted(cores.values(), key=lambda g: g[0])
        per = len(groups) // local_world
        if per < 1:
            return None
        mine = sorted(c for g in groups[local * per:(local + 1) * per] for c in g)
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pinned = None
    if world > 1 and args.pin_cpus:   # before torch spawns its threads
        pinned = pin_cpus(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    import torch
    import torch.distributed as dist
    from radardistill_b200 import _lib, ops
    ddp = world > 1
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if ddp:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()
    mode, frames = args.mode, FRAMES_PER_GPU
    lidar, radar = make_clouds(rank, frames)
    n_rows = len(lidar) + len(radar)
    lid, rad, call = build_modules(device, mode, args.dp if ddp else False)
    lidar_dev, radar_dev = torch.from_numpy(lidar).to(device), torch.from_numpy(radar).to(device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    upstream = make_upstream(device, len(lidar), len(radar), list(lid.parameters()) + list(rad.parameters()))
    if ddp and args.dp == "lean":
        from radardistill_b200.sharding import GradientAllReduce
        upstream["reducer"] = GradientAllReduce(upstream["params"])

    def barrier():
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps(m, k):
        evs = []
        # the step is within ~10 % of host bound and every step ends in a rendezvous of all ranks: a cyclic-GC pause on any
        # one rank stalls all of them, so the collector is parked over the timed steps (as long-running training loops do:
        # collect at a step boundary of your choosing, not in the middle of a launch sequence)
        gc.collect()
        gc.freeze()
        gc.disable()
        try:
            return _timed_steps(m, k, evs)
        finally:
            gc.enable()
            gc.unfreeze()

    def _timed_steps(m, k, evs):
        for _ in range(k):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            gpu_step(call, lidar_dev, radar_dev, m, frames, upstream)
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in evs]

    clk = ClockSampler(local)
    if rank == 0:     # one sampler per job: the step is host-bound, eight pollers would perturb what they measure
        clk.__enter__()   # started before the warm-up so that nvidia-smi is already streaming when the timed steps run
    for _ in range(max(args.warmup, 3)):
        gpu_step(call, lidar_dev, radar_dev, mode, frames, upstream)
    barrier()
    clk.rows.clear()  # keep only samples taken during the timed region
    t_wall0 = time.perf_counter()
    ms = timed_steps(mode, args.steps)
    barrier()
    wall = time.perf_counter() - t_wall0
    step_ms = sum(ms) / len(ms)
    if ddp:
        t = torch.tensor([step_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms = float(t.item())
        tot = torch.tensor([float(n_rows)], device=device, dtype=torch.float64)
        dist.all_reduce(tot)
        total_rows = float(tot.item())
    else:
        total_rows = float(n_rows)
    value = total_rows / (step_ms * 1e-3)
    if wall < 0.3:  # short timed region: keep the same load running (same count on every rank) until nvidia-smi has samples
        for _ in range(int(0.3 / (step_ms * 1e-3)) + 1):
            gpu_step(call, lidar_dev, radar_dev, mode, frames, upstream)
        barrier()
    if rank == 0:
        clk.__exit__(None, None, None)

    # ---- mode A beside it (reference-faithful step) when the headline is mode B, N == 1 only
    extra = {}
    if mode == "B" and not ddp:
        lidA, radA, callA = build_modules(device, "A", False)
        upstreamA = dict(upstream, params=list(lidA.parameters()) + list(radA.parameters()))
        for _ in range(3):
            gpu_step(callA, lidar_dev, radar_dev, "A", frames, upstreamA)
        evs = []
        for _ in range(args.steps):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); gpu_step(callA, lidar_dev, radar_dev, "A", frames, upstreamA); e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        msA = sum(s.elapsed_time(e) for s, e in evs) / len(evs)
        extra["mode_a"] = {"value": n_rows / (msA * 1e-3), "unit": "points/s", "ms_per_step": msA,
                           "what": "LiDAR frozen eval-BN forward + radar train-BN forward+backward (radar_distill_train.yaml)"}

    # ---- e2e: pinned host buffers in, the step's result out, through the package's host pipeline
    #      (radardistill_b200.pipeline.HostPipeline: uploads / downloads overlap the kernels; every step's H2D of its
    #      points and D2H of its result complete inside the timed region).  Result of a training step = the parameter
    #      gradients; the second pass also downloads the pillar features + coords.
    from radardistill_b200.pipeline import HostPipeline
    # host side of the e2e arm: the frames as the dataset yields them -- back to back, no batch column -- plus int32 frame
    # offsets; the batch index is attached on the device (rdp_index_fwd_frames) instead of by collate_batch's np.pad
    lidar_raw, radar_raw = lidar[:, 1:], radar[:, 1:]
    def offsets_of(rows):
        return np.concatenate([[0], np.cumsum(np.bincount(rows[:, 0].astype(np.int64), minlength=frames))]).astype(np.int32)
    host_in = {"points": torch.from_numpy(np.ascontiguousarray(lidar_raw)).pin_memory(),
               "radar_points": torch.from_numpy(np.ascontiguousarray(radar_raw)).pin_memory(),
               "points_offsets": torch.from_numpy(offsets_of(lidar)).pin_memory(),
               "radar_points_offsets": torch.from_numpy(offsets_of(radar)).pin_memory()}
    h2d_bytes = sum(int(v.numel() * v.element_size()) for v in host_in.values())
    mods = (lid, rad)

    def e2e_step(d):
        for p in upstream["params"]:
            p.grad = None
        bd = call(dict(d))
        outs = [bd["radar_pillar_features"]]
        grads = [upstream["radar"][:outs[0].shape[0]]]
        if mode == "B":
            outs.append(bd["pillar_features"])
            grads.append(upstream["lidar"][:outs[1].shape[0]])
        torch.autograd.backward(outs, grads)
        if upstream.get("reducer") is not None:
            upstream["reducer"].reduce()
        bd["param_grads"] = grads_vector(mods).unsqueeze(1)
        return bd

    def time_e2e(keys):
        pipe = HostPipeline(e2e_step, device, keys)
        for _ in range(3):
            d2h = pipe.submit(host_in, host_in)
        pipe.finish()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            d2h = pipe.submit(host_in, host_in if i + 1 < args.steps else None)
        pipe.finish()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        if ddp:
            t = torch.tensor([dt], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, int(d2h)

    e2e_dt, d2h = time_e2e(("param_grads",))
    e2e = {"value": total_rows / e2e_dt, "unit": "points/s", "h2d_bytes_per_step": h2d_bytes,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_dt * 1e3,
           "how": "HostPipeline: pinned H2D of the frames (no batch column; int32 frame offsets instead) on an input stream, D2H of "
                  "the step's parameter gradients on an output stream, overlapped with the kernels of the neighbouring steps"}
    out_dt, out_d2h = time_e2e(("param_grads", "pillar_features", "pillar_coords", "radar_pillar_features", "radar_pillar_coords"))
    e2e_out = {"value": total_rows / out_dt, "unit": "points/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": out_d2h, "ms_per_step": out_dt * 1e3,
               "how": "as e2e, plus the D2H of the pillar features and coords of both encoders (PCIe bound)"}

    # ---- roofline: the PFN kernels on the LiDAR batch, each timed live with CUDA events on the launching stream, L2
    #      flushed before every launch.  Headline = pfn_tile_kernel<APPLY> (one launch per eval-mode rdp_pfn_fwd: the
    #      kernel both step modes spend the most forward time in); the train-mode forward and the backward calls
    #      (several launches each, dominated by one tile kernel) are listed beside it.
    roof = None
    if rank == 0:
        import ctypes as C
        spec = lid.spec
        norm = lid.pfn_layers[0].norm
        lib = _lib.load()
        geom, layout = spec.geom(frames), spec.layout_struct()
        w = lid.pfn_layers[0].linear.weight.detach()
        res = ops.encode_forward(lidar_dev, spec, frames, w, None, norm.weight, norm.bias, norm.running_mean.clone(),
                                 norm.running_var.clone(), True, True)   # train-mode state for the backward timing
        torch.cuda.synchronize()
        P_ = ops._ptr
        feats = torch.empty((len(lidar), spec.c_out), dtype=torch.float32, device=device)
        argp = torch.empty((len(lidar), spec.c_out), dtype=torch.int32, device=device)
        dw, dg, db = (torch.empty(sh, device=device) for sh in ((spec.c_out, spec.c_in), (spec.c_out,), (spec.c_out,)))
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rm, rv = norm.running_mean.clone(), norm.running_var.clone()
        prm_eval = ops._params_struct(spec, w, None, norm.weight.detach(), norm.bias.detach(), rm, rv, False)
        prm_train = ops._params_struct(spec, w, None, norm.weight.detach(), norm.bias.detach(), rm, rv, True)

        def call_index():
            _lib.check(lib.rdp_index_fwd(P_(lidar_dev), len(lidar), C.byref(geom), spec.coord_cols, P_(res.workspace),
                                         res.workspace.numel(), P_(res.coords), P_(res.inverse), P_(res.counts),
                                         P_(res.counters), st), "rdp_index_fwd")

        def call_fwd(prm, arg, state):
            _lib.check(lib.rdp_pfn_fwd(P_(lidar_dev), len(lidar), C.byref(geom), C.byref(layout), C.byref(prm), P_(res.workspace),
                                       res.workspace.numel(), P_(res.counters), P_(feats), arg, None, state, st), "rdp_pfn_fwd")

        def call_bwd():
            _lib.check(lib.rdp_pfn_bwd(P_(lidar_dev), len(lidar), C.byref(geom), C.byref(layout), C.byref(prm_train),
                                       P_(res.workspace), res.workspace.numel(), P_(res.counters), P_(upstream["lidar"]),
                                       P_(feats), P_(argp), P_(res.bn_state), P_(dw), P_(dg), P_(db), st), "rdp_pfn_bwd")

        def timed(fn, reps):
            ts = []
            for i in range(reps + 3):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            return sum(ts) / len(ts)

        reps = max(args.steps, 10)
        idx = timed(call_index, reps)
        dur = timed(lambda: call_fwd(prm_eval, None, None), reps)
        dur_train = timed(lambda: call_fwd(prm_train, P_(argp), P_(res.bn_state)), reps)
        dur_bwd = timed(call_bwd, reps)
        n_kept, n_pil = res.n_kept, res.n_pillars
        row_bytes = 4 * spec.cols
        alg = row_bytes * n_kept + 4 * spec.c_out * n_pil           # rows read once + feature rows written once
        alg_fwd = row_bytes * len(lidar) + 4 * n_kept + n_pil * (4 * spec.c_out + 4 * spec.coord_cols + 4)  # SURVEY 8(d) B_fwd
        alg_train = 2 * row_bytes * n_kept + 8 * spec.c_out * n_pil  # moments pass re-reads the rows; features + argmax written
        alg_bwd = 8 * spec.c_out * n_pil + row_bytes * n_kept + 4 * n_kept   # SURVEY 8(d) B_bwd
        peak, peak_src = peaks()
        traffic, traffic_src, tj_all = None, None, {}
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):  # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full captures
            tj_all = json.load(open(tp))
            tj = tj_all.get("pfn_tile_apply_eval", {})
            if tj.get("rows") == n_kept and tj.get("pillars") == n_pil:
                traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]

        def entry(name, alg_bytes, ms, key=None):
            e = {"kernel": name, "algorithmic_bytes": int(alg_bytes), "ms": ms, "achieved": alg_bytes / (ms * 1e-3) / 1e9,
                 "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak}
            t = tj_all.get(key) if key else None
            if t and t.get("rows") == n_kept and t.get("pillars") == n_pil:
                e["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
            return e

        roof = {"bound": "hbm", "kernel": "pfn_tile_kernel<PfnCfg<6 cols, Simple2D, 32 ch>, APPLY> (LiDAR batch of 8 frames, eval BN)",
                "achieved": alg / (dur * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (dur * 1e-3) / 1e9 / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": dur, "algorithmic_bytes": int(alg),
                "frac_of_8000_nominal": alg / (dur * 1e-3) / 1e9 / 8000.0,
                "whole_forward": {"algorithmic_bytes": int(alg_fwd), "ms": idx + dur, "index_ms": idx,
                                  "achieved": alg_fwd / ((idx + dur) * 1e-3) / 1e9,
                                  "frac": alg_fwd / ((idx + dur) * 1e-3) / 1e9 / peak},
                "other_calls": [
                    entry("rdp_index_fwd: 7 index kernels + publish", row_bytes * len(lidar) + 4 * n_kept + n_pil * (4 * spec.coord_cols + 4), idx),
                    entry("rdp_pfn_fwd train: pfn_tile<STATS> + bn_finalize + pfn_tile<APPLY_ARG>", alg_train, dur_train, "pfn_train_fwd"),
                    entry("rdp_pfn_bwd: pfn_tile<BWD> + bwd_finalize", alg_bwd, dur_bwd, "pfn_tile_bwd")]}

    if rank == 0:
        cpu = None
        if not ddp:
            cores = os.cpu_count() or 1
            fr = 2
            l2, r2 = make_clouds(0, fr)
            v, dt = time_cpu(l2, r2, mode, 3, 1, cores)
            cpu = {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"{fr} of the {frames} paired frames ({len(l2)} LiDAR + {len(r2)} radar rows) x 3 steps, C oracle "
                             f"(oracle/pillar_oracle.c), {cores} threads in the per-point loops, {dt:.2f} s/step"}
        # librdp kernels per step: index 8 (quantise, scan, publish, zero, rank, scan, group, table); forward 3 in train mode
        # (moments, finalize, apply) or 1; backward 2 (tile, finalize).  The radar encoder always trains.
        launches_lidar = 8 + (3 + 2 if mode == "B" else 1)
        launches = (launches_lidar + 13) * args.steps
        line = {"metric": "pillar-encoder points/s (fwd+bwd)", "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(mode, frames, len(lidar), len(radar)),
                               cpus_pinned=(pinned if pinned is None else len(pinned)),
                               data_parallel=("none" if not ddp else ("GradientAllReduce: one NCCL all_reduce of the flat PFN gradients "
                                              "per step" if args.dp == "lean" else "torch DistributedDataParallel"))),
                "e2e": e2e, "e2e_host_outputs": e2e_out,
                "gpu_launches": launches,
                "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu, "wall_s_timed_region": wall}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if ddp:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="B", choices=["A", "B"])
    ap.add_argument("--pin-cpus", type=int, default=1,
                    help="N > 1: give every rank its own physical cores (sched_setaffinity); 0 = leave the scheduler alone")
    ap.add_argument("--dp", default="lean", choices=["lean", "ddp"],
                    help="N > 1: gradient all-reduce by radardistill_b200.sharding.GradientAllReduce (one all_reduce of a flat "
                         "buffer per step) or by torch DistributedDataParallel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- pillar-encoder throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode B|A]
                    [--scaling weak|strong] [--config paired|stress]

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
roofline = the dominant kernel of the timed step (mode B: pfn_apply_kernel<ARG>, the train-mode forward tile kernel), timed
         live with CUDA events; roofline.whole_step = algorithmic bytes of fwd+bwd of both encoders / the step time
         (the north_star quantity); "configs" = the other BASELINE.json configurations, measured briefly at N = 1
--impl reference: the CPU port of the path (oracle/pillar_oracle.c, all host threads) on the same workload, same frame
         count and warm-up; when build() staged the reference's own torch files (oracle/_ref/) its rows are listed too.
--scaling strong: BASELINE.json configs[3] -- global batch 64, 64 / N frames per rank.   --config stress: configs[4] --
         1 M-point clouds at 0.05 m pillars, global batch 16, 16 / N frames per rank.
Multi-GPU: frames shard over ranks; the only collective is the gradient all-reduce of the PFN parameters.
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
STRONG_GLOBAL_BATCH = 64      # BASELINE.json configs[3]
STRESS_GLOBAL_BATCH = 16      # BASELINE.json configs[4]
S2D_CFG = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
DYN_CFG = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64])   # DynamicPillarVFE 10 -> 64


class Cfg(dict):
    __getattr__ = dict.__getitem__


def make_clouds(rank: int, frames: int, config: str = "paired"):
    from radardistill_b200 import synth
    if config == "stress":   # configs[4]: dense 1 M-point clouds (skewed pillars, 20 % exact duplicates) + the usual radar frames
        lidar = synth.collate([synth.stress_frame(rank * frames + b) for b in range(frames)])
    else:
        lidar = synth.collate([synth.lidar_frame(rank * frames + b) for b in range(frames)])
    radar = synth.collate([synth.radar_frame(rank * frames + b) for b in range(frames)])
    return lidar, radar


def voxel_of(config: str):
    from radardistill_b200 import synth
    return synth.STRESS_VOXEL_SIZE if config == "stress" else synth.VOXEL_SIZE


def frames_per_rank(args, world: int) -> int:
    if args.config == "stress":
        return max(STRESS_GLOBAL_BATCH // world, 1)
    if args.scaling == "strong":
        return max(STRONG_GLOBAL_BATCH // world, 1)
    return FRAMES_PER_GPU


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
def oracle_pair(seed=0, voxel=None):
    from oracle import oracle as orc
    from radardistill_b200 import synth
    rng = np.random.default_rng(seed)
    voxel = list(voxel if voxel is not None else synth.VOXEL_SIZE)
    out = {}
    for kind, c in (("lidar", 5), ("radar", 6)):
        cfg = orc.OracleConfig(num_point_features=c, voxel_size=tuple(voxel), grid_size=tuple(synth.grid_size_of(voxel_size=voxel)),
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


def time_cpu(lidar, radar, mode, steps, warmup, threads, voxel=None):
    from oracle import oracle as orc
    orc.set_threads(threads)
    oracles = oracle_pair(voxel=voxel)
    for _ in range(warmup):
        cpu_step(oracles, lidar, radar, mode)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(oracles, lidar, radar, mode)
    dt = (time.perf_counter() - t0) / steps
    return (len(lidar) + len(radar)) / dt, dt


def time_torch_reference(mode, cores, frames=1):
    """The reference's OWN torch path (DynamicPillarVFESimple2D + Radar_DynamicPillarVFESimple2D loaded from the files that
    build() staged in oracle/_ref/, torch_scatter restated by oracle/ref_loader.py) on the host cores: `frames` paired
    frames, one warm-up + two timed steps per row.  Two rows, as BASELINE.md section 3 plans: as-written, and with
    `dim=0` dropped from the 1-D torch.unique call (bit-identical output, removes the aten::unique_dim pathology).
    Returns [] when the files are not staged."""
    try:
        import torch
        from oracle import ref_loader
        if not ref_loader.reference_available():
            return []
        from radardistill_b200 import synth
        torch.set_num_threads(cores)
        lidar, radar = make_clouds(0, frames)
        grid = synth.grid_size_of()
        mods = {"lidar": ref_loader.build_reference("DynamicPillarVFESimple2D", S2D_CFG, 5, synth.VOXEL_SIZE, grid, synth.PC_RANGE, force_cpu=True),
                "radar": ref_loader.build_reference("Radar_DynamicPillarVFESimple2D", S2D_CFG, 6, synth.VOXEL_SIZE, grid, synth.PC_RANGE, force_cpu=True)}
        if mode == "A":
            mods["lidar"].eval()
        pts = {"lidar": torch.from_numpy(lidar), "radar": torch.from_numpy(radar)}

        def step():
            for kind, key, out_key in (("lidar", "points", "pillar_features"), ("radar", "radar_points", "radar_pillar_features")):
                m = mods[kind]
                train = mode == "B" or kind == "radar"
                with ref_loader.cpu_cuda_identity(True), torch.set_grad_enabled(train):
                    out = m({key: pts[kind]})
                if train:
                    out[out_key].backward(torch.ones_like(out[out_key]))
                    m.zero_grad(set_to_none=True)

        rows = []
        orig_unique = torch.unique

        def unique_1d(x, *a, **k):
            if x.dim() == 1:
                k.pop("dim", None)
            return orig_unique(x, *a, **k)

        for variant, patch in (("as written", None), ("dim=0 dropped from the 1-D torch.unique (same output)", unique_1d)):
            if patch is not None:
                torch.unique = patch
            try:
                step()
                t0 = time.perf_counter()
                for _ in range(2):
                    step()
                dt = (time.perf_counter() - t0) / 2
            finally:
                torch.unique = orig_unique
            rows.append({"value": (len(lidar) + len(radar)) / dt, "unit": "points/s", "cores": cores, "kind": "reference",
                         "variant": variant, "s_per_step": dt,
                         "sample": f"{frames} paired frame(s) ({len(lidar)} LiDAR + {len(radar)} radar rows), mode {mode}, 1 warm-up + 2 timed "
                                   f"steps, torch {torch.__version__} CPU, reference classes from oracle/_ref, torch_scatter restated"})
        return rows
    except Exception as e:   # the baseline rows are a report, never a reason to lose the bench line
        return [{"kind": "reference", "unavailable": f"{type(e).__name__}: {e}"}]


def run_reference(args):
    """--impl reference: the CPU port of the reference path on the host cores (rank 0 only), on the GPU arm's workload:
    same frame count, same warm-up, same step count."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    frames = frames_per_rank(args, world)
    if args.scaling == "strong" or args.config == "stress":
        frames = min(frames, 8)   # bounded sample of the rank-0 shard: the rate is per point
    lidar, radar = make_clouds(0, frames, args.config)
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    value, dt = time_cpu(lidar, radar, args.mode, steps, warmup, cores, voxel_of(args.config))
    torch_rows = time_torch_reference(args.mode, cores) if args.config == "paired" else []
    line = {"impl": "reference", "metric": "pillar-encoder points/s (fwd+bwd)", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.mode, frames, len(lidar), len(radar)),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": f"{frames} paired frames/step ({len(lidar)} LiDAR + {len(radar)} radar rows), C oracle "
                                       f"(oracle/pillar_oracle.c), {cores} threads in the per-point loops"},
            "cpu_baseline_reference_torch": torch_rows,
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, mode, frames, n_lidar, n_radar):
    from radardistill_b200 import synth
    vox = voxel_of(args.config)
    grid = synth.grid_size_of(voxel_size=vox)
    what = ("paired radar+LiDAR pillar encoding fwd+bwd (BASELINE.json configs[2]; radar_distill_train.yaml grid 1440x1440, 0.075 m pillars)"
            if args.config == "paired" else
            "stress: 1 M-point dense clouds at 0.05 m pillars (2160x2160), skewed pillars + 20 % exact duplicates, paired with the radar "
            "frames, fwd+bwd (BASELINE.json configs[4], global batch 16)")
    if args.config == "paired" and args.scaling == "strong":
        what += f"; strong scaling: global batch {STRONG_GLOBAL_BATCH} (configs[3])"
    return {"workload": what, "mode": "B: both encoders train-BN fwd+bwd" if mode == "B" else
            "A: LiDAR frozen eval fwd, radar train fwd+bwd", "frames_per_gpu": frames, "lidar_rows_per_gpu": int(n_lidar),
            "radar_rows_per_gpu": int(n_radar), "grid": [int(grid[0]), int(grid[1])],
            "encoders": "DynamicPillarVFESimple2D 14->32 + Radar_DynamicPillarVFESimple2D 15->32",
            "l2": "256 MiB scratch written between timed steps (L2 flushed)"}


# ------------------------------------------------------------------------------------------------ GPU arm
def build_modules(device, mode, ddp, voxel=None):
    import torch
    from radardistill_b200 import synth, vfe
    from radardistill_b200.vfe import forward_pair
    torch.manual_seed(1234)
    voxel = list(voxel if voxel is not None else synth.VOXEL_SIZE)
    grid = synth.grid_size_of(voxel_size=voxel)
    lid = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D_CFG), num_point_features=5, voxel_size=voxel,
                                       grid_size=grid, point_cloud_range=synth.PC_RANGE).to(device)
    rad = vfe.Radar_DynamicPillarVFESimple2D(model_cfg=Cfg(S2D_CFG), num_point_features=6, voxel_size=voxel,
                                             grid_size=grid, point_cloud_range=synth.PC_RANGE).to(device)
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
    mode, frames = args.mode, frames_per_rank(args, world)
    lidar, radar = make_clouds(rank, frames, args.config)
    n_rows = len(lidar) + len(radar)
    lid, rad, call = build_modules(device, mode, args.dp if ddp else False, voxel_of(args.config))
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
        return _timed_steps(m, k, [])

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
    # Every step ends in a rendezvous of all ranks (the gradient all-reduce), so a cyclic-GC pause on any one rank stalls all of
    # them: the collector is parked over the measurement (as long-running training loops do: collect at a step boundary of
    # your choosing, not in the middle of a launch sequence).  It is parked BEFORE the warm-up and the barrier, so that the
    # ranks enter the timed region aligned -- a collection of different length per rank right before the first timed step
    # would be charged to that step on every rank.
    gc.collect()
    gc.freeze()
    gc.disable()
    for _ in range(max(args.warmup, 3)):
        gpu_step(call, lidar_dev, radar_dev, mode, frames, upstream)
    barrier()
    clk.rows.clear()  # keep only samples taken during the timed region
    t_wall0 = time.perf_counter()
    ms = timed_steps(mode, args.steps)
    barrier()
    wall = time.perf_counter() - t_wall0
    step_ms = sum(ms) / len(ms)
    step_stats = {"median": statistics.median(ms), "min": min(ms), "max": max(ms)}   # this rank's steps (rank 0 reports its own)
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
        lidA, radA, callA = build_modules(device, "A", False, voxel_of(args.config))
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

    # ---- roofline: the librdp calls on the LiDAR batch, each timed live with CUDA events on the launching stream, L2
    #      flushed before every launch.  Headline kernel = the dominant kernel of the timed step: mode B (train forward +
    #      backward) spends its largest share in pfn_apply_kernel<ARG = true> -- the forward tile kernel that also writes the
    #      argmax; an eval-BN rdp_pfn_fwd call that asks for the argmax launches exactly that kernel and nothing else.  Mode A
    #      (frozen LiDAR) is dominated by the eval instantiation <ARG = false>.
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
        rres = ops.encode_forward(radar_dev, rad.spec, frames, rad.pfn_layers[0].linear.weight.detach(), None,
                                  rad.pfn_layers[0].norm.weight, rad.pfn_layers[0].norm.bias,
                                  rad.pfn_layers[0].norm.running_mean.clone(), rad.pfn_layers[0].norm.running_var.clone(), True, True)
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
        dur_eval = timed(lambda: call_fwd(prm_eval, None, None), reps)            # pfn_apply_kernel<ARG = false> alone
        dur_arg = timed(lambda: call_fwd(prm_eval, P_(argp), None), reps)          # pfn_apply_kernel<ARG = true> alone
        dur_train = timed(lambda: call_fwd(prm_train, P_(argp), P_(res.bn_state)), reps)   # table+moments kernel, then <ARG = true>
        dur_bwd = timed(call_bwd, reps)
        n_kept, n_pil = res.n_kept, res.n_pillars
        row_bytes = 4 * spec.cols
        peak, peak_src = peaks()

        def alg_forward(sp, n0, nk, npil, train):   # SURVEY 8(d) B_fwd (+ the argmax in train mode)
            return 4 * sp.cols * n0 + 4 * nk + npil * (4 * sp.c_out + 4 * sp.coord_cols + 4) + (4 * sp.c_out * npil if train else 0)

        def alg_backward(sp, nk, npil):             # SURVEY 8(d) B_bwd
            return 8 * sp.c_out * npil + 4 * sp.cols * nk + 4 * nk

        alg_apply = row_bytes * n_kept + 4 * spec.c_out * n_pil           # rows read once + feature rows written once
        alg_apply_arg = alg_apply + 4 * spec.c_out * n_pil                # + the argmax rows
        alg_fwd = alg_forward(spec, len(lidar), n_kept, n_pil, False)
        alg_train = 2 * row_bytes * n_kept + 8 * spec.c_out * n_pil       # the statistics pass re-reads the rows; features + argmax written
        alg_bwd = alg_backward(spec, n_kept, n_pil)
        # the whole timed step: forward (+ backward where the mode trains) of BOTH encoders
        step_bytes = alg_forward(rad.spec, len(radar), rres.n_kept, rres.n_pillars, True) + alg_backward(rad.spec, rres.n_kept, rres.n_pillars)
        step_bytes += alg_forward(spec, len(lidar), n_kept, n_pil, mode == "B") + (alg_bwd if mode == "B" else 0)
        tj_all = {}
        tp = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tp):  # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full captures
            tj_all = json.load(open(tp))

        def traffic_of(key):
            t = tj_all.get(key)
            if t and t.get("rows") == n_kept and t.get("pillars") == n_pil:
                return t["dram_bytes_read"] + t["dram_bytes_write"], t.get("source")
            return None, None

        def entry(name, alg_bytes, ms, key=None):
            e = {"kernel": name, "algorithmic_bytes": int(alg_bytes), "ms": ms, "achieved": alg_bytes / (ms * 1e-3) / 1e9,
                 "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak}
            t, _ = traffic_of(key) if key else (None, None)
            if t is not None:
                e["traffic"] = t
            return e

        if mode == "B":
            head_name, head_alg, head_ms, head_key = ("pfn_apply_kernel<PfnCfg<6 cols, no distance, 32 ch>, ARG = true> (LiDAR batch, train forward: "
                                                      "features + argmax)"), alg_apply_arg, dur_arg, "pfn_apply_arg"
        else:
            head_name, head_alg, head_ms, head_key = ("pfn_apply_kernel<PfnCfg<6 cols, no distance, 32 ch>, ARG = false> (LiDAR batch, eval BN "
                                                      "forward)"), alg_apply, dur_eval, "pfn_apply_eval"
        traffic, traffic_src = traffic_of(head_key)
        roof = {"bound": "hbm", "kernel": head_name, "achieved": head_alg / (head_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": head_alg / (head_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms": head_ms, "algorithmic_bytes": int(head_alg),
                "frac_of_8000_nominal": head_alg / (head_ms * 1e-3) / 1e9 / 8000.0,
                "whole_step": {"what": f"mode {mode}: forward{' + backward' if mode == 'B' else ''} of the LiDAR encoder and forward + backward "
                                       "of the radar encoder, SURVEY 8(d) algorithmic bytes / the timed step (ms_per_step)",
                               "algorithmic_bytes": int(step_bytes), "ms": step_ms, "achieved": step_bytes / (step_ms * 1e-3) / 1e9,
                               "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                               "frac_of_8000_nominal": step_bytes / (step_ms * 1e-3) / 1e9 / 8000.0},
                "whole_forward": {"algorithmic_bytes": int(alg_fwd), "ms": idx + dur_eval, "index_ms": idx,
                                  "achieved": alg_fwd / ((idx + dur_eval) * 1e-3) / 1e9,
                                  "frac": alg_fwd / ((idx + dur_eval) * 1e-3) / 1e9 / peak},
                "other_calls": [
                    entry("rdp_index_fwd: quantize_mark, bitmap_rank, rank_count, count_scan, group_rows, pillar_table",
                          row_bytes * len(lidar) + 4 * n_kept + n_pil * (4 * spec.coord_cols + 4), idx, "index_call"),
                    entry("pfn_apply_kernel<ARG = false> (eval forward)", alg_apply, dur_eval, "pfn_apply_eval"),
                    entry("pfn_apply_kernel<ARG = true> (train forward)", alg_apply_arg, dur_arg, "pfn_apply_arg"),
                    entry("rdp_pfn_fwd train: pillar_table_stats_kernel (table + BatchNorm moments) + pfn_apply_kernel<ARG = true>",
                          alg_train, dur_train, "pfn_train_fwd"),
                    entry("rdp_pfn_bwd: pfn_bwd_kernel (closed-form epilogue in its last CTA)", alg_bwd, dur_bwd, "pfn_bwd")]}

    # ---- the other BASELINE.json configurations, measured briefly (N = 1, rank 0, default workload only)
    others = None
    if rank == 0 and not ddp and args.config == "paired" and args.scaling == "weak" and not args.no_configs:
        others = other_configs(device, flush, max(min(args.steps, 10), 3))

    if rank == 0:
        cpu, cpu_torch = None, []
        if not ddp:
            cores = os.cpu_count() or 1
            fr = min(frames, 8)
            l2, r2 = (lidar, radar) if fr == frames else make_clouds(0, fr, args.config)
            v, dt = time_cpu(l2, r2, mode, 3, 1, cores, voxel_of(args.config))
            cpu = {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"{fr} of the {frames} paired frames ({len(l2)} LiDAR + {len(r2)} radar rows) x 3 steps after 1 warm-up, C oracle "
                             f"(oracle/pillar_oracle.c), {cores} threads in the per-point loops, {dt:.2f} s/step"}
            if args.config == "paired":
                cpu_torch = time_torch_reference(mode, cores)
        # librdp kernels per step and encoder: index 5 (quantise, bitmap scan, rank, count scan, group) + table (train: table +
        # moments) + apply = 7 per forward, + 1 backward tile kernel (its epilogue runs in its last CTA).  The radar encoder always trains.
        launches_lidar = 7 + (1 if mode == "B" else 0)
        red = upstream.get("reducer")
        per_step = launches_lidar + 8 + (1 if red is not None and getattr(red, "p2p", None) is not None else 0)   # + rdp_allreduce_small
        launches = per_step * args.steps
        line = {"metric": "pillar-encoder points/s (fwd+bwd)", "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(workload_config(args, mode, frames, len(lidar), len(radar)),
                               cpus_pinned=(pinned if pinned is None else len(pinned)),
                               data_parallel=("none" if not ddp else (
                                   ("GradientAllReduce: rdp_allreduce_small, one kernel per step over NVLink peer memory"
                                    if getattr(upstream.get("reducer"), "p2p", None) is not None else
                                    "GradientAllReduce: one NCCL all_reduce of the flat PFN gradients per step")
                                   if args.dp == "lean" else "torch DistributedDataParallel"))),
                "e2e": e2e, "e2e_host_outputs": e2e_out,
                "gpu_launches": launches,
                "ms_per_step_rank0": step_stats, "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu, "cpu_baseline_reference_torch": cpu_torch,
                "wall_s_timed_region": wall}
        if others is not None:
            line["configs"] = others
        line.update(extra)
        print(json.dumps(line), flush=True)
    gc.enable()
    gc.unfreeze()
    if ddp:
        dist.barrier()
        dist.destroy_process_group()


def other_configs(device, flush, steps):
    """BASELINE.json configs[0], [1] and [4] at N = 1, a few L2-flushed steps each (device-timed with CUDA events):
    one radar frame (fwd+bwd), one LiDAR frame forward through DynamicPillarVFESimple2D 14->32 and DynamicPillarVFE 10->64
    (us per frame), and the 1 M-point stress clouds at 0.05 m pillars (2 frames = one rank's share of the batch of 16)."""
    import torch
    from radardistill_b200 import synth, vfe
    out = {}

    def timed(fn):
        ts = []
        for i in range(steps + 3):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    def module(cls, cfg, c, voxel=synth.VOXEL_SIZE):
        torch.manual_seed(7)
        m = cls(model_cfg=Cfg(cfg), num_point_features=c, voxel_size=list(voxel), grid_size=synth.grid_size_of(voxel_size=list(voxel)),
                point_cloud_range=synth.PC_RANGE).to(device)
        n = m.pfn_layers[-1].norm
        with torch.no_grad():
            n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
        return m

    # configs[1]: one 10-sweep LiDAR frame, eval forward
    frame = synth.collate([synth.lidar_frame(0)])
    pts = torch.from_numpy(frame).to(device)
    m = module(vfe.DynamicPillarVFESimple2D, S2D_CFG, 5).eval()
    with torch.no_grad():
        ms = timed(lambda: m({"points": pts, "batch_size": 1}))
        p = int(m({"points": pts, "batch_size": 1})["pillar_features"].shape[0])
    out["lidar_frame_forward_simple2d_14to32"] = {"config": "configs[1]: LiDAR-teacher DynamicPillarVFESimple2D eval forward, one 10-sweep frame, batch 1",
                                                  "us_per_frame": ms * 1e3, "points": int(len(frame)), "pillars": p,
                                                  "points_per_s": len(frame) / (ms * 1e-3)}
    pts4 = torch.from_numpy(np.ascontiguousarray(frame[:, :5])).to(device)     # b, x, y, z, intensity: DynamicPillarVFE 4 + 6 = 10 -> 64
    m = module(vfe.DynamicPillarVFE, DYN_CFG, 4).eval()
    with torch.no_grad():
        ms = timed(lambda: m({"points": pts4, "batch_size": 1}))
    out["lidar_frame_forward_dynpillar_10to64"] = {"config": "configs[1] with north_star's DynamicPillarVFE (PillarNet nuScenes 0.075 m grid) 10->64, eval forward, batch 1",
                                                   "us_per_frame": ms * 1e3, "points": int(len(frame)), "points_per_s": len(frame) / (ms * 1e-3)}
    # configs[0]: one radar frame (the reference's CPU-runnable case), train forward + backward
    rframe = synth.collate([synth.radar_frame(0)])
    rpts = torch.from_numpy(rframe).to(device)
    m = module(vfe.Radar_DynamicPillarVFESimple2D, S2D_CFG, 6).train()
    g = torch.randn((len(rframe), 32), device=device)

    def radar_step():
        for p_ in m.parameters():
            p_.grad = None
        f = m({"radar_points": rpts, "batch_size": 1})["radar_pillar_features"]
        f.backward(g[:f.shape[0]])
    ms = timed(radar_step)
    from oracle import oracle as orc
    orc.set_threads(os.cpu_count() or 1)
    o = oracle_pair()["radar"]
    t0 = time.perf_counter()
    for _ in range(5):
        r = o.forward(rframe, training=True, keep_intermediates=True)
        o.backward(r, np.ones_like(r["features"]))
    cpu_dt = (time.perf_counter() - t0) / 5
    out["radar_frame_fwd_bwd"] = {"config": "configs[0]: radar-branch encoder, one radar frame, batch 1, train forward + backward",
                                  "us_per_frame": ms * 1e3, "points": int(len(rframe)), "points_per_s": len(rframe) / (ms * 1e-3),
                                  "cpu_port_us_per_frame": cpu_dt * 1e6, "note": "launch-latency bound on the GPU: 8 dependent kernels for 3 k points"}
    # configs[4]: stress clouds, one rank's share (2 of 16 frames), train forward + backward
    sframes = STRESS_GLOBAL_BATCH // 8
    stress = synth.collate([synth.stress_frame(b) for b in range(sframes)])
    spts = torch.from_numpy(stress).to(device)
    m = module(vfe.DynamicPillarVFESimple2D, S2D_CFG, 5, synth.STRESS_VOXEL_SIZE).train()
    g = torch.randn((len(stress), 32), device=device)
    torch.cuda.reset_peak_memory_stats(device)
    base_mem = torch.cuda.memory_allocated(device)

    def stress_step():
        for p_ in m.parameters():
            p_.grad = None
        f = m({"points": spts, "batch_size": sframes})["pillar_features"]
        f.backward(g[:f.shape[0]])
    ms = timed(stress_step)
    out["stress_fwd_bwd"] = {"config": f"configs[4]: 1 M-point dense clouds, 0.05 m pillars (2160x2160), {sframes} frames = one of 8 ranks' share of "
                                       "the batch of 16, train forward + backward",
                             "ms_per_step": ms, "points": int(len(stress)), "pillars": int(m.last_result.n_pillars),
                             "points_per_s": len(stress) / (ms * 1e-3),
                             "peak_step_memory_bytes": int(torch.cuda.max_memory_allocated(device) - base_mem)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="B", choices=["A", "B"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 8 frames per GPU (default); strong: global batch 64 split over the ranks (BASELINE.json configs[3])")
    ap.add_argument("--config", default="paired", choices=["paired", "stress"],
                    help="paired: configs[2]/[3] (default); stress: configs[4], 1 M-point clouds at 0.05 m pillars, global batch 16")
    ap.add_argument("--no-configs", action="store_true", help="skip the brief runs of the other BASELINE.json configurations")
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

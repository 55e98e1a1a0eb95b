"""CPU, world_size 2 over gloo: the frame-sharded (N > 1) host path.

Each rank shards the same global batch, encodes its frames with the CPU oracle, and the ranks exchange pillar
coords / counts with all_gather: the union, mapped back to global frame ids, must equal the oracle on the full
batch bit for bit (frames are independent -- no data-path collective is needed).  Also checks the max-over-ranks
timing reduction bench.py uses and a DDP-style gradient all-reduce of the parameter gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from radardistill_b200 import sharding, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle(c=6):
    cfg = orc.OracleConfig(num_point_features=c, voxel_size=tuple(synth.VOXEL_SIZE), grid_size=tuple(synth.grid_size_of()),
                           point_cloud_range=tuple(synth.PC_RANGE))
    rng = np.random.default_rng(5)
    return orc.PillarOracle(cfg, (rng.standard_normal((32, cfg.c_in)) * 0.2).astype(np.float32), rng.uniform(0.5, 1.5, 32),
                            rng.normal(0, 0.2, 32), rng.normal(0, 1, 32), rng.uniform(0.5, 4, 32))


def _worker(rank, world, port, global_batch, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = synth.radar_batch(global_batch, n_points=400)
        local, local_batch = sharding.shard_points(pts, global_batch, rank, world)
        assert local_batch == len(sharding.frames_of_rank(global_batch, rank, world))
        o = _oracle()
        r = o.forward(local, training=True)
        b = o.backward(r, np.ones_like(r["features"]))
        coords = sharding.unshard_coords(r["coords"], rank, world)
        # exchange variable-length results (pad to the max P)
        p = torch.tensor([r["p"]])
        ps = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(ps, p)
        pmax = int(max(int(x) for x in ps))
        pad = torch.zeros((pmax, 3), dtype=torch.int32)
        pad[:r["p"]] = torch.from_numpy(coords)
        gathered = [torch.zeros((pmax, 3), dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gathered, pad)
        # DDP-style gradient averaging of the PFN parameters (the step's only collective)
        g = torch.from_numpy(b["d_weight"].copy())
        dist.all_reduce(g)
        g /= world
        # the package's reducer: one all-reduce of the flat PFN gradients, p.grad <- average over ranks
        w = torch.nn.Parameter(torch.zeros(32, 15)); gm = torch.nn.Parameter(torch.zeros(32)); frozen = torch.nn.Parameter(torch.zeros(3), requires_grad=False)
        w.grad = torch.full((32, 15), float(rank + 1)); gm.grad = torch.arange(32, dtype=torch.float32) * (rank + 1)
        red = sharding.GradientAllReduce([w, gm, frozen])
        red.reduce()
        assert red.flat.numel() == 32 * 15 + 32
        assert torch.equal(w.grad, torch.full((32, 15), (1 + world) / 2.0))
        assert torch.allclose(gm.grad, torch.arange(32, dtype=torch.float32) * (1 + world) / 2.0)
        assert w.grad.data_ptr() == red.flat.data_ptr()   # gradients are views of the flat buffer
        # max-over-ranks timing, as bench.py reports it
        t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            out_q.put(dict(ps=[int(x) for x in ps], coords=[gt[:int(n)].numpy() for gt, n in zip(gathered, ps)],
                           grad=g.numpy(), tmax=float(t), rows=len(local)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_matches_global_oracle():
    world, global_batch = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, global_batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _oracle().index(synth.radar_batch(global_batch, n_points=400))
    union = np.concatenate(res["coords"], 0)
    assert sum(res["ps"]) == full["p"]
    # pillar order inside a frame is preserved; across ranks only the frame interleaving differs
    key = lambda c: (c[:, 0].astype(np.int64) * 1440 + c[:, 2]) * 1440 + c[:, 1]
    order = np.argsort(key(union), kind="stable")
    np.testing.assert_array_equal(union[order], full["coords"])
    assert res["tmax"] == pytest.approx(0.020)
    assert np.isfinite(res["grad"]).all()


def test_shard_points_properties():
    pts = synth.lidar_batch(3, sweeps=1, beams=4, azimuths=64)
    total = 0
    for r in range(2):
        loc, lb = sharding.shard_points(pts, 3, r, 2)
        assert lb == (2 if r == 0 else 1)
        assert set(np.unique(loc[:, 0]).astype(int)) <= set(range(lb))
        total += len(loc)
        back = loc.copy()
        back[:, 0] = back[:, 0] * 2 + r
        ref = pts[(pts[:, 0].astype(int) % 2) == r]
        np.testing.assert_array_equal(back, ref)
    assert total == len(pts)
    assert sharding.aggregate_throughput([100, 300], [0.5, 1.0]) == 400.0
    with pytest.raises(ValueError):
        sharding.frames_of_rank(4, 2, 2)

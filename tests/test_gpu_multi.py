"""GPU, world_size 2 over NCCL (skipped on a single-GPU box): the frame-sharded CUDA path.

Each rank encodes its frames of the same global batch with the CUDA modules (train-mode forward + backward) and
``sharding.GradientAllReduce`` averages the PFN parameter gradients -- the step's only collective, what DDP does at
``tools/train.py:175-176``.  Checked against the C oracle:
  * per-rank BatchNorm (the reference default, no --sync_bn): the union of the ranks' pillar coords equals the oracle on the
    global batch, and the averaged gradients equal the mean of the oracle's per-shard gradients;
  * SyncBatchNorm (``tools/train.py:34,144-145``): features, running statistics and summed gradients equal the oracle run
    once on the UNION batch (global batch statistics).
"""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _module(kind, sync_bn):
    from oracle.ref_loader import Cfg
    from radardistill_b200 import synth, vfe
    cfg = Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])
    cls, c = (vfe.DynamicPillarVFESimple2D, 5) if kind == "lidar" else (vfe.Radar_DynamicPillarVFESimple2D, 6)
    torch.manual_seed(7)
    m = cls(model_cfg=cfg, num_point_features=c, voxel_size=synth.VOXEL_SIZE, grid_size=synth.grid_size_of(),
            point_cloud_range=synth.PC_RANGE)
    n = m.pfn_layers[0].norm
    with torch.no_grad():
        n.weight.uniform_(0.5, 1.5); n.bias.normal_(0, 0.2); n.running_mean.normal_(0, 1); n.running_var.uniform_(0.5, 4)
    if sync_bn:
        m = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)   # what tools/train.py:144-145 does to the whole model
    return m.cuda().train()


def _oracle_of(m, c):
    from radardistill_b200 import synth
    pfn = m.pfn_layers[0]
    cfg = orc.OracleConfig(num_point_features=c, voxel_size=tuple(synth.VOXEL_SIZE), grid_size=tuple(synth.grid_size_of()),
                           point_cloud_range=tuple(synth.PC_RANGE))
    cp = lambda t: t.detach().cpu().numpy().copy()
    return orc.PillarOracle(cfg, cp(pfn.linear.weight), cp(pfn.norm.weight), cp(pfn.norm.bias), cp(pfn.norm.running_mean),
                            cp(pfn.norm.running_var))


def _global_batch(kind, frames):
    from radardistill_b200 import synth
    return synth.lidar_batch(frames, sweeps=2) if kind == "lidar" else synth.radar_batch(frames)


def _worker(rank, world, port, kind, frames, sync_bn, out_q):
    import torch.distributed as dist
    from radardistill_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pts = _global_batch(kind, frames)
        local, local_batch = sharding.shard_points(pts, frames, rank, world)
        m = _module(kind, sync_bn)
        key = "points" if kind == "lidar" else "radar_points"
        out = m({key: torch.from_numpy(local).cuda(), "batch_size": local_batch})
        fk = [k for k in out if k.endswith("pillar_features")][0]
        ck = [k for k in out if k.endswith("_coords")][0]
        f = out[fk]
        gout = torch.randn(f.shape, generator=torch.Generator().manual_seed(100 + rank)).cuda()
        f.backward(gout)
        red = sharding.GradientAllReduce(list(m.parameters()))
        red.reduce()
        torch.cuda.synchronize()
        pfn = m.pfn_layers[0]
        out_q.put(dict(rank=rank, coords=sharding.unshard_coords(out[ck].cpu().numpy(), rank, world),
                       features=f.detach().cpu().numpy(), gout=gout.cpu().numpy(),
                       dW=pfn.linear.weight.grad.cpu().numpy(), dg=pfn.norm.weight.grad.cpu().numpy(),
                       db=pfn.norm.bias.grad.cpu().numpy(), rm=pfn.norm.running_mean.cpu().numpy(),
                       rv=pfn.norm.running_var.cpu().numpy(), nbt=int(pfn.norm.num_batches_tracked)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(kind, frames, sync_bn):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, frames, sync_bn, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["radar", "lidar"])
def test_two_gpu_sharded_step_matches_oracle(kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from radardistill_b200 import sharding
    frames, c = 4, (5 if kind == "lidar" else 6)
    res = _run(kind, frames, sync_bn=False)
    pts = _global_batch(kind, frames)
    o = _oracle_of(_module(kind, False), c)
    full = o.index(pts)
    union = np.concatenate([r["coords"] for r in res], 0)
    nx = ny = 1440
    keyf = lambda cc: (cc[:, 0].astype(np.int64) * nx + cc[:, 2]) * ny + cc[:, 1]
    np.testing.assert_array_equal(union[np.argsort(keyf(union), kind="stable")], full["coords"])
    exp = []
    for r in res:
        local, _ = sharding.shard_points(pts, frames, r["rank"], 2)
        fw = o.forward(local, training=True)
        assert H.norm_rel_err(r["features"], fw["features"]) <= 1e-6
        exp.append(o.backward(fw, r["gout"]))
    for name, k in (("d_weight", "dW"), ("d_gamma", "dg"), ("d_beta", "db")):
        want = (exp[0][name].astype(np.float64) + exp[1][name]) / 2
        for r in res:      # every rank holds the same average
            assert H.norm_rel_err(r[k], want) <= H.RTOL_GRADS


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["radar", "lidar"])
def test_two_gpu_sync_bn_matches_oracle_on_the_union_batch(kind):
    """SyncBatchNorm: batch statistics over the points of ALL ranks (tools/train.py:34,144-145)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from radardistill_b200 import sharding
    frames, c = 4, (5 if kind == "lidar" else 6)
    res = _run(kind, frames, sync_bn=True)
    pts = _global_batch(kind, frames)
    # the oracle on the union batch, frames in rank-major order so that its pillar order is rank 0's pillars then rank 1's
    shards = [sharding.shard_points(pts, frames, r, 2) for r in range(2)]
    merged = []
    off = 0
    for loc, lb in shards:
        loc = loc.copy()
        loc[:, 0] += off
        off += lb
        merged.append(loc)
    o = _oracle_of(_module(kind, False), c)
    fw = o.forward(np.concatenate(merged, 0), training=True)
    feats = np.concatenate([r["features"] for r in res], 0)
    assert feats.shape == fw["features"].shape
    assert H.norm_rel_err(feats, fw["features"]) <= 1e-6
    b = o.backward(fw, np.concatenate([r["gout"] for r in res], 0))
    for name, k in (("d_weight", "dW"), ("d_gamma", "dg"), ("d_beta", "db")):
        for r in res:      # GradientAllReduce averages: each rank's local SyncBN gradient sums to the union gradient
            assert H.norm_rel_err(r[k] * 2.0, b[name]) <= H.RTOL_GRADS
    for r in res:
        assert H.norm_rel_err(r["rm"], fw["new_running_mean"]) <= 1e-6
        assert H.norm_rel_err(r["rv"], fw["new_running_var"]) <= 1e-6
        assert r["nbt"] == 1

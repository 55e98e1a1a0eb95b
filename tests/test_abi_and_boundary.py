"""CPU: the C-ABI library loads and exports every symbol include/rdp.h declares (no compute calls), and the
drop-in modules keep the reference's constructor / state_dict / registry contract."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from radardistill_b200 import _lib, ops, synth, vfe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rdp.h")).read()
    return sorted(set(re.findall(r"RDP_API\s+[\w\s\*]+?\b(rdp_\w+)\s*\(", txt)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 9 and set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.rdp_abi_version() == _lib.RDP_ABI_VERSION
    assert lib.rdp_status_string(0) == b"ok"
    assert b"workspace" in lib.rdp_status_string(-2)


def test_workspace_bytes_and_argument_checks_need_no_gpu():
    lib = _lib.load()
    spec = ops.make_spec(5, synth.VOXEL_SIZE, synth.grid_size_of(), synth.PC_RANGE, _lib.LAYOUT_SIMPLE2D, True, True, True, False, 32)
    assert (spec.c_in, spec.coord_cols, spec.nx, spec.ny) == (14, 3, 1440, 1440)
    assert spec.off == (float(np.float32(-53.9625)), float(np.float32(-53.9625)), float(np.float32(-4.9)))
    geom, layout = spec.geom(8), spec.layout_struct()
    n = C.c_size_t(0)
    assert lib.rdp_workspace_bytes(2_700_000, C.byref(geom), C.byref(layout), C.byref(n)) == 0
    assert 100e6 < n.value < 400e6
    small = C.c_size_t(0)
    assert lib.rdp_workspace_bytes(10, C.byref(geom), C.byref(layout), C.byref(small)) == 0 and small.value < n.value
    geom_big = spec.geom(2000)  # 2000 * 1440^2 > 2^31 keys
    assert lib.rdp_workspace_bytes(10, C.byref(geom_big), C.byref(layout), C.byref(small)) == -5
    assert lib.rdp_workspace_bytes(10, C.byref(geom), C.byref(layout), None) == -1
    assert lib.rdp_index_fwd(None, 10, C.byref(geom), 3, None, 0, None, None, None, None, None) == -1
    assert lib.rdp_bn_state_doubles(C.byref(layout)) == 4 * 32 + 1 + 14 + 14 * 14


class Cfg(dict):
    __getattr__ = dict.__getitem__


S2D = dict(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, USE_CLUSTER_XYZ=True, NUM_FILTERS=[32])


def test_modules_keep_reference_contract():
    grid = synth.grid_size_of()
    lid = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D), num_point_features=5, voxel_size=synth.VOXEL_SIZE, grid_size=grid,
                                       point_cloud_range=synth.PC_RANGE, depth_downsample_factor=None)
    rad = vfe.Radar_DynamicPillarVFESimple2D(model_cfg=Cfg(S2D), num_point_features=6, voxel_size=synth.VOXEL_SIZE,
                                             grid_size=grid, point_cloud_range=synth.PC_RANGE)
    dyn = vfe.DynamicPillarVFE(model_cfg=Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[64]),
                               num_point_features=4, voxel_size=synth.VOXEL_SIZE, grid_size=grid, point_cloud_range=synth.PC_RANGE)
    keys = ["pfn_layers.0.linear.weight", "pfn_layers.0.norm.weight", "pfn_layers.0.norm.bias", "pfn_layers.0.norm.running_mean",
            "pfn_layers.0.norm.running_var", "pfn_layers.0.norm.num_batches_tracked"]
    assert list(lid.state_dict().keys()) == keys
    assert tuple(lid.state_dict()[keys[0]].shape) == (32, 14)      # LiDAR 14 -> 32
    assert tuple(rad.state_dict()[keys[0]].shape) == (32, 15)      # radar 15 -> 32
    assert tuple(dyn.state_dict()[keys[0]].shape) == (64, 10)      # DynamicPillarVFE 10 -> 64
    assert (lid.get_output_feature_dim(), dyn.get_output_feature_dim()) == (32, 64)
    # PillarNet freezes by class __name__ (pillarnet.py:19-23)
    assert [type(m).__name__ for m in (lid, rad, dyn)] == ["DynamicPillarVFESimple2D", "Radar_DynamicPillarVFESimple2D",
                                                            "DynamicPillarVFE"]
    assert isinstance(rad, vfe.DynamicPillarVFESimple2D) and isinstance(lid, vfe.VFETemplate)
    nonorm = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D, USE_NORM=False), num_point_features=5, voxel_size=synth.VOXEL_SIZE,
                                          grid_size=grid, point_cloud_range=synth.PC_RANGE)
    assert list(nonorm.state_dict().keys()) == ["pfn_layers.0.linear.weight", "pfn_layers.0.linear.bias"]
    reg = vfe.register({"MeanVFE": object})
    assert reg["DynPillarVFE"] is vfe.DynamicPillarVFE and reg["Radar_DynamicPillarVFESimple2D_Test"] is vfe.Radar_DynamicPillarVFESimple2D_Test
    assert reg["MeanVFE"] is object
    # stacked PFN layers (dynamic_pillar_vfe.py:63-72): non-last layers have out_channels // 2 outputs and feed a concat of 2x that
    two = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D, NUM_FILTERS=[32, 64]), num_point_features=5, voxel_size=synth.VOXEL_SIZE,
                                       grid_size=grid, point_cloud_range=synth.PC_RANGE)
    assert tuple(two.state_dict()["pfn_layers.0.linear.weight"].shape) == (16, 14)
    assert tuple(two.state_dict()["pfn_layers.1.linear.weight"].shape) == (64, 32)
    assert not two.fused and lid.fused and two.get_output_feature_dim() == 64
    vox = vfe.DynamicVoxelVFE(model_cfg=Cfg(USE_NORM=True, WITH_DISTANCE=False, USE_ABSLOTE_XYZ=True, NUM_FILTERS=[128, 256]),
                              num_point_features=4, voxel_size=[0.1, 0.1, 0.2], grid_size=[1080, 1080, 40], point_cloud_range=synth.PC_RANGE)
    assert tuple(vox.state_dict()["pfn_layers.0.linear.weight"].shape) == (64, 10) and vox.spec.nz == 40 and not vox.fused
    mean = vfe.DynamicMeanVFE(model_cfg=Cfg(), num_point_features=5, voxel_size=[0.1, 0.1, 0.2], grid_size=[1080, 1080, 40],
                              point_cloud_range=synth.PC_RANGE)
    assert mean.get_output_feature_dim() == 5 and len(mean.state_dict()) == 0
    assert reg["DynamicVoxelVFE"] is vfe.DynamicVoxelVFE and reg["DynMeanVFE"] is vfe.DynamicMeanVFE
    flip = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D, DOUBLE_FLIP=True), num_point_features=5, voxel_size=synth.VOXEL_SIZE,
                                        grid_size=grid, point_cloud_range=synth.PC_RANGE)
    with pytest.raises(NotImplementedError):
        flip({"points": torch.zeros((0, 6))})


def test_no_cpu_fallback():
    lid = vfe.DynamicPillarVFESimple2D(model_cfg=Cfg(S2D), num_point_features=5, voxel_size=synth.VOXEL_SIZE,
                                       grid_size=synth.grid_size_of(), point_cloud_range=synth.PC_RANGE).eval()
    with pytest.raises(_lib.RdpError):
        lid({"points": torch.zeros((4, 6)), "batch_size": 1})


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_state_dict_matches_the_reference_class():
    from oracle import ref_loader as rl
    ref = rl.build_reference("Radar_DynamicPillarVFESimple2D", S2D, 6, synth.VOXEL_SIZE, synth.grid_size_of(), synth.PC_RANGE)
    ours = vfe.Radar_DynamicPillarVFESimple2D(model_cfg=Cfg(S2D), num_point_features=6, voxel_size=synth.VOXEL_SIZE,
                                              grid_size=synth.grid_size_of(), point_cloud_range=synth.PC_RANGE)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(tuple(a[k].shape) == tuple(b[k].shape) and a[k].dtype == b[k].dtype for k in a)
    ours.load_state_dict(a, strict=True)


def test_torch_library_ops_are_registered_with_fake_kernels():
    """torch.ops.rdp.* exist, carry schemas, and their fake kernels give the data-dependent sizes unbacked symbols."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.fx.experimental.symbolic_shapes import ShapeEnv
    from radardistill_b200 import torch_ops
    spec = ops.make_spec(6, synth.VOXEL_SIZE, synth.grid_size_of(), synth.PC_RANGE, _lib.LAYOUT_SIMPLE2D, True, True, True, False, 32)
    ints, floats = torch_ops.spec_to_lists(spec)
    assert torch_ops.spec_from_lists(ints, floats) == spec
    assert "Tensor points" in str(torch.ops.rdp.pillar_encode.default._schema)
    with FakeTensorMode(shape_env=ShapeEnv()):
        pts, w = torch.empty((500, 7), device="cuda"), torch.empty((32, 15), device="cuda")
        v = lambda: torch.empty(32, device="cuda")
        out = torch.ops.rdp.pillar_encode(pts, w, None, v(), v(), v(), v(), ints, floats, 2, True, True)
        feats, coords, inverse, counts, argpos, bn_state = out[:6]
        assert feats.shape[1] == 32 and coords.shape[1] == 3 and feats.dtype == torch.float32 and coords.dtype == torch.int32
        assert feats.shape[0] == coords.shape[0] == counts.shape[0] == argpos.shape[0]      # one symbol: P
        assert bn_state.shape[0] == 4 * 32 + 1 + 14 + 14 * 14   # reduced basis, upper bound G = 14
        g = torch.ops.rdp.pillar_encode_backward(pts, feats, feats, argpos, bn_state, out[6], out[7], w, None, v(), v(), v(), v(),
                                                 ints, floats, 2, True)
        assert [tuple(t.shape) for t in g] == [(32, 15), (32,), (32,)]


def test_plan_carves_one_scratch_allocation():
    """ops._plan: librdp workspace | counters | BN state | inverse | counts in ONE buffer, 256-byte aligned regions."""
    spec = ops.make_spec(6, synth.VOXEL_SIZE, synth.grid_size_of(), synth.PC_RANGE, _lib.LAYOUT_SIMPLE2D, True, True, True, False, 32)
    for train in (False, True):
        pl = ops._plan(spec, 8, 21_827, train)
        assert pl is ops._plan(spec, 8, 21_827, train)            # cached
        offs = [0, pl.off_counters, pl.off_bn, pl.off_inverse, pl.off_counts, pl.total_bytes]
        assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
        assert pl.off_bn - pl.off_counters >= 4 * _lib.RDP_NUM_COUNTERS
        assert pl.off_inverse - pl.off_bn >= 8 * pl.bn_doubles
        assert pl.off_counts - pl.off_inverse >= 4 * (pl.cap + 4) and pl.total_bytes - pl.off_counts >= 4 * (pl.cap + 4)
        assert pl.bn_doubles == (4 * 32 + 1 + 14 + 14 * 14 if train else 0)
    assert ops._plan(spec, 8, 21_827, True) is not ops._plan(spec, 8, 21_827, False)


def test_collate_frames_is_collate_without_the_batch_column():
    frames = [synth.radar_frame(s, n_points=50 + 7 * s) for s in range(4)]
    frames[2] = frames[2][:0]
    padded = synth.collate(frames)
    raw, offs = synth.collate_frames(frames)
    assert offs.dtype == np.int32 and offs[0] == 0 and offs[-1] == len(raw) == len(padded)
    np.testing.assert_array_equal(raw, padded[:, 1:])
    for b in range(4):
        assert np.all(padded[offs[b]:offs[b + 1], 0] == b)


def test_pin_cpus_partitions_the_allowed_cores():
    import bench
    before = os.sched_getaffinity(0)
    try:
        n = 2 if len(before) >= 2 else 1
        sets = []
        for r in range(n):
            os.sched_setaffinity(0, before)
            got = bench.pin_cpus(r, n)
            assert got and set(got) <= before and os.sched_getaffinity(0) == set(got)
            sets.append(set(got))
        if n == 2:
            assert not (sets[0] & sets[1])
    finally:
        os.sched_setaffinity(0, before)


def test_host_extension_loads_and_matches_the_abi():
    """rdp_torch_ext.so (csrc/rdp_torch.cpp, built by build_ext) is plumbing around the same librdp calls; it must load on a box
    without a GPU and agree with the library's ABI version."""
    from radardistill_b200 import _ext, build
    if build.ext_path() is None:
        pytest.skip("rdp_torch_ext.so has not been built (python -m radardistill_b200.build)")
    mod = _ext.load()
    assert mod is not None and mod.abi_version() == _lib.RDP_ABI_VERSION and hasattr(mod, "pair_forward")


def test_round2_entry_points_validate_their_arguments_without_a_gpu():
    import ctypes as C
    lib = _lib.load()
    spec = ops.make_spec(5, synth.VOXEL_SIZE, synth.grid_size_of(), synth.PC_RANGE, _lib.LAYOUT_SIMPLE2D, True, True, True, False, 32)
    geom, layout = spec.geom(2), spec.layout_struct()
    inv = _lib.RdpError
    # n_points == 0 is a no-op everywhere; missing buffers are refused before anything is launched
    assert lib.rdp_decorate(0, C.byref(geom), C.byref(layout), None, 0, 1, None, None) == 0
    assert lib.rdp_decorate(10, C.byref(geom), C.byref(layout), None, 0, 1, None, None) == -1
    bad = spec.layout_struct(); bad.c_in = 13
    assert lib.rdp_decorate(10, C.byref(geom), C.byref(bad), 1, 1 << 30, 1, 1, None) == -1          # c_in does not match the layout
    assert lib.rdp_segment_max_fwd(None, 32, 10, C.byref(geom), None, 0, 1, None, None, None) == -1
    assert lib.rdp_segment_max_bwd(None, None, 0, 32, None, None) == 0
    assert lib.rdp_voxel_mean(10, C.byref(geom), None, 0, 1, None, None) == -1
    rng = (C.c_float * 4)(-54, -54, 54, 54)
    assert lib.rdp_prepare_points(None, 10, 5, 4, rng, 0, None, 0, None, 1, None) == -1              # y column outside the row
    assert lib.rdp_prepare_scratch_bytes(1000) >= 4 * 4
    assert lib.rdp_allreduce_staging_bytes(1056, 8) >= 2 * 1056 * 4 + 8 * 4
    assert lib.rdp_allreduce_small(None, None, 1, None, 0, 8, 1, 0.125, None, None) == -1
    seg, cnt = (C.c_void_p * 1)(16), (C.c_int32 * 1)(4)
    assert lib.rdp_allreduce_small(seg, cnt, 1, 16, 8, 8, 1, 0.125, 16, None) == -1                  # rank outside the world
    assert lib.rdp_allreduce_small(seg, cnt, 1, 16, 0, 8, 0, 0.125, 16, None) == -1                  # step numbers start at 1
    voxel = ops.make_spec(4, [0.1, 0.1, 0.2], [1080, 1080, 40], synth.PC_RANGE, _lib.LAYOUT_DYNVOXEL, True, True, False, False, 32)
    assert voxel.nz == 40 and voxel.coord_cols == 4 and voxel.c_in == 10
    assert lib.rdp_config_supported(C.byref(voxel.geom(1)), C.byref(voxel.layout_struct())) == 0      # voxels: layer-stack path
    nbytes = C.c_size_t(0)
    assert lib.rdp_workspace_bytes(1000, C.byref(voxel.geom(64)), C.byref(voxel.layout_struct()), C.byref(nbytes)) == -5   # key space > int32
    del inv


def test_shuffle_permutation_of_prepare_points_is_a_bijection():
    """The seeded permutation of rdp_prepare_points (csrc/rdp_stack.cu: perm_index) restated in integers: an odd multiplier, an
    xor-shift and an addition are bijections on [0, 2^k); cycle-walking restricts them to [0, n)."""
    M = (1 << 64) - 1

    def perm(j, n, seed):
        k = 1
        while (1 << k) < n:
            k += 1
        mask = (1 << k) - 1
        v = j
        while True:
            v = (v * (((2 * seed + 0x9E3779B97F4A7C15) & M) | 1)) & mask
            v ^= v >> (k // 2 + 1)
            v = (v * 0xD6E8FEB86659FD93) & mask
            v = (v + seed) & mask
            if v < n:
                return v

    for n in (1, 2, 3, 17, 1000, 4097):
        for seed in (1, 12345, (1 << 63) + 5):
            assert sorted(perm(j, n, seed) for j in range(n)) == list(range(n))

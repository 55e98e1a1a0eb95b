"""ctypes binding of librdp.so (include/rdp.h).  There is no fallback: if the CUDA library is
missing or does not load, importing the encoder fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDP_LIB_PATH", os.path.join(_HERE, "librdp.so"))  # override: kernel-variant experiments

RDP_ABI_VERSION = 7
RDP_NUM_COUNTERS = 16
CNT_N, CNT_P, CNT_ERRFLAGS = 0, 1, 2
LAYOUT_SIMPLE2D, LAYOUT_DYNPILLAR, LAYOUT_DYNVOXEL = 0, 1, 2

EXPORTS = ["rdp_abi_version", "rdp_status_string", "rdp_last_cuda_error", "rdp_workspace_bytes", "rdp_index_fwd", "rdp_index_fwd_publish", "rdp_index_fwd_frames", "rdp_encode_fwd_frames",
           "rdp_pfn_fwd", "rdp_encode_fwd", "rdp_bn_state_doubles", "rdp_pfn_bwd", "rdp_argmax_kept", "rdp_pillar_lookup", "rdp_publish_counters",
           "rdp_encode_host", "rdp_config_supported", "rdp_stats_buffers", "rdp_decorate", "rdp_segment_max_fwd", "rdp_segment_max_bwd",
           "rdp_voxel_mean", "rdp_prepare_scratch_bytes", "rdp_prepare_points", "rdp_allreduce_staging_bytes", "rdp_allreduce_small"]


class Geom(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("vsz", C.c_float * 3), ("off", C.c_float * 3),
                ("nx", C.c_int32), ("ny", C.c_int32), ("batch_size", C.c_int32), ("cols", C.c_int32), ("nz", C.c_int32)]


class Layout(C.Structure):
    _fields_ = [("layout", C.c_int32), ("use_abs", C.c_int32), ("use_cluster", C.c_int32), ("use_relative", C.c_int32),
                ("with_distance", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32), ("coord_cols", C.c_int32)]


class PfnParams(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("eps", C.c_double), ("momentum", C.c_double),
                ("train_bn", C.c_int32), ("num_batches_tracked", C.c_void_p),
                ("stats_phase", C.c_int32), ("local_stats", C.c_void_p), ("global_bwd", C.c_void_p)]


class RdpError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Loads librdp.so (built by ``python -m radardistill_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RdpError(f"{LIB_PATH} is missing: build it with `python -m radardistill_b200.build` "
                       "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the pillar encoder.")
    lib = C.CDLL(LIB_PATH)
    lib.rdp_abi_version.restype = C.c_int
    if lib.rdp_abi_version() != RDP_ABI_VERSION:
        raise RdpError("librdp.so ABI version mismatch -- rebuild")
    lib.rdp_status_string.restype = C.c_char_p
    lib.rdp_status_string.argtypes = [C.c_int]
    lib.rdp_last_cuda_error.restype = C.c_char_p
    lib.rdp_workspace_bytes.restype = C.c_int
    lib.rdp_workspace_bytes.argtypes = [C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(C.c_size_t)]
    vp = C.c_void_p
    lib.rdp_index_fwd.restype = C.c_int
    lib.rdp_index_fwd.argtypes = [vp, C.c_int64, C.POINTER(Geom), C.c_int32, vp, C.c_size_t, vp, vp, vp, vp, vp]
    lib.rdp_index_fwd_publish.restype = C.c_int
    lib.rdp_index_fwd_publish.argtypes = [vp, C.c_int64, C.POINTER(Geom), C.c_int32, vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp]
    lib.rdp_pfn_fwd.restype = C.c_int
    lib.rdp_pfn_fwd.argtypes = [vp, C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(PfnParams), vp, C.c_size_t,
                                vp, vp, vp, vp, vp, vp]
    lib.rdp_encode_fwd.restype = C.c_int
    lib.rdp_encode_fwd.argtypes = [vp, C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(PfnParams), vp, C.c_size_t,
                                   vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rdp_index_fwd_frames.restype = C.c_int
    lib.rdp_index_fwd_frames.argtypes = [vp, vp, C.c_int64, C.POINTER(Geom), C.c_int32, vp, C.c_size_t, vp, vp, vp, vp, vp, vp, vp]
    lib.rdp_encode_fwd_frames.restype = C.c_int
    lib.rdp_encode_fwd_frames.argtypes = [vp, vp, C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(PfnParams), vp, C.c_size_t,
                                          vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rdp_bn_state_doubles.restype = C.c_int64
    lib.rdp_bn_state_doubles.argtypes = [C.POINTER(Layout)]
    lib.rdp_pfn_bwd.restype = C.c_int
    lib.rdp_pfn_bwd.argtypes = [vp, C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(PfnParams), vp, C.c_size_t,
                                vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rdp_argmax_kept.restype = C.c_int
    lib.rdp_argmax_kept.argtypes = [C.c_int64, C.POINTER(Geom), C.POINTER(Layout), vp, C.c_size_t, vp, vp, vp, vp]
    lib.rdp_pillar_lookup.restype = C.c_int
    lib.rdp_pillar_lookup.argtypes = [C.c_int64, C.POINTER(Geom), vp, C.c_size_t, vp, vp]
    lib.rdp_publish_counters.restype = C.c_int
    lib.rdp_publish_counters.argtypes = [vp, vp, vp]
    lib.rdp_config_supported.restype = C.c_int
    lib.rdp_config_supported.argtypes = [C.POINTER(Geom), C.POINTER(Layout)]
    lib.rdp_stats_buffers.restype = C.c_int
    lib.rdp_stats_buffers.argtypes = [C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(C.c_size_t), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
    lib.rdp_encode_host.restype = C.c_int
    lib.rdp_encode_host.argtypes = [vp, C.c_int64, C.POINTER(Geom), C.POINTER(Layout), C.POINTER(PfnParams), vp, vp, vp, vp,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.rdp_decorate.restype = C.c_int
    lib.rdp_decorate.argtypes = [C.c_int64, C.POINTER(Geom), C.POINTER(Layout), vp, C.c_size_t, vp, vp, vp]
    lib.rdp_segment_max_fwd.restype = C.c_int
    lib.rdp_segment_max_fwd.argtypes = [vp, C.c_int32, C.c_int64, C.POINTER(Geom), vp, C.c_size_t, vp, vp, vp, vp]
    lib.rdp_segment_max_bwd.restype = C.c_int
    lib.rdp_segment_max_bwd.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp]
    lib.rdp_voxel_mean.restype = C.c_int
    lib.rdp_voxel_mean.argtypes = [C.c_int64, C.POINTER(Geom), vp, C.c_size_t, vp, vp, vp]
    lib.rdp_prepare_scratch_bytes.restype = C.c_size_t
    lib.rdp_prepare_scratch_bytes.argtypes = [C.c_int64]
    lib.rdp_prepare_points.restype = C.c_int
    lib.rdp_prepare_points.argtypes = [vp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_uint64, vp, C.c_size_t, vp, vp, vp]
    lib.rdp_allreduce_staging_bytes.restype = C.c_size_t
    lib.rdp_allreduce_staging_bytes.argtypes = [C.c_int64, C.c_int32]
    lib.rdp_allreduce_small.restype = C.c_int
    lib.rdp_allreduce_small.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_uint32, C.c_float, vp, vp]
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        lib = load()
        msg = lib.rdp_status_string(status).decode()
        cuda = lib.rdp_last_cuda_error().decode()
        raise RdpError(f"{what} failed: {msg}" + (f" [{cuda}]" if status == -3 and cuda else ""))

"""Loader of the native host path (rdp_torch_ext.so, built from csrc/rdp_torch.cpp by radardistill_b200.build.build_ext).

The extension only replaces Python glue (allocation, streams, events, autograd wiring) around the same librdp.so calls; when
it has not been built the modules use the Python implementation of that glue (ops.py), with identical results.  Set
RDP_NO_HOST_EXT=1 to force the Python glue."""
from __future__ import annotations

import importlib.util
import os

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None
_tried = False


def load():
    """The extension module, or None."""
    global _mod, _tried
    if _tried:
        return _mod
    _tried = True
    path = os.path.join(_HERE, "rdp_torch_ext.so")
    if os.environ.get("RDP_NO_HOST_EXT") or not os.path.exists(path):
        return None
    _lib.load()   # librdp.so first (the extension links against it)
    import torch  # noqa: F401  (libtorch / libc10_cuda must be in the process before the extension is opened)
    spec = importlib.util.spec_from_file_location("rdp_torch_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.abi_version() != _lib.RDP_ABI_VERSION:
        raise _lib.RdpError("rdp_torch_ext.so was built against another librdp ABI -- rebuild (python -m radardistill_b200.build)")
    _mod = mod
    return mod

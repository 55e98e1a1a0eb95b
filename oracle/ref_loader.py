"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference encoder from /root/reference.

Used by ``oracle/gen_golden.py`` (to produce ``tests/golden/*.npz``) and by the
``-m "not gpu"`` tests that pin the C oracle against the reference when the reference
tree is present.  Nothing on the product path (``radardistill_b200/``) imports this,
and ``/root/reference`` does not exist on the GPU box, so nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may call it.

The reference path is ``pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py`` (+ its
base ``vfe_template.py``).  ``import pcdet`` fails here (needs a generated
``version.py``, spconv, SharedArray ...), so the two files are loaded *by path* as a
synthetic package.  Two shims are needed:

1. ``torch_scatter`` (third-party, ``torch-scatter==2.1.1`` per the reference's
   ``docs/INSTALL.md:39``; absent from this image, no network).  Its two functions on
   this path are restated from the library's documented semantics:
     * ``scatter_mean(src, index, dim=0)``: per-segment sum / clamp(count, 1);
     * ``scatter_max(src, index, dim=0)``: ``(out, argmax)``; CPU implementation
       updates on strict ``>`` while walking rows in order => the FIRST (lowest) row
       index wins ties; backward routes the gradient to the single argmax row.
   Call sites: dynamic_pillar_vfe.py:40 (scatter_max), :105/:226/:287/:347 (scatter_mean).
2. ``Tensor.cuda()`` in the constructors (dynamic_pillar_vfe.py:83-85,187-189): there is
   no GPU here, so it is patched to identity while the reference runs on CPU.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch

_REL = os.path.join("pcdet", "models", "backbones_3d", "vfe")
# oracle/_ref/: git-ignored copy of the two reference files this path consists of, made by __graft_entry__.build() in the
# build container so that bench.py's CPU-baseline legs can time the reference's own torch path on the GPU box (where
# /root/reference does not exist).  Never committed, never imported by the product.
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("RDP_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REF_ROOT, _REL, "dynamic_pillar_vfe.py")) and \
        os.path.isfile(os.path.join(STAGED_ROOT, _REL, "dynamic_pillar_vfe.py")):
    REF_ROOT = STAGED_ROOT
_VFE_DIR = os.path.join(REF_ROOT, _REL)
REF_FILES = ("vfe_template.py", "dynamic_pillar_vfe.py", "dynamic_voxel_vfe.py", "dynamic_mean_vfe.py")


def stage_reference(src_root: str = "/root/reference") -> bool:
    """Copies the reference files of this path into oracle/_ref/ (build container only).  Returns True if staged."""
    import shutil
    src = os.path.join(src_root, _REL)
    if not os.path.isfile(os.path.join(src, "dynamic_pillar_vfe.py")):
        return os.path.isfile(os.path.join(STAGED_ROOT, _REL, "dynamic_pillar_vfe.py"))
    dst = os.path.join(STAGED_ROOT, _REL)
    os.makedirs(dst, exist_ok=True)
    for fn in REF_FILES:
        if os.path.isfile(os.path.join(src, fn)):
            shutil.copyfile(os.path.join(src, fn), os.path.join(dst, fn))
    return True


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_VFE_DIR, "dynamic_pillar_vfe.py"))


# --------------------------------------------------------------------------- torch_scatter shim
def _scatter_mean(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    n_seg = int(index.max()) + 1 if (dim_size is None and index.numel()) else int(dim_size or 0)
    total = torch.zeros((n_seg,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    total.index_add_(0, index, src)  # CPU: sequential, in row order
    cnt = torch.zeros(n_seg, dtype=src.dtype, device=src.device)
    cnt.index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype, device=src.device))
    return total / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, n_seg):
        n, c = src.shape
        out = torch.full((n_seg, c), float("-inf"), dtype=src.dtype, device=src.device)
        idx2 = index.view(-1, 1).expand(n, c)
        out.scatter_reduce_(0, idx2, src, reduce="amax", include_self=True)
        # first (lowest) row index attaining the max -- torch_scatter CPU rule (strict '>')
        rows = torch.arange(n, device=src.device).view(-1, 1).expand(n, c)
        cand = torch.where(src == out[index], rows, torch.full_like(rows, n))
        arg = torch.full((n_seg, c), n, dtype=torch.long, device=src.device)
        arg.scatter_reduce_(0, idx2, cand, reduce="amin", include_self=True)
        out = torch.where(arg == n, torch.zeros_like(out), out)  # empty segments -> 0
        ctx.save_for_backward(arg)
        ctx.n = n
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g_out, _g_arg):
        (arg,) = ctx.saved_tensors
        n, c = ctx.n, g_out.shape[1]
        g_src = torch.zeros((n + 1, c), dtype=g_out.dtype, device=g_out.device)
        g_src.scatter_(0, arg, g_out)  # one winner per (segment, channel)
        return g_src[:n], None, None


def _scatter_max(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and src.dim() == 2
    n_seg = int(index.max()) + 1 if (dim_size is None and index.numel()) else int(dim_size or 0)
    return _ScatterMax.apply(src, index, n_seg)


def _install_torch_scatter_shim():
    mod = types.ModuleType("torch_scatter")
    mod.scatter_mean = _scatter_mean
    mod.scatter_max = _scatter_max
    mod.__rdp_shim__ = True
    sys.modules["torch_scatter"] = mod
    return mod


# --------------------------------------------------------------------------- loader
_PKG = "_rdp_ref_vfe"
_loaded = None


@contextlib.contextmanager
def cpu_cuda_identity(force: bool = False):
    """While active, ``Tensor.cuda()`` is the identity (the reference ctor calls it).  ``force``: also on a box with a
    GPU (the CPU-baseline legs of bench.py run the reference on the host cores there)."""
    if torch.cuda.is_available() and not force:
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def load_reference():
    """Returns the reference ``dynamic_pillar_vfe`` module (classes + PFNLayerV2)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise FileNotFoundError(f"reference VFE not found under {_VFE_DIR}")
    if "torch_scatter" not in sys.modules:
        _install_torch_scatter_shim()
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [_VFE_DIR]
    sys.modules[_PKG] = pkg
    for name in ("vfe_template", "dynamic_pillar_vfe", "dynamic_voxel_vfe", "dynamic_mean_vfe"):
        if not os.path.isfile(os.path.join(_VFE_DIR, f"{name}.py")):
            continue
        spec = importlib.util.spec_from_file_location(f"{_PKG}.{name}", os.path.join(_VFE_DIR, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"{_PKG}.{name}"] = m
        spec.loader.exec_module(m)
    _loaded = sys.modules[f"{_PKG}.dynamic_pillar_vfe"]
    return _loaded


class Cfg(dict):
    """Attribute-dict stand-in for the reference's EasyDict ``model_cfg``."""
    __getattr__ = dict.__getitem__


def build_reference(name, model_cfg, num_point_features, voxel_size, grid_size, point_cloud_range, force_cpu=False):
    ref = load_reference()
    if name in ("DynamicVoxelVFE", "DynMeanVFE", "DynamicMeanVFE"):
        mod = sys.modules[f"{_PKG}.dynamic_voxel_vfe" if name == "DynamicVoxelVFE" else f"{_PKG}.dynamic_mean_vfe"]
        cls = mod.DynamicVoxelVFE if name == "DynamicVoxelVFE" else mod.DynamicMeanVFE
        with cpu_cuda_identity(force_cpu):
            return cls(model_cfg=Cfg(model_cfg), num_point_features=num_point_features, voxel_size=voxel_size,
                       grid_size=grid_size, point_cloud_range=point_cloud_range)
    cls = {"DynPillarVFE": ref.DynamicPillarVFE,
           "DynamicPillarVFE": ref.DynamicPillarVFE,
           "DynamicPillarVFESimple2D": ref.DynamicPillarVFESimple2D,
           "Radar_DynamicPillarVFESimple2D": ref.Radar_DynamicPillarVFESimple2D,
           "Radar_DynamicPillarVFESimple2D_Test": ref.Radar_DynamicPillarVFESimple2D_Test}[name]
    with cpu_cuda_identity(force_cpu):
        return cls(model_cfg=Cfg(model_cfg), num_point_features=num_point_features, voxel_size=voxel_size,
                   grid_size=grid_size, point_cloud_range=point_cloud_range, depth_downsample_factor=None)


def run_reference(module, points: torch.Tensor, points_key="points", capture=None, extra=None):
    """Runs the reference forward on CPU; ``capture`` (dict) receives inverse / argmax / counts.

    The reference does not expose ``unq_inv`` / ``unq_cnt`` / the scatter_max argmax, so
    the shim functions are wrapped for the duration of the call to record them.
    """
    ts = sys.modules["torch_scatter"]
    orig_max, orig_unique = ts.scatter_max, torch.unique
    rec = {} if capture is None else capture

    def rec_max(src, index, dim=0, **kw):
        out, arg = orig_max(src, index, dim=dim, **kw)
        rec["argmax"], rec["inverse"] = arg.detach(), index.detach()
        return out, arg

    def rec_unique(*a, **k):
        r = orig_unique(*a, **k)
        if k.get("return_counts"):
            rec["unq"], rec["counts"] = r[0].detach(), r[2].detach()
        return r

    ts.scatter_max, torch.unique = rec_max, rec_unique
    try:
        with cpu_cuda_identity():
            out = module(dict({points_key: points}, **(extra or {})))
    finally:
        ts.scatter_max, torch.unique = orig_max, orig_unique
    return out

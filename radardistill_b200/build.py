"""Builds librdp.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m radardistill_b200.build [--force]

One nvcc invocation per translation unit (the PFN configurations compile in parallel), then one link.
The .so lands next to this file (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "_build")
LIB = os.environ.get("RDP_LIB_OUT", os.path.join(HERE, "librdp.so"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA = os.environ.get("RDP_EXTRA_FLAGS", "").split()
FLAGS = [*EXTRA, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "--expt-relaxed-constexpr", "-I", INCLUDE, "-I", CSRC]


def _config_ids():
    txt = open(os.path.join(CSRC, "rdp_pfn_host.h")).read()
    return sorted({int(m) for m in re.findall(r"\bX\((\d+),", txt)})


def _units():
    units = [("rdp_abi", "rdp_abi.cu", []), ("rdp_index", "rdp_index.cu", []), ("rdp_pfn", "rdp_pfn.cu", []),
             ("rdp_stack", "rdp_stack.cu", []),
             ("rdp_allreduce", "rdp_allreduce.cu", [])]
    units += [(f"rdp_pfn_inst_{i}", "rdp_pfn_inst.cu", [f"-DRDP_CFG_ID={i}"]) for i in _config_ids()]
    return units


def _source_digest() -> str:
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for fn in sorted(os.listdir(d)):
            with open(os.path.join(d, fn), "rb") as f:
                h.update(fn.encode()); h.update(f.read())
    h.update(" ".join(FLAGS + ARCH).encode())
    return h.hexdigest()


def _compile(unit, verbose):
    name, src, defs = unit
    out = os.path.join(OBJ, name + ".o")
    cmd = [NVCC, *ARCH, *FLAGS, *defs, "-c", os.path.join(CSRC, src), "-o", out]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    return out, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    units = _units()
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda u: _compile(u, verbose), units))
    if verbose:
        for (name, _, _), (_, log) in zip(units, results):
            print(f"==== {name}\n{log}")
    objs = [o for o, _ in results]
    cmd = [NVCC, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


EXT_NAME = "rdp_torch_ext"
EXT_DIR = os.path.join(HERE, "_build_ext")


def ext_path():
    """Path of the built host extension (rdp_torch_ext.so next to librdp.so), or None."""
    p = os.path.join(HERE, EXT_NAME + ".so")
    return p if os.path.exists(p) else None


def build_ext(force: bool = False, verbose: bool = False):
    """Builds the native host path (csrc/rdp_torch.cpp: a torch C++ extension, g++ only -- the kernels stay in librdp.so) in-tree,
    next to librdp.so, so that it travels to the GPU box and is imported there without a rebuild.  Compiled and linked with
    plain g++ invocations (torch's include / library paths from torch.utils.cpp_extension): `cpp_extension.load` would also
    LOAD the result, and a second load of the same TORCH_LIBRARY from its installed path aborts the process."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension
    src = os.path.join(CSRC, "rdp_torch.cpp")
    h = hashlib.sha256()
    for fn in (src, os.path.join(INCLUDE, "rdp.h")):
        with open(fn, "rb") as f:
            h.update(f.read())
    h.update(torch.__version__.encode())
    digest, stamp, out = h.hexdigest(), os.path.join(EXT_DIR, "digest.txt"), os.path.join(HERE, EXT_NAME + ".so")
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == digest:
        return out
    os.makedirs(EXT_DIR, exist_ok=True)
    cxx = cpp_extension.get_cxx_compiler()   # what cpp_extension.load would use (CXX, else the toolchain torch was configured with)
    obj = os.path.join(EXT_DIR, "rdp_torch.o")
    inc = ["-I", INCLUDE, "-I", "/usr/local/cuda/include"]
    for d in cpp_extension.include_paths() + [sysconfig.get_paths()["include"]]:
        inc += ["-isystem", d]
    compile_cmd = [cxx, f"-DTORCH_EXTENSION_NAME={EXT_NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, "-fPIC", "-std=c++17", "-O2",
                   "-c", src, "-o", obj]
    libdirs = []
    for d in cpp_extension.library_paths():
        libdirs += ["-L", d]
    link_cmd = [cxx, obj, "-shared", "-L", HERE, "-lrdp", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{HERE}", "-L", "/usr/local/cuda/lib64", "-lcudart",
                *libdirs, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-o", out]
    for cmd in (compile_cmd, link_cmd):
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building {EXT_NAME} failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_ext(force="--force" in sys.argv, verbose="-v" in sys.argv))

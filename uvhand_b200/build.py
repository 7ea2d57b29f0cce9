"""In-tree build of the two native artefacts (sm_100a only, no other arch, no fallbacks):

  uvhand_b200/lib/libmsda_b200.so                   kernels + torch-free C ABI   (nvcc)
  uvhand_b200/MultiScaleDeformableAttention*.so     PyTorch extension host layer (g++), links the above

Replaces the reference's models/ops/setup.py + make.sh (/root/reference/models/ops/setup.py:20-81),
which passes no -gencode and refuses to build without a visible GPU (setup.py:36-47); this build
cross-compiles on a GPU-less machine.  Both outputs are git-ignored and travel to the GPU box.

    python -m uvhand_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
CABI_SO = os.path.join(LIBDIR, "libmsda_b200.so")
EXT_NAME = "MultiScaleDeformableAttention"
EXT_SO = os.path.join(PKG, EXT_NAME + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _cxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def cabi_sources():
    return [os.path.join(CSRC, f) for f in ("msda_b200.cu", "msda_device.cuh", "msda_generic.cuh", "msda_d32.cuh", "msda_d32_tiles.cuh", "msda_d32_wide.cuh", "msda_d32_slab.cuh")] + [
        os.path.join(ROOT, "include", "msda_b200.h")]


def build_cabi(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _stale(CABI_SO, cabi_sources()):
        _run([_nvcc(), *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-ccbin", _cxx(),
              "-o", CABI_SO, os.path.join(CSRC, "msda_b200.cu")], verbose)
    return CABI_SO


CHECKED_SO = os.path.join(LIBDIR, "libmsda_b200_checked.so")


def build_checked(force: bool = False, verbose: bool = True) -> str:
    """The same library with -DMSDA_BOUNDS_CHECK=1: every data-dependent gather / reduction address is
    compared with the tensor's extent inside the kernels (tests run the parity suite through it; it is
    never loaded by the product path -- select it with MSDA_B200_LIB)."""
    os.makedirs(LIBDIR, exist_ok=True)
    if force or _stale(CHECKED_SO, cabi_sources()):
        _run([_nvcc(), *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-ccbin", _cxx(),
              "-DMSDA_BOUNDS_CHECK=1", "-o", CHECKED_SO, os.path.join(CSRC, "msda_b200.cu")], verbose)
    return CHECKED_SO


def build_torch_ext(force: bool = False, verbose: bool = True) -> str:
    build_cabi(force, verbose)
    src = os.path.join(CSRC, "msda_torch.cpp")
    if not (force or _stale(EXT_SO, [src, CABI_SO, os.path.join(ROOT, "include", "msda_b200.h")])):
        return EXT_SO
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]:
        inc += ["-isystem", p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = [_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", f"-DTORCH_EXTENSION_NAME={EXT_NAME}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", *inc, src, "-o", EXT_SO,
           f"-L{LIBDIR}", "-lmsda_b200", f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
           "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-Wl,-rpath,$ORIGIN/lib", f"-Wl,-rpath,{torch_lib}"]
    _run(cmd, verbose)
    return EXT_SO


def build_all(force: bool = False, verbose: bool = True):
    build_checked(force, verbose)
    return build_cabi(force, verbose), build_torch_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)

"""Builds libmaai_ntxent.so (in-tree, next to this file) with nvcc for sm_100a only.

    python multimodal-active-ai_b200/build.py [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmaai_ntxent.so")
SOURCES = ["maai_ntxent.cu"]
DEPS = ["maai_ntxent.cu", "ntxent_tile.cuh", "ntxent_aux.cuh", "ptx_sm100.cuh",
        os.path.join("..", "..", "include", "maai_ntxent.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-soname=libmaai_ntxent.so"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, defs: tuple = (), out: str | None = None) -> str:
    """defs/out: A/B variants for tuning (e.g. defs=("MAAI_POLY_FWD=4",), out="/tmp/x.so")."""
    target = out or LIB
    if not force and not out and not needs_build():
        return LIB
    tmp = target + ".tmp"  # compile aside, then rename: a concurrent reader (gpurun snapshot) never sees a torn file
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defs], "-o", tmp,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, target)
    if verbose:
        print(res.stderr)
    return target


EXT_LIB = os.path.join(HERE, "maai_torch_ext.so")
EXT_SRC = os.path.join(CSRC, "maai_torch_ext.cpp")


def ext_needs_build() -> bool:
    if not os.path.exists(EXT_LIB):
        return True
    t = os.path.getmtime(EXT_LIB)
    return any(os.path.getmtime(f) > t for f in (EXT_SRC, os.path.join(HERE, "..", "include", "maai_ntxent.h")))


def build_torch_ext(force: bool = False, verbose: bool = False) -> str:
    """maai_torch_ext.so: the C++ autograd binding (csrc/maai_torch_ext.cpp), compiled with g++ against the
    torch headers and linked to the in-tree libmaai_ntxent.so (rpath $ORIGIN).  No CUDA code of its own."""
    if not force and not ext_needs_build():
        return EXT_LIB
    if not os.path.exists(LIB):
        build()
    import shutil
    from torch.utils import cpp_extension
    bdir = os.path.join(HERE, "build", "torch_ext")
    os.makedirs(bdir, exist_ok=True)
    import ctypes
    ctypes.CDLL(LIB, mode=ctypes.RTLD_GLOBAL)  # so that load()'s trial import of the result resolves its NEEDED entry
    cpp_extension.load(
        name="maai_torch_ext", sources=[EXT_SRC], build_directory=bdir, with_cuda=True, verbose=verbose,
        extra_cflags=["-O2", "-std=c++17"], is_python_module=False,
        # ninja ($$) and the shell (quotes) both leave $ORIGIN alone: the binding finds the kernels next to it
        extra_ldflags=[f"-L{HERE}", "-l:libmaai_ntxent.so", "'-Wl,-rpath,$$ORIGIN'"])
    tmp = EXT_LIB + ".tmp"
    shutil.copyfile(os.path.join(bdir, "maai_torch_ext.so"), tmp)
    os.replace(tmp, EXT_LIB)
    return EXT_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--ext" in sys.argv:
        print(build_torch_ext(force=True, verbose=True))

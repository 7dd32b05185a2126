"""Build libtofgpu.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtofgpu.so")
SOURCES = ["tofgpu.cu"]
HEADERS = ["tof_device.cuh", "tof_kernels.cuh", "tof_common.cuh", "adv_rk4.cuh", "adv_range.cuh", "adv_planned.cuh", "adv_zrank.cuh", "simple_model.cuh",
           "simult_model.cuh", "onebd_model.cuh", "sampler.cuh", os.path.join("..", "..", "include", "tofgpu.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtofgpu.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


CHECKED_LIB_PATH = os.path.join(PKG_DIR, "libtofgpu_checked.so")


def build_checked(verbose: bool = False) -> str:
    """Same sources with -DTOF_CHECKED: every shared-memory index of the range kernels is asserted on the device
    (tof_device.cuh).  Not shipped; load it with TOFGPU_LIB=<path> to run the parity suite against it."""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DTOF_CHECKED"] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", CHECKED_LIB_PATH]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for the checked build")
    return CHECKED_LIB_PATH


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libtofgpu.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB_PATH]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(PKG_DIR, "csrc", "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB_PATH


if __name__ == "__main__":
    if "--checked" in sys.argv:
        print(build_checked(verbose=True))
    else:
        print(build_library(force="--force" in sys.argv, verbose=True))

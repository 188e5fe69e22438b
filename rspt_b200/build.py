"""In-tree build of the CUDA library and the C++ drop-in (nvcc, sm_100a only).

    python -m rspt_b200.build            # build if sources are newer than the libraries
    python -m rspt_b200.build --force
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_GPU = os.path.join(HERE, "librspt_gpu.so")
LIB_PACKER = os.path.join(HERE, "librspt_packer.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = False) -> None:
    gpu_src = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                     glob.glob(os.path.join(ROOT, "include", "*.h")))
    if force or _stale(LIB_GPU, gpu_src):
        cmd = [NVCC, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               "-o", LIB_GPU, os.path.join(CSRC, "rspt_gpu.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    cpp = os.path.join(CSRC, "signal_packer_gpu.cpp")
    if os.path.exists(cpp) and (force or _stale(LIB_PACKER, [cpp, LIB_GPU] + glob.glob(os.path.join(ROOT, "include", "*.h")))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                               "-o", LIB_PACKER, cpp, "-L", HERE, "-lrspt_gpu", "-Wl,-rpath,$ORIGIN"])


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_GPU)

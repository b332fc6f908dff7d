"""Builds librfslam_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librfslam_b200.so")
# the same sources with device-side checks of the entry lists and the DFS state (-DRF_CHECK); tests/ runs its
# time-slicing stress against it (compute-sanitizer is not available on the test boxes)
CHECK_LIB = os.path.join(HERE, "librfslam_b200_check.so")
SOURCES = ["api.cu", "dataset.cu", "grow.cu", "predict.cu", "hostpool.cu", "plugin.cu", "comm.cu", "cpiu.cu", "grow_t128.cu", "grow_t256f.cu", "grow_t320.cu", "grow_t512.cu"]
HEADERS = ["common.cuh", "internal.h", "grow_kernel.cuh", "grow_variant.h", "perf.cuh", os.path.join("..", "..", "include", "rfslam_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False, check: bool = False) -> str:
    """check=True builds CHECK_LIB (-DRF_CHECK) instead of the product library."""
    lib = CHECK_LIB if check else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".check.o" if check else ".o"))
        extra = os.environ.get("RFSLAM_NVCC_EXTRA", "").split() + (["-DRF_CHECK"] if check else [])
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed on " + src)
    cmd = [nvcc, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl", "-Xlinker", "-Bsymbolic"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, check="--check" in sys.argv))

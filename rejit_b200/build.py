"""Builds librejit_b200.so (host front end + sm_100a engine) in-tree with nvcc.

    python -m rejit_b200.build [--force]

The library is built for sm_100a only (no fallback architectures, no PTX JIT
target) and linked against the static CUDA runtime so that it can be loaded
next to any other CUDA user in the same process.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librejit_b200.so")

HOST_SOURCES = ["host/parser.cc", "host/lower.cc", "host/automaton.cc", "host/capi.cc", "host/regej.cc"]
CUDA_SOURCES = ["cuda/engine.cu"]
HEADERS = ["host/ir.h", "host/automaton.h", "cuda/engine.h", "cuda/device_program.h", "cuda/kernels.cuh", "cuda/scan_emit.cuh", "cuda/replace.cuh", "cuda/stitch.cuh",
           "../../include/rejit.h", "../../include/rejit_b200.h"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function", "-Xcompiler", "-Wno-unknown-pragmas",
              "-cudart", "static", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for rel in HOST_SOURCES + CUDA_SOURCES + HEADERS:
        p = os.path.normpath(os.path.join(CSRC, rel))
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("RJ_NVCC_EXTRA", "").split()       # tuning builds only (e.g. -DRJ_KMER_PROBE)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in HOST_SOURCES + CUDA_SOURCES]
    print("[rejit_b200.build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)

"""In-tree build of libgmix_b200.so (nvcc, sm_100a only). Used by __graft_entry__.build()."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB = os.path.join(ROOT, "lib", "libgmix_b200.so")
SOURCES = ["gmix_b200.cu"]
DEPS = ["gmix_b200.cu", "stream_kernel.cuh", "ppmd.cuh", "dmath.cuh", "spec.cuh", "layout.h", "nonstationary.inc",
        os.path.join("..", "..", "include", "gmix_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-shared", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    # exactness: no FMA contraction on the device, none on the host either
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources. Returns the .so path."""
    if not force and not _stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    print("[gmix_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)

"""In-tree build of libgmix_b200.so (nvcc, sm_100a only). Used by __graft_entry__.build()."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB = os.path.join(ROOT, "lib", "libgmix_b200.so")
SOURCES = ["host.cu", "kernel_compress.cu", "kernel_decompress.cu", "kernel_compress_prof.cu"]
DEPS = ["host.cu", "kernel_compress.cu", "kernel_decompress.cu", "kernel_compress_prof.cu", "kernels.h", "stream_kernel.cuh", "ppmd.cuh", "dmath.cuh", "spec.cuh", "layout.h", "nonstationary.inc",
        os.path.join("..", "..", "include", "gmix_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
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
    objdir = os.path.join(ROOT, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-Xptxas", "-v"] if verbose else []
    procs, objs = [], []
    for src in SOURCES:  # the translation units compile in parallel (each stream kernel takes ptxas > 1 min)
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        print("[gmix_b200] " + " ".join(cmd), file=sys.stderr)
        log = open(obj + ".log", "w")
        procs.append((subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), obj, log))
    failed = False
    for p, obj, log in procs:
        rc = p.wait()
        log.close()
        out = open(obj + ".log").read()
        if rc != 0 or verbose:
            print(out, file=sys.stderr)
        failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    print("[gmix_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)

"""In-tree build of libgmix_b200.so (nvcc, sm_100a only). Used by __graft_entry__.build()."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB = os.path.join(ROOT, "lib", "libgmix_b200.so")
SOURCES = ["host.cu", "kernel_compress.cu", "kernel_decompress.cu", "kernel_compress_prof.cu", "kernel_step.cu", "kernel_generate.cu", "kernel_genstep.cu", "kernel_gate.cu"]
DEPS = ["host.cu", "kernel_compress.cu", "kernel_decompress.cu", "kernel_compress_prof.cu", "kernel_step.cu", "kernel_generate.cu", "kernel_genstep.cu", "kernel_gate.cu", "gate_gemm.cuh", "checkpoint.h", "kernels.h", "stream_kernel.cuh", "ppmd.cuh", "dmath.cuh", "spec.cuh", "layout.h", "nonstationary.inc",
        os.path.join("..", "..", "include", "gmix_b200.h"), os.path.join("..", "host", "runner.cpp"), os.path.join("..", "host", "predictor.h"),
        os.path.join("..", "host", "coder.h"), os.path.join("..", "host", "dictionary.h"), os.path.join("..", "host", "dictionary_prep.cpp"), os.path.join("..", "host", "multi_gpu.h"), os.path.join("..", "host", "shard.h"),
        os.path.join("..", "..", "scripts", "ncu_case.cpp")]

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


def build_variant(name, defines):
    """Development: the same library compiled with extra -D flags into lib/variants/<name>/ (kernel A/B measurements;
    select it with GMIX_B200_LIB=<path>)."""
    vdir = os.path.join(ROOT, "lib", "variants", name)
    os.makedirs(vdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(vdir, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT))
    if any(p.wait() != 0 for p in procs):
        raise RuntimeError("nvcc failed")
    out = os.path.join(vdir, "libgmix_b200.so")
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs, check=True)
    for o in objs:
        os.remove(o)
    return out


def build_library(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources. Returns the .so path."""
    if not force and not _stale():
        if not os.path.exists(os.path.join(os.path.dirname(LIB), "gmixb200")):
            build_host_tools()
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objdir = os.path.join(ROOT, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("GMX_NVCC_EXTRA", "").split()
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
    build_host_tools()
    return LIB


def build_host_tools():
    """C++ host layer above the C ABI: the gmixb200 runner (reference CLI mirror) and the profiler driver."""
    libdir = os.path.dirname(LIB)
    cxx = os.environ.get("CXX", "g++")
    # host-only tool: the reference's dictionary preprocessing (host/dictionary.h), no GPU library behind it
    cmd = [cxx, "-std=c++17", "-O2", "-o", os.path.join(libdir, "dictionary-prep"), os.path.join(ROOT, "host", "dictionary_prep.cpp")]
    print("[gmix_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    for src, exe in ((os.path.join(ROOT, "host", "runner.cpp"), "gmixb200"), (os.path.join(ROOT, "..", "scripts", "ncu_case.cpp"), "ncu_case")):
        cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-o", os.path.join(libdir, exe), src, "-L" + libdir, "-lgmix_b200", "-Wl,-rpath,$ORIGIN"]
        if exe == "gmixb200":   # the multi-GPU host (host/multi_gpu.h) talks to the CUDA runtime and NCCL directly
            # NCCL: the library PyTorch ships (the one bench.py's process group runs on) when present, else the system's
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            nccl_dir = os.path.join(list(spec.submodule_search_locations)[0], "lib") if spec and spec.submodule_search_locations else None
            cmd += ["-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart", "-pthread", "-Wl,-rpath,/usr/local/cuda/lib64"]
            if nccl_dir and os.path.exists(os.path.join(nccl_dir, "libnccl.so.2")):
                cmd += [os.path.join(nccl_dir, "libnccl.so.2"), "-Wl,-rpath," + nccl_dir]
            else:
                cmd += ["-lnccl"]
        print("[gmix_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)

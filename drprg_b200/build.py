"""Builds drprg_b200/libdrprg_cuda.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libdrprg_cuda.so")
SOURCES = ["kernels.cu", "capi.cu", "ingest.cu", "prg_graph.cpp", "genotype_host.cpp"]
EXTRA = ["pandora_cuda_main.cpp"]
HEADERS = ["kernels.cuh", "prg_graph.hpp", "genotype_host.hpp", "ingest.hpp", "../../include/drprg_cuda.h"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS + EXTRA)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
           "-ccbin", host_cxx, "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-shared", "-cudart", "shared",
           "-o", SO] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lz"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    # pandora-argv-compatible front end (drprg -p/--pandora can point at it)
    exe = os.path.join(HERE, "pandora_cuda")
    subprocess.check_call([host_cxx, "-std=c++17", "-O2", "-o", exe, os.path.join(CSRC, "pandora_cuda_main.cpp"),
                           "-L" + HERE, "-ldrprg_cuda", "-Wl,-rpath,$ORIGIN"])
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)

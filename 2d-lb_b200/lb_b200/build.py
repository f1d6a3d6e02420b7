"""Builds csrc/liblb_d2q9.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG), "csrc")
LIB = os.path.join(CSRC, "liblb_d2q9.so")
SOURCES = ["lb_d2q9.cu"]
HEADERS = ["lb_device.cuh", "lb_fused.cuh", "lb_cython.cuh", "lb_oldcl.cuh", "lb_tma.cuh", "lb_tb2.cuh", "lb_tb2v.cuh", os.path.join("..", "..", "include", "lb_d2q9.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # STRICT math mirrors the reference; FAST code calls fma() explicitly
    "--shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",      # self-contained: no loader-path dependency on the GPU box; streams and
                              # device pointers are driver-level objects and interoperate with torch's runtime
]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))

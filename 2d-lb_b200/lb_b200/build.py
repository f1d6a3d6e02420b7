"""Builds csrc/liblb_d2q9.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build().

Three translation units, compiled in parallel and linked into one shared library:
  lb_d2q9.cu     C ABI, auxiliary kernels, the Cython-order / OLD-OpenCL-order kernels
  lb_k_step.cu   instantiations of the one-update kernel (register-shuffle and TMA-staged)
  lb_k_march.cu  instantiations of the two-update marching kernel
`LB_EXPERIMENTS=1` in the environment adds the round-1 tuning variants and shared-memory tiles
(-DLB_EXPERIMENTS); the default library carries the shipped kernels only.
"""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(PKG), "csrc")
LIB = os.path.join(CSRC, "liblb_d2q9.so")
SOURCES = ["lb_d2q9.cu", "lb_k_step.cu", "lb_k_march.cu"]
HEADERS = ["lb_device.cuh", "lb_f32x2.cuh", "lb_fused.cuh", "lb_march.cuh", "lb_march_rim.cuh", "lb_host.h", "lb_cython.cuh", "lb_oldcl.cuh", "lb_tma.cuh",
           "lb_tb2.cuh", "lb_tb2v.cuh", os.path.join("..", "..", "include", "lb_d2q9.h")]
OBJ_DIR = os.path.join(os.path.dirname(os.path.dirname(PKG)), "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # STRICT math mirrors the reference; FAST code calls fma() explicitly
    "-diag-suppress", "550",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "--shared",
    "-cudart", "static",      # self-contained: no loader-path dependency on the GPU box; streams and
                              # device pointers are driver-level objects and interoperate with torch's runtime
]


def _experiments():
    return os.environ.get("LB_EXPERIMENTS", "") not in ("", "0")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, extra_flags=(), out=None, tag=""):
    """Build the library.  `extra_flags` / `out` / `tag` make a side build for A/B measurements (tools/):
    extra nvcc flags, another output path, a suffix for the object files."""
    out = out or LIB
    if out == LIB and not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = (NVCC_FLAGS + (["-DLB_EXPERIMENTS"] if _experiments() else []) + (["-Xptxas", "-v"] if verbose else [])
             + list(extra_flags))

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", tag + ".o"))
        subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], cwd=CSRC, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([nvcc] + LINK_FLAGS + ["-o", out] + objs, cwd=CSRC, check=True)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))

#!/usr/bin/env python
"""TEST INFRASTRUCTURE: compile the UNMODIFIED reference Cython modules and OpenCL kernels.

Sources are read where they lie under the reference tree (never copied into this
repository); every output (generated C, objects, .so) goes to oracle/_ref/,
which is git-ignored but travels to the GPU box with the repo snapshot.

    LB_D2Q9/dimensionless/cython_dim.pyx -> oracle/_ref/cython_dim.<abi>.so
    LB_D2Q9/OLD/cython.pyx               -> oracle/_ref/old_cython/cython.<abi>.so

    LB_D2Q9/D2Q9.cl, LB_D2Q9/D2Q9i.cl    -> oracle/_ref/clshim/<digest>.so  (+ <alias>.key)
        the OpenCL C kernel files compiled as C11 through oracle/clshim/opencl_c.h and executed
        work-item by work-item on the host (oracle/shims/pyopencl is the launcher); the source is
        piped to gcc on stdin.  `pyopencl.Program.from_cache(ctx, "D2Q9")` loads the result where the
        reference tree is not mounted (the GPU box).

The reference is Python-2 era code: `language_level=2` makes Cython accept its
print statements; at import time `refload.py` supplies the skimage shim and
`np.bool`.  The reference's own setup.py is not used (it needs skimage and
writes into the source tree).
"""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

MODULES = [
    # (pyx path relative to the reference root, output sub-directory, module name)
    ("LB_D2Q9/dimensionless/cython_dim.pyx", "", "cython_dim"),
    ("LB_D2Q9/OLD/cython.pyx", "old_cython", "cython"),
]


def build(ref_root="/root/reference", verbose=False):
    import numpy as np

    if not os.path.isdir(ref_root):
        raise FileNotFoundError(f"reference tree not found at {ref_root}")
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = sysconfig.get_paths()["include"]
    built = []
    for rel, sub, name in MODULES:
        src = os.path.join(ref_root, rel)
        outdir = os.path.join(OUT, sub)
        os.makedirs(outdir, exist_ok=True)
        c_file = os.path.join(outdir, name + ".c")
        so_file = os.path.join(outdir, name + ext)
        if os.path.exists(so_file) and os.path.getmtime(so_file) >= os.path.getmtime(src):
            built.append(so_file)
            continue
        cy = [sys.executable, "-m", "cython", "-2", "-o", c_file, src]
        cc = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-fno-strict-aliasing",
              "-DNPY_NO_DEPRECATED_API=0", "-I", inc, "-I", np.get_include(),
              c_file, "-o", so_file]
        for cmd in (cy, cc):
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL,
                           stderr=None if verbose else subprocess.PIPE)
        os.remove(c_file)        # generated C embeds reference source text: do not keep it
        built.append(so_file)
    return built


CL_PROGRAMS = [
    # (kernel file relative to the reference root, alias)
    ("LB_D2Q9/D2Q9.cl", "D2Q9"),
    ("LB_D2Q9/D2Q9i.cl", "D2Q9i"),
    # the periodic-box semantic of BASELINE configs 3 and 5: float and double
    ("LB_D2Q9/rocket_yeast/rocket_yeast.cl", "rocket_yeast"),
    ("LB_D2Q9/multicomponent_multiphase/multi.cl", "multi"),
]


def build_cl(ref_root="/root/reference", verbose=False):
    """Compile the reference's OpenCL C kernel files for the CPU emulation; returns the .so paths."""
    shim = os.path.join(HERE, "shims")
    if shim not in sys.path:
        sys.path.append(shim)
    import pyopencl as cl

    if not getattr(cl, "VERSION_TEXT", "").startswith("clshim"):
        raise RuntimeError("a real pyopencl is installed; the CPU emulation is not needed")
    ctx = cl.Context()
    built = []
    for rel, alias in CL_PROGRAMS:
        with open(os.path.join(ref_root, rel)) as fh:
            prg = cl.Program(ctx, fh.read()).build()
        prg.remember_as(alias)
        built.append(os.path.join(cl.CACHE, prg._key + ".so"))
        if verbose:
            print(alias, "->", built[-1], "kernels:", prg.kernel_names)
    return built


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    for p in build(a.ref, verbose=True) + build_cl(a.ref, verbose=True):
        print("built", p)

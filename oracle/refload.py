"""TEST INFRASTRUCTURE: import the compiled reference modules from oracle/_ref/.

Used by tests/ (golden-vector generation and oracle pinning) and by bench.py's
CPU-baseline / `--impl reference` legs.  Never imported by the product path.
"""
import contextlib
import glob
import importlib.machinery
import importlib.util
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_cache = {}


def available():
    return bool(glob.glob(os.path.join(REF_DIR, "cython_dim*.so")))


def _load(name, pattern):
    if name in _cache:
        return _cache[name]
    import numpy as np

    hits = glob.glob(pattern)
    if not hits:
        raise ImportError(f"{pattern} not built; run `python oracle/build_ref.py` where /root/reference exists")
    shim = os.path.join(HERE, "shims")
    if shim not in sys.path:
        sys.path.append(shim)           # provides `skimage` only if the real one is absent
    if not hasattr(np, "bool"):
        np.bool = bool                  # cython_dim.pyx:424 uses the removed alias
    loader = importlib.machinery.ExtensionFileLoader(name, hits[0])
    spec = importlib.util.spec_from_file_location(name, hits[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _cache[name] = mod
    return mod


def cython_dim():
    """LB_D2Q9.dimensionless.cython_dim (Pipe_Flow, Pipe_Flow_Cylinder)."""
    return _load("cython_dim", os.path.join(REF_DIR, "cython_dim*.so"))


def old_cython():
    """LB_D2Q9.OLD.cython (Pipe_Flow, Pipe_Flow_Obstacles, ...)."""
    return _load("cython", os.path.join(REF_DIR, "old_cython", "cython*.so"))


@contextlib.contextmanager
def quiet():
    """The reference constructors print their parameters; swallow that."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


# ---------------------------------------------------------------------------------------------
# The reference's OpenCL host modules, executed on the CPU emulation of pyopencl
# (oracle/shims/pyopencl + oracle/clshim).  Needs the reference tree: the host files are Python 2
# (print statements), so they are read where they lie, their print statements are rewritten in
# memory, and the module is executed with __file__ pointing at the original so that it finds its
# own .cl file.  Nothing is written back or copied.
REF_ROOT = os.environ.get("LB_REFERENCE_ROOT", "/root/reference")


def opencl_host_available():
    return os.path.isfile(os.path.join(REF_ROOT, "LB_D2Q9", "dimensionless", "opencl_dim.py"))


def _py2_prints_to_py3(text):
    """`print a, b` -> `print(a, b)`; every print in the three host files is a one-line statement."""
    import re

    out = []
    for line in text.splitlines():
        m = re.match(r"^(\s*)print\s+(?!\()(.*\S)\s*$", line)
        out.append(f"{m.group(1)}print({m.group(2)})" if m else line)
    return "\n".join(out) + "\n"


class _LegacyScalarNumpy:
    """`numpy` as the host modules see it, with NumPy-1 scalar promotion restored where they rely on it.

    The host code builds `inlet_rho * np.ones(..., dtype=np.float32)` (opencl_dim.py:277) with
    `inlet_rho = 1. + np.abs(...)`, an `np.float64`.  Under the NumPy 1.x the reference was written
    for, value-based casting kept that product float32; under NumPy >= 2 (NEP 50) it silently becomes
    float64 and `cl.Buffer(hostbuf=...)` would hand float64 bytes to a `float *` kernel argument.
    NumPy 2.3 has no switch back, so scalar results of abs/sqrt/ceil/floor are returned as Python
    floats -- "weak" scalars that promote exactly like NumPy 1 did in these expressions.
    """

    _SCALAR_FUNCS = ("abs", "absolute", "sqrt", "ceil", "floor")

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        attr = getattr(self._real, name)
        if name in self._SCALAR_FUNCS:
            real = self._real

            def weak(*a, **k):
                r = attr(*a, **k)
                return float(r) if isinstance(r, real.floating) and r.ndim == 0 else r

            return weak
        return attr


def _load_host_module(name, rel):
    key = "host:" + name
    if key in _cache:
        return _cache[key]
    import types

    import numpy as np

    path = os.path.join(REF_ROOT, rel)
    if not os.path.isfile(path):
        raise ImportError(f"{path} not found (the OpenCL host modules only load where the reference is mounted)")
    shim = os.path.join(HERE, "shims")
    if shim not in sys.path:
        sys.path.append(shim)           # `pyopencl` and `skimage` stand-ins, only if the real ones are absent
    if not hasattr(np, "bool"):
        np.bool = bool
    with open(path) as fh:
        code = compile(_py2_prints_to_py3(fh.read()), path, "exec")
    mod = types.ModuleType(name)
    mod.__file__ = path
    exec(code, mod.__dict__)
    mod.np = _LegacyScalarNumpy(np)     # the classes look `np` up at call time
    _cache[key] = mod
    return mod


def opencl_dim():
    """LB_D2Q9.dimensionless.opencl_dim (Pipe_Flow, Pipe_Flow_Cylinder) on the CPU emulation."""
    return _load_host_module("ref_opencl_dim", "LB_D2Q9/dimensionless/opencl_dim.py")


def opencl_dim_D2Q9i():
    return _load_host_module("ref_opencl_dim_D2Q9i", "LB_D2Q9/dimensionless/opencl_dim_D2Q9i.py")


def old_opencl():
    """LB_D2Q9.OLD.opencl (Pipe_Flow, Pipe_Flow_Obstacles, *_PeriodicBC_VelocityInlet).

    As shipped the module opens `LB_D2Q9/OLD/D2Q9.cl` (OLD/opencl.py:153), which does not exist: the
    kernel file lives one directory up (and still holds the *_PeriodicBC_VelocityInlet kernels only
    this module launches).  Its `file_dir` global is pointed there; nothing else is touched."""
    mod = _load_host_module("ref_old_opencl", "LB_D2Q9/OLD/opencl.py")
    mod.file_dir = os.path.join(REF_ROOT, "LB_D2Q9")
    return mod

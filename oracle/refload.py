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

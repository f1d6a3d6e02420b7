"""The C-ABI library loads and exports every symbol include/lb_d2q9.h declares (no GPU needed)."""
import ctypes as ct
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lb_d2q9.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from lb_b200 import native
    names = _declared()
    assert len(names) >= 25
    lib = ct.CDLL(native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in lb_d2q9.h but not exported"
    assert sorted(native.SYMBOLS) == names, "lb_b200.native binds a different set than the header declares"


def test_library_loads_and_reports_variants():
    from lb_b200 import native
    L = native.lib()
    assert L.lb_abi_version() == native.ABI_VERSION
    v = native.variants()
    assert any(n.startswith("f32.fast.") for n in v) and any(n.startswith("f64.strict.") for n in v)
    assert ct.sizeof(native.LBConfig) == 128      # 14 x int32 + 8 x double + pointer


def test_no_device_means_loud_failure_not_fallback():
    """Without a CUDA device construction must raise; with one it must succeed.  Either way nothing
    is ever computed on the CPU."""
    import numpy as np
    import pytest
    from lb_b200 import Lattice, native
    if native.lib().lb_device_count() == 0:
        with pytest.raises(native.LBError, match="no CUDA device"):
            Lattice(16, 8, 1.0)
    else:
        with Lattice(16, 8, 1.0, dtype=np.float32) as s:
            assert s.launch_count == 0


def test_bad_configs_are_rejected_before_touching_the_device():
    import pytest
    from lb_b200 import Lattice, native
    if native.lib().lb_device_count() == 0:
        pytest.skip("argument validation order is checked on the GPU box")
    for kw in (dict(nx=1, ny=8, omega=1.0), dict(nx=16, ny=8, omega=2.5),
               dict(nx=16, ny=8, omega=1.0, bc="periodic", west_edge="boundary"),
               dict(nx=16, ny=8, omega=1.0, bc="pipe", east_edge="wrap")):
        with pytest.raises(native.LBError):
            Lattice(**kw)


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py and __graft_entry__.smoke() may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "2d-lb_b200")
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|liboracle|oracle/|oracle\.oracle|refload|d2q9_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(text), f"{os.path.join(dirpath, f)} references the oracle package"

"""ctypes binding of include/lb_d2q9.h.

The shared library is built in-tree by `lb_b200.build.build_library()` (nvcc, sm_100a).
Loading fails loudly when it is missing -- there is no Python or CPU substitute for it.
"""
import ctypes as ct
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# LB_D2Q9_LIB: another build of the same library (A/B measurements, tools/); the product path never sets it
LIB_PATH = os.environ.get("LB_D2Q9_LIB") or os.path.join(os.path.dirname(_PKG), "csrc", "liblb_d2q9.so")

# enums of lb_d2q9.h
F32, F64 = 0, 1
BC_PIPE, BC_PERIODIC, BC_VELOCITY_YPERIODIC = 0, 1, 2
MATH_STRICT, MATH_FAST = 0, 1
FIELD_F, FIELD_FEQ, FIELD_RHO, FIELD_U, FIELD_V = 0, 1, 2, 3, 4
WEST, EAST = 0, 1
EDGE_BOUNDARY, EDGE_WRAP, EDGE_HALO = 0, 1, 2
SYNTH_PIPE_RAMP, SYNTH_SHEAR_LAYERS = 0, 1
IPC_HANDLE_BYTES = 64
ABI_VERSION = 5
SCHEME_OPENCL, SCHEME_CYTHON, SCHEME_CYTHON_OLD, SCHEME_OPENCL_OLD = 0, 1, 2, 3
MODEL_D2Q9, MODEL_D2Q9I = 0, 1

# every symbol include/lb_d2q9.h declares (tests/test_abi.py checks the library exports them all)
SYMBOLS = [
    "lb_abi_version", "lb_device_count", "lb_create", "lb_destroy", "lb_last_error",
    "lb_set_mask", "lb_upload_f", "lb_upload_moments", "lb_step", "lb_run_streamed", "lb_sync", "lb_download", "lb_download_strided",
    "lb_stage_move", "lb_stage_move_bcs", "lb_stage_update_hydro", "lb_stage_update_feq",
    "lb_stage_collide", "lb_stage_zero_velocity", "lb_init_synthetic", "lb_set_mask_disk",
    "lb_total_mass", "lb_checksum", "lb_selftest_rcp", "lb_selftest_copy", "lb_plan_march_launch", "lb_set_temporal_blocking", "lb_temporal_blocking", "lb_segment_rows", "lb_tb2_shape_count", "lb_tb2_shape_name", "lb_launch_count", "lb_set_variant", "lb_variant_count", "lb_variant_name",
    "lb_device_ptr", "lb_stream", "lb_halo_ipc_handle", "lb_halo_connect_ipc", "lb_halo_connect_local",
    "lb_halo_prime", "lb_set_halo_timeout",
    "lb_multi_create", "lb_multi_destroy", "lb_multi_last_error", "lb_multi_slab_count", "lb_multi_slab",
    "lb_multi_set_mask", "lb_multi_set_mask_disk", "lb_multi_upload_f", "lb_multi_upload_moments",
    "lb_multi_init_synthetic", "lb_multi_set_temporal_blocking", "lb_multi_temporal_blocking", "lb_multi_prime",
    "lb_multi_step", "lb_multi_sync", "lb_multi_download", "lb_multi_total_mass", "lb_multi_checksum",
    "lb_multi_launch_count",
]


class LBError(RuntimeError):
    """Raised for any non-zero status of the C library (message from lb_last_error)."""

    def __init__(self, code, message):
        super().__init__(f"lb_d2q9 error {code}: {message}")
        self.code = code


class LBConfig(ct.Structure):
    _fields_ = [
        ("struct_size", ct.c_int32), ("device", ct.c_int32),
        ("nx", ct.c_int32), ("ny", ct.c_int32),
        ("dtype", ct.c_int32), ("bc", ct.c_int32), ("math", ct.c_int32),
        ("zero_obstacle_velocity", ct.c_int32),
        ("global_nx", ct.c_int32), ("x_offset", ct.c_int32),
        ("west_edge", ct.c_int32), ("east_edge", ct.c_int32),
        ("scheme", ct.c_int32), ("model", ct.c_int32),
        ("omega", ct.c_double), ("inlet_rho", ct.c_double), ("outlet_rho", ct.c_double),
        ("cs2", ct.c_double), ("cs22", ct.c_double), ("two_cs4", ct.c_double),
        ("u_west", ct.c_double), ("u_east", ct.c_double),
        ("stream", ct.c_void_p),
    ]


_lib = None


def _declare(lib):
    vp, i, d = ct.c_void_p, ct.c_int, ct.c_double
    sig = {
        "lb_abi_version": (i, []),
        "lb_device_count": (i, []),
        "lb_create": (i, [ct.POINTER(LBConfig), ct.POINTER(vp)]),
        "lb_destroy": (i, [vp]),
        "lb_last_error": (ct.c_char_p, [vp]),
        "lb_set_mask": (i, [vp, vp, i]),
        "lb_upload_f": (i, [vp, vp]),
        "lb_upload_moments": (i, [vp, vp, vp, vp]),
        "lb_step": (i, [vp, i]),
        "lb_run_streamed": (i, [vp, vp, i, vp, vp, vp]),
        "lb_sync": (i, [vp]),
        "lb_download": (i, [vp, i, vp]),
        "lb_download_strided": (i, [vp, i, i, i, vp]),
        "lb_stage_move": (i, [vp]),
        "lb_stage_move_bcs": (i, [vp]),
        "lb_stage_update_hydro": (i, [vp]),
        "lb_stage_update_feq": (i, [vp]),
        "lb_stage_collide": (i, [vp]),
        "lb_stage_zero_velocity": (i, [vp]),
        "lb_init_synthetic": (i, [vp, i, d, d, ct.c_uint64]),
        "lb_set_mask_disk": (i, [vp, d, d, d]),
        "lb_total_mass": (i, [vp, ct.POINTER(d)]),
        "lb_checksum": (i, [vp, ct.POINTER(ct.c_uint64)]),
        "lb_selftest_rcp": (i, [i, ct.c_uint32, ct.c_uint32, ct.POINTER(ct.c_uint64)]),
        "lb_selftest_copy": (i, [vp, i, ct.POINTER(ct.c_double)]),
        "lb_plan_march_launch": (i, [i] * 11 + [ct.POINTER(i)] * 5),
        "lb_set_temporal_blocking": (i, [vp, i]),
        "lb_temporal_blocking": (i, [vp]),
        "lb_segment_rows": (i, [vp]),
        "lb_tb2_shape_count": (i, []),
        "lb_tb2_shape_name": (ct.c_char_p, [i]),
        "lb_launch_count": (ct.c_int64, [vp]),
        "lb_set_variant": (i, [vp, i]),
        "lb_variant_count": (i, []),
        "lb_variant_name": (ct.c_char_p, [i]),
        "lb_device_ptr": (i, [vp, i, ct.POINTER(vp), ct.POINTER(ct.c_int64)]),
        "lb_stream": (vp, [vp]),
        "lb_halo_ipc_handle": (i, [vp, vp]),
        "lb_halo_connect_ipc": (i, [vp, i, vp, i]),
        "lb_halo_connect_local": (i, [vp, i, vp]),
        "lb_halo_prime": (i, [vp]),
        "lb_set_halo_timeout": (i, [vp, d]),
        "lb_multi_create": (i, [ct.POINTER(LBConfig), i, ct.POINTER(ct.c_int), ct.POINTER(vp)]),
        "lb_multi_destroy": (i, [vp]),
        "lb_multi_last_error": (ct.c_char_p, [vp]),
        "lb_multi_slab_count": (i, [vp]),
        "lb_multi_slab": (i, [vp, i, ct.POINTER(vp), ct.POINTER(ct.c_int), ct.POINTER(ct.c_int)]),
        "lb_multi_set_mask": (i, [vp, vp, i]),
        "lb_multi_set_mask_disk": (i, [vp, d, d, d]),
        "lb_multi_upload_f": (i, [vp, vp]),
        "lb_multi_upload_moments": (i, [vp, vp, vp, vp]),
        "lb_multi_init_synthetic": (i, [vp, i, d, d, ct.c_uint64]),
        "lb_multi_set_temporal_blocking": (i, [vp, i]),
        "lb_multi_temporal_blocking": (i, [vp]),
        "lb_multi_prime": (i, [vp]),
        "lb_multi_step": (i, [vp, i]),
        "lb_multi_sync": (i, [vp]),
        "lb_multi_download": (i, [vp, i, vp]),
        "lb_multi_total_mass": (i, [vp, ct.POINTER(d)]),
        "lb_multi_checksum": (i, [vp, ct.POINTER(ct.c_uint64)]),
        "lb_multi_launch_count": (ct.c_int64, [vp]),
    }
    assert sorted(sig) == sorted(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def lib():
    """The loaded C library.  Raises (never substitutes) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  lb_b200 has no CPU fallback.")
        handle = ct.CDLL(LIB_PATH)
        _declare(handle)
        if handle.lb_abi_version() != ABI_VERSION:
            raise ImportError("liblb_d2q9.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(status, handle=None):
    if status != 0:
        msg = lib().lb_last_error(handle)
        raise LBError(status, msg.decode() if msg else "unknown error")


def check_multi(status, handle=None):
    if status != 0:
        msg = lib().lb_multi_last_error(handle)
        raise LBError(status, msg.decode() if msg else "unknown error")


def variants():
    """Names of the compiled tile configurations of the fused kernel, by index."""
    L = lib()
    return [L.lb_variant_name(k).decode() for k in range(L.lb_variant_count())]

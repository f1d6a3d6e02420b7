"""TEST INFRASTRUCTURE -- a CPU stand-in for the slice of `pyopencl` the reference's host code uses.

The reference's OpenCL path (`LB_D2Q9/dimensionless/opencl_dim.py` and friends) cannot run in this
image: there is no pyopencl and no OpenCL ICD.  This package lets the *unmodified* host classes run
anyway: `Program(ctx, src).build()` hands the OpenCL C source -- untouched, piped to gcc on stdin --
to `gcc -x c -include oracle/clshim/opencl_c.h`, and a kernel launch executes the NDRange on the
host, work-group by work-group (oracle/clshim/ndrange.c; barrier() is honoured with fibers).

What is emulated: get_platforms / Context / CommandQueue / Program.build / kernel launch with
(queue, global_size, local_size, *args) / Buffer(COPY_HOST_PTR) / LocalMemory / enqueue_copy /
Event.wait -- the call sites of opencl_dim.py:165-176, 203-255, 291-407.  Arithmetic: C11 with
-ffp-contract=off, i.e. OpenCL C's own promotion rules without fused multiply-adds.

Only tests/ (and the golden-vector generator) import this; the product path never does.
"""
import ctypes as ct
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(os.path.dirname(HERE))
CLSHIM = os.path.join(ORACLE, "clshim")
CACHE = os.path.join(ORACLE, "_ref", "clshim")
CFLAGS = ["-O2", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing", "-w"]

PARSER_VERSION = "2"          # part of the cache key: bump when parse_kernels / _trampolines change
VERSION = (2015, 1)
VERSION_TEXT = "clshim (gcc CPU emulation)"


class Error(Exception):
    pass


class RuntimeError(Error):              # noqa: A001 - pyopencl's own name
    pass


class LogicError(Error):
    pass


class MemoryError(Error):               # noqa: A001
    pass


# --------------------------------------------------------------------------- enumerations
class mem_flags:
    READ_WRITE = 1 << 0
    WRITE_ONLY = 1 << 1
    READ_ONLY = 1 << 2
    USE_HOST_PTR = 1 << 3
    ALLOC_HOST_PTR = 1 << 4
    COPY_HOST_PTR = 1 << 5


class device_type:
    DEFAULT = 1
    CPU = 2
    GPU = 4
    ACCELERATOR = 8
    ALL = 0xFFFFFFFF

    @staticmethod
    def to_string(value):
        return {1: "DEFAULT", 2: "CPU", 4: "GPU", 8: "ACCELERATOR"}.get(value, str(value))


class command_queue_properties:
    OUT_OF_ORDER_EXEC_MODE_ENABLE = 1
    PROFILING_ENABLE = 2


class context_properties:
    PLATFORM = 0x1084


# --------------------------------------------------------------------------- platform objects
class Device:
    name = "host CPU through gcc (clshim)"
    type = device_type.CPU
    vendor = "oracle/clshim"
    max_clock_frequency = 0
    max_mem_alloc_size = 1 << 34
    max_work_group_size = 4096
    max_work_item_dimensions = 3
    max_work_item_sizes = [4096, 4096, 4096]
    double_fp_config = 63


class Platform:
    name = "clshim"
    vendor = "oracle/clshim (test infrastructure)"
    version = "OpenCL C 1.2 subset compiled as C11"

    def __init__(self):
        self._devices = [Device()]

    def get_devices(self, device_type=device_type.ALL):
        return list(self._devices)


_PLATFORM = Platform()


def get_platforms():
    return [_PLATFORM]


class Context:
    def __init__(self, devices=None, properties=None, dev_type=None):
        self.devices = list(devices) if devices else _PLATFORM.get_devices()
        self.properties = properties


def create_some_context(interactive=False):
    return Context()


class CommandQueue:
    def __init__(self, context, device=None, properties=None):
        self.context = context
        self.device = device or context.devices[0]
        self.properties = properties

    def finish(self):
        pass

    def flush(self):
        pass


class _Profile:
    start = end = queued = submit = 0


class Event:
    profile = _Profile()

    def wait(self):
        return self


# --------------------------------------------------------------------------- memory objects
class Buffer:
    """Device buffer = an owned byte array.  COPY_HOST_PTR snapshots the host array's memory as it
    lies (Fortran-ordered arrays stay Fortran-ordered), like clCreateBuffer does."""

    def __init__(self, context, flags, size=0, hostbuf=None):
        if hostbuf is not None:
            host = np.asarray(hostbuf)
            if not (host.flags.c_contiguous or host.flags.f_contiguous):
                raise LogicError("hostbuf must be contiguous")
            raw = np.frombuffer(host.tobytes(order="A"), dtype=np.uint8)
            self.size = raw.size
            self._mem = np.empty(max(self.size, 1), dtype=np.uint8)
            self._mem[: self.size] = raw
        else:
            if size <= 0:
                raise LogicError("Buffer needs a size or a hostbuf")
            self.size = int(size)
            self._mem = np.zeros(self.size, dtype=np.uint8)
        self.flags = flags
        self.context = context

    @property
    def ptr(self):
        return self._mem.ctypes.data

    def release(self):
        pass


class LocalMemory:
    def __init__(self, size):
        self.size = int(size)


def _host_bytes_view(arr):
    """Flat uint8 view of a contiguous host array's memory (either order)."""
    a = np.asarray(arr)
    if a.flags.c_contiguous:
        flat = a.reshape(-1)
    elif a.flags.f_contiguous:
        flat = a.reshape(-1, order="F")
    else:
        raise LogicError("host array must be contiguous")
    return flat.view(np.uint8)


def enqueue_copy(queue, dest, src, is_blocking=True, wait_for=None, **kwargs):
    if isinstance(dest, Buffer) and isinstance(src, Buffer):
        n = min(dest.size, src.size)
        dest._mem[:n] = src._mem[:n]
    elif isinstance(src, Buffer):
        view = _host_bytes_view(dest)
        if view.size > src.size:
            raise LogicError("host array larger than the buffer")
        view[...] = src._mem[: view.size]
    elif isinstance(dest, Buffer):
        view = _host_bytes_view(src)
        if view.size > dest.size:
            raise LogicError("buffer smaller than the host array")
        dest._mem[: view.size] = view
    else:
        raise LogicError("enqueue_copy needs at least one Buffer")
    return Event()


def enqueue_barrier(queue, wait_for=None):
    return Event()


# --------------------------------------------------------------------------- programs and kernels
_QUALIFIERS = {"__global", "__local", "__constant", "__private", "__read_only", "__write_only", "__read_write",
               "global", "local", "constant", "const", "restrict", "volatile"}
_SCALARS = {
    "float": (ct.c_float, np.float32), "double": (ct.c_double, np.float64),
    "int": (ct.c_int32, np.int32), "uint": (ct.c_uint32, np.uint32), "unsigned int": (ct.c_uint32, np.uint32),
    "long": (ct.c_int64, np.int64), "ulong": (ct.c_uint64, np.uint64),
    "short": (ct.c_int16, np.int16), "ushort": (ct.c_uint16, np.uint16),
    "char": (ct.c_int8, np.int8), "uchar": (ct.c_uint8, np.uint8),
}
_KERNEL_RE = re.compile(r"\b(?:__kernel|kernel)\s+void\s+(\w+)\s*\(([^)]*)\)\s*\{", re.S)


def _strip_comments(src):
    """One left-to-right pass, so that `//**** title ****` is a line comment, not the start of a block."""
    return re.sub(r"//[^\n]*|/\*.*?\*/", " ", src, flags=re.S)


def _matching_brace(text, open_pos):
    depth = 0
    for i in range(open_pos, len(text)):
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                return i
    return len(text)


def parse_kernels(src):
    """[{name, params: [{kind: 'global'|'local'|'scalar', ctype, text}], uses_barrier}] from OpenCL C."""
    clean = _strip_comments(src)
    kernels = []
    for m in _KERNEL_RE.finditer(clean):
        name, plist = m.group(1), m.group(2)
        body = clean[m.end() - 1: _matching_brace(clean, m.end() - 1)]
        params = []
        for raw in [p.strip() for p in plist.split(",") if p.strip()]:
            words = raw.replace("*", " * ").split()
            is_ptr = "*" in words
            is_local = "__local" in words or "local" in words
            base = [w for w in words[:-1] if w not in _QUALIFIERS and w != "*"]
            ctype = " ".join(base)
            if is_ptr:
                kind = "local" if is_local else "global"
            else:
                if ctype not in _SCALARS:
                    raise LogicError(f"kernel {name}: unsupported scalar parameter type {ctype!r}")
                kind = "scalar"
            params.append({"kind": kind, "ctype": ctype, "text": raw})
        kernels.append({"name": name, "params": params, "uses_barrier": bool(re.search(r"\bbarrier\s*\(", body))})
    return kernels


def _trampolines(kernels):
    """C source of one `launch_<kernel>` per kernel: packs the arguments, runs the NDRange."""
    out = ['#include "opencl_c.h"\n']
    for k in kernels:
        n = k["name"]
        decl = ", ".join(p["text"] for p in k["params"]) or "void"
        out.append(f"void {n}({decl});\n")
        fields, unpack, sig, fill = [], [], [], []
        for i, p in enumerate(k["params"]):
            t = "void *" if p["kind"] != "scalar" else p["ctype"] + " "
            fields.append(f"    {t}a{i};\n")
            unpack.append(f"a->a{i}")
            sig.append(f"{t}a{i}")
            fill.append(f"    a.a{i} = a{i};\n")
        out.append(f"struct args_{n} {{\n{''.join(fields) or '    int unused;\n'}}};\n")
        out.append(f"static void thunk_{n}(void *p) {{ struct args_{n} *a = (struct args_{n} *)p; (void)a; "
                   f"{n}({', '.join(unpack)}); }}\n")
        head = "unsigned dim, const size_t *g, const size_t *l" + ("".join(", " + s for s in sig))
        out.append(f"int launch_{n}({head}) {{\n    struct args_{n} a;\n{''.join(fill)}"
                   f"    return clshim_run(dim, g, l, thunk_{n}, &a, {int(k['uses_barrier'])});\n}}\n")
    return "".join(out)


def _support_digest():
    h = hashlib.sha256()
    for fn in ("opencl_c.h", "ndrange.c"):
        with open(os.path.join(CLSHIM, fn), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(CFLAGS).encode())
    h.update(PARSER_VERSION.encode())
    return h


class _Kernel:
    def __init__(self, lib, meta):
        self._meta = meta
        self.function_name = meta["name"]
        self._fn = getattr(lib, "launch_" + meta["name"])
        argtypes = [ct.c_uint, ct.POINTER(ct.c_size_t), ct.POINTER(ct.c_size_t)]
        for p in meta["params"]:
            argtypes.append(ct.c_void_p if p["kind"] != "scalar" else _SCALARS[p["ctype"]][0])
        self._fn.argtypes = argtypes
        self._fn.restype = ct.c_int

    def __call__(self, queue, global_size, local_size, *args, **kwargs):
        meta = self._meta
        if len(args) != len(meta["params"]):
            raise LogicError(f"{meta['name']}: expected {len(meta['params'])} arguments, got {len(args)}")
        dim = len(global_size)
        if local_size is None:
            local_size = (1,) * dim
        if len(local_size) != dim:
            raise LogicError("global and local size must have the same rank")
        g = (ct.c_size_t * 3)(*(list(global_size) + [1] * (3 - dim)))
        loc = (ct.c_size_t * 3)(*(list(local_size) + [1] * (3 - dim)))
        cargs, keep = [], []
        for p, a in zip(meta["params"], args):
            if p["kind"] == "global":
                if not isinstance(a, Buffer):
                    raise LogicError(f"{meta['name']}: parameter `{p['text']}` needs a Buffer")
                cargs.append(a.ptr)
            elif p["kind"] == "local":
                if not isinstance(a, LocalMemory):
                    raise LogicError(f"{meta['name']}: parameter `{p['text']}` needs LocalMemory")
                scratch = np.zeros(max(a.size, 1), dtype=np.uint8)
                keep.append(scratch)
                cargs.append(scratch.ctypes.data)
            else:
                ctype, npt = _SCALARS[p["ctype"]]
                if not isinstance(a, np.generic):
                    raise LogicError(f"{meta['name']}: scalar `{p['text']}` must be a sized numpy scalar")
                if a.dtype.itemsize != np.dtype(npt).itemsize:
                    raise LogicError(f"{meta['name']}: `{p['text']}` got a {a.dtype} ({a.dtype.itemsize} bytes)")
                cargs.append(ctype(a.item()))
        rc = self._fn(dim, g, loc, *cargs)
        if rc:
            raise RuntimeError(f"{meta['name']}: NDRange launch failed ({rc}); global {tuple(global_size)} "
                               f"local {tuple(local_size)}")
        return Event()


class Program:
    def __init__(self, context, src=None):
        self.context = context
        self._src = src
        self._lib = None
        self._kernels = {}

    # -- cache of compiled programs, keyed by the source text --------------------------------------
    @staticmethod
    def key_of(src):
        h = _support_digest()
        h.update(src.encode())
        return h.hexdigest()[:20]

    def build(self, options="", devices=None, cache_dir=None):
        key = self.key_of(self._src)
        return self._load(key, self._src)

    @classmethod
    def from_cache(cls, context, alias):
        """Load a program built earlier under `alias` (oracle/build_ref.py) -- for places where the
        reference sources are not mounted (the GPU box)."""
        path = os.path.join(CACHE, alias + ".key")
        if not os.path.exists(path):
            raise Error(f"no cached program {alias!r} under {CACHE}; run oracle/build_ref.py where the reference is mounted")
        with open(path) as fh:
            key = fh.read().strip()
        return cls(context)._load(key, None)

    def remember_as(self, alias):
        os.makedirs(CACHE, exist_ok=True)
        with open(os.path.join(CACHE, alias + ".key"), "w") as fh:
            fh.write(self._key + "\n")

    def _load(self, key, src):
        so = os.path.join(CACHE, key + ".so")
        meta_path = os.path.join(CACHE, key + ".json")
        if not (os.path.exists(so) and os.path.exists(meta_path)):
            if src is None:
                raise Error(f"cached program {key} is missing and no source was given")
            self._compile(src, so, meta_path)
        with open(meta_path) as fh:
            kernels = json.load(fh)
        self._key = key
        self._lib = ct.CDLL(so)
        self._kernels = {k["name"]: _Kernel(self._lib, k) for k in kernels}
        return self

    @staticmethod
    def _compile(src, so, meta_path):
        os.makedirs(CACHE, exist_ok=True)
        kernels = parse_kernels(src)
        if not kernels:
            raise RuntimeError("no __kernel found in the program source")
        tmp = so + f".{os.getpid()}"
        obj_k, obj_t, obj_n = tmp + ".k.o", tmp + ".t.o", tmp + ".n.o"
        inc = ["-I", CLSHIM, "-include", os.path.join(CLSHIM, "opencl_c.h")]
        try:
            # the kernel source goes to gcc on stdin: no copy of it is written anywhere
            for cmd, text in ((["gcc", *CFLAGS, *inc, "-x", "c", "-c", "-", "-o", obj_k], src),
                              (["gcc", *CFLAGS, "-I", CLSHIM, "-x", "c", "-c", "-", "-o", obj_t], _trampolines(kernels))):
                r = subprocess.run(cmd, input=text.encode(), capture_output=True)
                if r.returncode:
                    raise RuntimeError("clBuildProgram failed:\n" + r.stderr.decode(errors="replace"))
            r = subprocess.run(["gcc", *CFLAGS, "-I", CLSHIM, "-c", os.path.join(CLSHIM, "ndrange.c"), "-o", obj_n],
                               capture_output=True)
            if r.returncode:
                raise RuntimeError("ndrange.c failed to compile:\n" + r.stderr.decode(errors="replace"))
            r = subprocess.run(["gcc", "-shared", "-o", tmp, obj_k, obj_t, obj_n, "-lm"], capture_output=True)
            if r.returncode:
                raise RuntimeError("link failed:\n" + r.stderr.decode(errors="replace"))
            os.replace(tmp, so)
            with open(meta_path, "w") as fh:
                json.dump(kernels, fh)
        finally:
            for p in (obj_k, obj_t, obj_n, tmp):
                if os.path.exists(p):
                    os.remove(p)

    @property
    def kernel_names(self):
        return ";".join(self._kernels)

    def all_kernels(self):
        return list(self._kernels.values())

    def __getattr__(self, name):
        kernels = self.__dict__.get("_kernels") or {}
        if name in kernels:
            return kernels[name]
        raise AttributeError(name)


from . import tools  # noqa: E402,F401  (the reference does `import pyopencl.tools`)

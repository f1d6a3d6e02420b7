"""TEST INFRASTRUCTURE -- `pyopencl.tools` stand-in (see the package docstring)."""


def get_gl_sharing_context_properties():
    """No GL interop on the CPU emulation; `use_interop=True` paths are out of scope."""
    return []

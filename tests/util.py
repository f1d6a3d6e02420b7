"""Shared helpers for the parity tests: seeded inputs in device layout ([9][ny][nx])."""
import numpy as np


def pipe_case(orc, nx, ny, dtype=np.float32, inlet_rho=1.01, outlet_rho=1.0, seed=0, amplitude=1e-3,
              mask="none"):
    """Initial populations the way opencl_dim builds them (opencl_dim.py:279-288, :315-321):
    density ramp, u=v=0, f = feq * (1 + amplitude*randn)."""
    rng = np.random.RandomState(seed)
    x = np.arange(nx)
    rho = (inlet_rho - x * (inlet_rho - outlet_rho) / float(nx)).astype(np.float32)[None, :].repeat(ny, 0)
    zero = np.zeros((ny, nx))
    feq = orc.feq_of(rho, zero, zero, dtype)
    f0 = (feq * (1.0 + amplitude * rng.randn(9, ny, nx))).astype(dtype)
    m = None
    if mask == "blocks":
        m = np.zeros((ny, nx), np.uint8)
        m[ny // 3: ny // 3 + max(2, ny // 6), nx // 4: nx // 4 + max(2, nx // 10)] = 1
        m[2 * ny // 3: 2 * ny // 3 + 2, nx // 2: nx // 2 + 3] = 1
    elif mask == "random":
        m = (rng.rand(ny, nx) < 0.05).astype(np.uint8)
        m[0, :] = m[-1, :] = 0
        m[:, 0] = m[:, -1] = 0
    elif mask == "bulky":             # a body wider than a warp's span: whole groups of 32 / 64 / 128 cells are solid
        m = np.zeros((ny, nx), np.uint8)
        m[ny // 5: ny - ny // 5, nx // 8: nx - nx // 6] = 1
        m[ny // 2, nx // 3: nx // 3 + 40] = 0                      # a slit: mixed groups inside the body
        m[1:3, :70] = 1                                            # and one hugging the inlet corner
    elif mask == "touching":          # solids on the walls, inlet and outlet columns and corners
        m = (rng.rand(ny, nx) < 0.05).astype(np.uint8)
        m[0, :3] = 1
        m[-1, -3:] = 1
        m[ny // 2, 0] = 1
        m[ny // 2, -1] = 1
    return f0, m


def periodic_case(orc, nx, ny, dtype=np.float32, u0=0.05, seed=0, amplitude=0.0):
    """Doubly periodic shear layers."""
    rng = np.random.RandomState(seed)
    yy = (np.arange(ny) / float(ny))[:, None].repeat(nx, 1)
    xx = (np.arange(nx) / float(nx))[None, :].repeat(ny, 0)
    u = np.where(yy < 0.5, u0 * np.tanh(80 * (yy - 0.25)), u0 * np.tanh(80 * (0.75 - yy)))
    v = 0.05 * u0 * np.sin(2 * np.pi * (xx + 0.25))
    f0 = orc.feq_of(np.ones((ny, nx)), u, v, dtype)
    if amplitude:
        f0 = (f0 * (1.0 + amplitude * rng.randn(9, ny, nx))).astype(dtype)
    return f0


def rel_err(a, b):
    """max|a-b| / max|b| (the tolerance form of BASELINE.json's north_star)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

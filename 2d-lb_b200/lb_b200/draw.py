"""Disk rasterisation used for the cylinder mask.

The reference calls `skimage.draw.circle(x_c, y_c, N)` (opencl_dim.py:474), a third-party
routine it does not pin (setup.py:29) and that newer scikit-image releases removed.  This
is its published contract (scikit-image <= 0.18): pixels (i, j) of the bounding box with
((i - r)/R)^2 + ((j - c)/R)^2 < 1.  The mask is an input of the hot path (shared by the CUDA
kernel and the oracle), so parity never depends on this routine.
"""
import numpy as np


def circle(r, c, radius, shape=None):
    center = np.array([r, c], dtype=float)
    radii = np.array([radius, radius], dtype=float)
    lo = np.ceil(center - radii).astype(int)
    hi = np.floor(center + radii).astype(int)
    if shape is not None:
        lo = np.maximum(lo, 0)
        hi = np.minimum(hi, np.array(shape[:2]) - 1)
    ii, jj = np.ogrid[0:float(hi[0] - lo[0] + 1), 0:float(hi[1] - lo[1] + 1)]
    rel = center - lo
    inside = ((ii - rel[0]) / radii[0]) ** 2 + ((jj - rel[1]) / radii[1]) ** 2 < 1
    rr, cc = np.nonzero(inside)
    return rr + lo[0], cc + lo[1]

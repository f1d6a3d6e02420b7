"""`skimage.draw.circle(r, c, radius)` as scikit-image <= 0.18 defined it.

Contract (not reference code; the reference pins no version, setup.py:29):
pixels of the bounding box [ceil(center - R), floor(center + R)] whose centre
satisfies ((i - r)/R)**2 + ((j - c)/R)**2 < 1.  Mask pixels are an INPUT shared
by the oracle and the CUDA path, so parity does not depend on this shim.
"""
import numpy as np


def circle(r, c, radius, shape=None):
    center = np.array([r, c], dtype=float)
    radii = np.array([radius, radius], dtype=float)
    upper_left = np.ceil(center - radii).astype(int)
    lower_right = np.floor(center + radii).astype(int)
    if shape is not None:
        upper_left = np.maximum(upper_left, np.array([0, 0]))
        lower_right = np.minimum(lower_right, np.array(shape[:2]) - 1)
    shifted = center - upper_left
    bshape = lower_right - upper_left + 1
    ii, jj = np.ogrid[0:float(bshape[0]), 0:float(bshape[1])]
    d = ((ii - shifted[0]) / radii[0]) ** 2 + ((jj - shifted[1]) / radii[1]) ** 2
    rr, cc = np.nonzero(d < 1)
    return rr + upper_left[0], cc + upper_left[1]

"""Obstacle-mask ingestion: image -> boolean (nx, ny) mask for Pipe_Flow_Obstacles.

The reference builds such masks in a notebook (docs/cs205_movie.ipynb cells 11-16: tifffile.imread,
skimage.transform.resize, threshold) and injects them with the `obstacle_mask_host` hack.  Here:
  from_image(path, nx, ny)    any image PIL can read (the reference's TIFFs included) -> (nx, ny) bool
  resample(src, nx, ny)       nearest-neighbour resampling  mask[x, y] = src[x*W//nx, y*H//ny]
  pack(mask) / unpack(blob)   1 bit per node, for fixtures
Masks use the reference's host convention: shape (nx, ny), True/1 = solid.
"""
import numpy as np


def resample(src, nx, ny):
    """Nearest-neighbour resampling of an (W, H) mask to (nx, ny): mask[x, y] = src[x*W//nx, y*H//ny]."""
    src = np.asarray(src)
    w, h = src.shape
    xi = (np.arange(nx, dtype=np.int64) * w) // nx
    yi = (np.arange(ny, dtype=np.int64) * h) // ny
    return np.ascontiguousarray(src[np.ix_(xi, yi)])


def from_image(path, nx=None, ny=None, threshold=0.5, solid_is_dark=False):
    """Read an image (PIL), convert to grey, threshold, transpose to (x, y) and resample to (nx, ny)."""
    from PIL import Image
    img = np.asarray(Image.open(path).convert("L"), dtype=np.float64) / 255.0      # (H, W), row 0 = top
    solid = (img < threshold) if solid_is_dark else (img >= threshold)
    mask = np.ascontiguousarray(solid.T)                                            # (W, H) = (x, y)
    if nx is not None and ny is not None:
        mask = resample(mask, nx, ny)
    return mask


def pack(mask):
    mask = np.asarray(mask, dtype=bool)
    return dict(bits=np.packbits(mask.ravel()), shape=np.array(mask.shape))


def unpack(blob):
    shape = tuple(int(v) for v in blob["shape"])
    n = int(np.prod(shape))
    return np.unpackbits(np.asarray(blob["bits"]))[:n].astype(bool).reshape(shape)

"""Obstacle-mask ingestion: image -> boolean (nx, ny) mask for Pipe_Flow_Obstacles.

The reference builds such masks in a notebook (docs/cs205_movie.ipynb cells 11-16: tifffile.imread,
skimage.transform.resize, threshold) and injects them with the `obstacle_mask_host` hack.  Here:
  from_image(path, nx, ny)    any image PIL can read (the reference's TIFFs included) -> (nx, ny) bool
  resample(src, nx, ny)       nearest-neighbour resampling  mask[x, y] = src[x*W//nx, y*H//ny]
  resize(img, nx, ny)         anti-aliased resize of a grey image (area average when shrinking, bilinear when
                              growing) -- what the notebook's skimage.transform.resize call is there for
                              (cells 11-16); threshold afterwards.  Same intent, not skimage's bits: its Gaussian
                              pre-filter is not reproduced.
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


def _resize_axis(a, n, axis):
    """Resize one axis of a float array to n samples: exact area average when shrinking (every output sample is the
    mean of the input interval it covers, fractional ends weighted), linear interpolation at pixel centres when
    growing."""
    a = np.moveaxis(np.asarray(a, dtype=np.float64), axis, 0)
    m = a.shape[0]
    if n == m:
        out = a
    elif n < m:
        # integral image along the axis, sampled at the (fractional) interval ends
        c = np.concatenate([np.zeros((1,) + a.shape[1:]), np.cumsum(a, axis=0)], axis=0)
        edges = np.arange(n + 1, dtype=np.float64) * m / n
        lo = np.minimum(np.floor(edges).astype(np.int64), m - 1)
        frac = (edges - lo).reshape((-1,) + (1,) * (a.ndim - 1))
        at = c[lo] + frac * a[lo]                     # integral of the step function up to each edge
        out = (at[1:] - at[:-1]) * (n / m)
    else:
        x = (np.arange(n, dtype=np.float64) + 0.5) * m / n - 0.5
        i0 = np.clip(np.floor(x).astype(np.int64), 0, m - 1)
        i1 = np.clip(i0 + 1, 0, m - 1)
        t = np.clip(x - i0, 0.0, 1.0).reshape((-1,) + (1,) * (a.ndim - 1))
        out = a[i0] * (1.0 - t) + a[i1] * t
    return np.moveaxis(out, 0, axis)


def resize(img, nx, ny):
    """Anti-aliased resize of a grey (W, H) image to (nx, ny), values kept in their range."""
    return _resize_axis(_resize_axis(img, nx, 0), ny, 1)


def from_image(path, nx=None, ny=None, threshold=0.5, solid_is_dark=False, antialias=False):
    """Read an image (PIL), convert to grey, transpose to (x, y), bring to (nx, ny) and threshold.
    antialias=False: threshold first, nearest-neighbour resampling of the mask (exact for binary masks and integer
    scale factors); antialias=True: area-average / bilinear resize of the grey image, then threshold -- the order of
    docs/cs205_movie.ipynb cells 11-16 (skimage.transform.resize, then `> threshold`)."""
    from PIL import Image
    img = np.asarray(Image.open(path).convert("L"), dtype=np.float64) / 255.0      # (H, W), row 0 = top
    grey = np.ascontiguousarray(img.T)                                              # (W, H) = (x, y)
    if antialias and nx is not None and ny is not None:
        grey = resize(grey, nx, ny)
    mask = (grey < threshold) if solid_is_dark else (grey >= threshold)
    if not antialias and nx is not None and ny is not None:
        mask = resample(mask, nx, ny)
    return np.ascontiguousarray(mask)


def pack(mask):
    mask = np.asarray(mask, dtype=bool)
    return dict(bits=np.packbits(mask.ravel()), shape=np.array(mask.shape))


def unpack(blob):
    shape = tuple(int(v) for v in blob["shape"])
    n = int(np.prod(shape))
    return np.unpackbits(np.asarray(blob["bits"]))[:n].astype(bool).reshape(shape)

"""TEST INFRASTRUCTURE: minimal stand-in for scikit-image (absent from this image).

The reference's cython_dim.pyx does `import skimage as ski` and calls
`ski.draw.circle` (cython_dim.pyx:11,430).  Only that one function is provided.
"""
from . import draw  # noqa: F401

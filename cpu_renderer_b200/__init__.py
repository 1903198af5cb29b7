"""B200-native rasterization back end for the hot path of MacSpain/cpu-renderer.

Only what the path needs lives here:
  csrc/     hand-written sm_100a CUDA kernels + the C ABI (include/b200_raster.h)
  api.py    ctypes mirror of the reference's host structs over that C ABI
  scene.py  the synthetic scenes pinned in SURVEY.md section 8d
The CPU oracle (oracle/) is test infrastructure and is never imported from this package.
"""
from . import scene  # noqa: F401
from .api import Renderer, B200RasterError, load_library  # noqa: F401

"""`soft_renderer.functional.soft_rasterize` (reference: functional/soft_rasterize.py:111-125),
routed to the sm_100a kernels.  Same positional arguments as the reference."""
from ...raster import SoftRasterizeFunction, soft_rasterize  # noqa: F401

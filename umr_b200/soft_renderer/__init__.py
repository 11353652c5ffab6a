"""Drop-in `soft_renderer` package (the names NVlabs/UMR imports from external/SoftRas,
`SoftRas/__init__.py:1-7`) backed by the sm_100a kernels of umr_b200.

Install it under the import name the reference uses with `umr_b200.compat.install()`:
`import soft_renderer as sr` then resolves to this package.
"""
from . import functional
from .mesh import Mesh
from .renderer import SoftRenderer
Renderer = SoftRenderer  # the reference's plain `Renderer` differs only in its rasteriser defaults
from .transform import Look, LookAt, Projection, Transform
from .lighting import AmbientLighting, DirectionalLighting, Lighting
from .rasterizer import SoftRasterizer
from .losses import LaplacianLoss, FlattenLoss

__version__ = "1.0.0+umr_b200"

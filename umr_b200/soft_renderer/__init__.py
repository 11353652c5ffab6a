"""Drop-in `soft_renderer` package backed by the sm_100a kernels of umr_b200.

It provides the names NVlabs/UMR takes from `external/SoftRas` (SoftRas/__init__.py:1-7): the mesh container,
the renderer and its three stages (lighting, camera transform, soft rasteriser), the two mesh regularisers
and the `functional` layer.  `umr_b200.compat.install()` registers it under the import name `soft_renderer`.
"""
__version__ = "1.0.0+umr_b200"

from . import functional  # noqa: E402
from .lighting import AmbientLighting, DirectionalLighting, Lighting  # noqa: E402
from .losses import FlattenLoss, LaplacianLoss  # noqa: E402
from .mesh import Mesh  # noqa: E402
from .rasterizer import SoftRasterizer  # noqa: E402
from .renderer import SoftRenderer  # noqa: E402
from .transform import Look, LookAt, Projection, Transform  # noqa: E402

Renderer = SoftRenderer  # the reference's plain `Renderer` differs only in its rasteriser defaults

__all__ = ["functional", "Mesh", "Renderer", "SoftRenderer", "Projection", "LookAt", "Look", "Transform",
           "AmbientLighting", "DirectionalLighting", "Lighting", "SoftRasterizer", "LaplacianLoss", "FlattenLoss"]

"""`sr.SoftRenderer`: lighting -> transform -> rasterise.  Reference: SoftRas/renderer.py:47-101 (argument names
and defaults; UMR constructs it with keywords only, nnutils/smr.py:56)."""
import torch.nn as nn

from ._args import bind, pick
from .lighting import Lighting
from .mesh import Mesh
from .rasterizer import SoftRasterizer
from .transform import Transform

# (name, default) in the reference's positional order
FIELDS = (
    # rasteriser
    ("image_size", 256), ("background_color", (0, 0, 0)), ("near", 1), ("far", 100), ("anti_aliasing", False),
    ("fill_back", True), ("eps", 1e-3), ("sigma_val", 1e-5), ("dist_func", "euclidean"), ("dist_eps", 1e-4),
    ("gamma_val", 1e-4), ("aggr_func_rgb", "softmax"), ("aggr_func_alpha", "prod"), ("texture_type", "surface"),
    # camera
    ("camera_mode", "projection"), ("P", None), ("dist_coeffs", None), ("orig_size", 512), ("perspective", True),
    ("viewing_angle", 30), ("viewing_scale", 1.0), ("eye", None), ("camera_direction", (0, 0, 1)),
    # light
    ("light_mode", "surface"), ("light_intensity_ambient", 0.5), ("light_color_ambient", (1, 1, 1)),
    ("light_intensity_directionals", 0.5), ("light_color_directionals", (1, 1, 1)), ("light_directions", (0, 1, 0)),
)
RASTER_KEYS = tuple((n, n) for n, _ in FIELDS[:14])
CAMERA_KEYS = tuple((n, n) for n, _ in FIELDS[14:23])
LIGHT_KEYS = (("light_mode", "light_mode"), ("light_intensity_ambient", "intensity_ambient"),
              ("light_color_ambient", "color_ambient"), ("light_intensity_directionals", "intensity_directionals"),
              ("light_color_directionals", "color_directionals"), ("light_directions", "directions"))


class SoftRenderer(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        cfg = bind("SoftRenderer", FIELDS, args, kwargs)
        self.lighting = Lighting(**pick(cfg, LIGHT_KEYS))
        self.transform = Transform(**pick(cfg, CAMERA_KEYS))
        self.rasterizer = SoftRasterizer(**pick(cfg, RASTER_KEYS))

    def set_sigma(self, sigma):
        self.rasterizer.sigma_val = sigma

    def set_gamma(self, gamma):
        self.rasterizer.gamma_val = gamma

    def set_texture_mode(self, mode):
        assert mode in ("vertex", "surface"), "Mode only support surface and vertex"
        self.lighting.light_mode = mode
        self.rasterizer.texture_type = mode

    def render_mesh(self, mesh, mode=None):
        self.set_texture_mode(mesh.texture_type)
        return self.rasterizer(self.transform(self.lighting(mesh)), mode)

    def forward(self, vertices, faces, textures=None, mode=None, texture_type="surface"):
        return self.render_mesh(Mesh(vertices, faces, textures=textures, texture_type=texture_type), mode)

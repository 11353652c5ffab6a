"""Camera transforms of the drop-in package.  Reference: SoftRas/transform.py (class / attribute names; UMR uses
`look_at` with `perspective=False` and overwrites `transform.transformer._eye`, nnutils/smr.py:56,60)."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as srf
from ._args import bind


def _default_eye(viewing_angle):
    return [0, 0, -(1. / math.tan(math.radians(viewing_angle)) + 1)]


class _EyeCamera(nn.Module):
    """Shared part of the two eye-based cameras: rigid transform, then perspective or orthogonal scaling."""

    def _configure(self, cfg):
        self.perspective = cfg["perspective"]
        self.viewing_angle = cfg["viewing_angle"]
        self.viewing_scale = cfg["viewing_scale"]
        self._eye = cfg["eye"] if cfg["eye"] is not None else _default_eye(cfg["viewing_angle"])

    def _project(self, vertices):
        if self.perspective:
            return srf.perspective(vertices, angle=self.viewing_angle)
        return srf.orthogonal(vertices, scale=self.viewing_scale)


class LookAt(_EyeCamera):
    FIELDS = (("perspective", True), ("viewing_angle", 30), ("viewing_scale", 1.0), ("eye", None))

    def __init__(self, *args, **kwargs):
        super().__init__()
        self._configure(bind("LookAt", self.FIELDS, args, kwargs))

    def forward(self, vertices):
        return self._project(srf.look_at(vertices, self._eye))


class Look(_EyeCamera):
    FIELDS = (("camera_direction", (0, 0, 1)),) + LookAt.FIELDS

    def __init__(self, *args, **kwargs):
        super().__init__()
        self._configure(bind("Look", self.FIELDS, args, kwargs))
        self.camera_direction = [0, 0, 1]  # the reference ignores its argument (transform.py:57)

    def forward(self, vertices):
        return self._project(srf.look(vertices, self._eye, self.camera_direction))


class Projection(nn.Module):
    def __init__(self, P, dist_coeffs=None, orig_size=512):
        super().__init__()
        if isinstance(P, np.ndarray):
            P = torch.from_numpy(P).cuda()
        if P is None or P.dim() != 3 or tuple(P.shape[1:]) != (3, 4):
            raise ValueError("You need to provide a valid (batch_size)x3x4 projection matrix")
        self.P, self.orig_size = P, orig_size
        self.dist_coeffs = dist_coeffs if dist_coeffs is not None else torch.zeros(P.shape[0], 5, device=P.device)

    def forward(self, vertices):
        return srf.projection(vertices, self.P, self.dist_coeffs, self.orig_size)


class Transform(nn.Module):
    FIELDS = (("camera_mode", "projection"), ("P", None), ("dist_coeffs", None), ("orig_size", 512), ("perspective", True),
              ("viewing_angle", 30), ("viewing_scale", 1.0), ("eye", None), ("camera_direction", (0, 0, 1)))

    def __init__(self, *args, **kwargs):
        super().__init__()
        c = bind("Transform", self.FIELDS, args, kwargs)
        self.camera_mode = c["camera_mode"]
        builders = {
            "look_at": lambda: LookAt(c["perspective"], c["viewing_angle"], c["viewing_scale"], c["eye"]),
            # INTENTIONAL DEVIATION: the reference passes (perspective, viewing_angle, viewing_scale,
            # eye, camera_direction) positionally into Look(camera_direction, perspective, ...) (transform.py:83-85),
            # shifting every argument by one; the arguments are bound by name here.  `look` is unused by UMR.
            "look": lambda: Look(c["camera_direction"], c["perspective"], c["viewing_angle"], c["viewing_scale"], c["eye"]),
            "projection": lambda: Projection(c["P"], c["dist_coeffs"], c["orig_size"]),
        }
        if self.camera_mode not in builders:
            raise ValueError("Camera mode has to be one of projection, look or look_at")
        self.transformer = builders[self.camera_mode]()

    def forward(self, mesh):
        mesh.vertices = self.transformer(mesh.vertices)
        return mesh

    def _require_eye_camera(self):
        if self.camera_mode not in ("look", "look_at"):
            raise ValueError("Projection does not need to set eyes")

    def set_eyes_from_angles(self, distances, elevations, azimuths):
        self._require_eye_camera()
        self.transformer._eye = srf.get_points_from_angles(distances, elevations, azimuths)

    def set_eyes(self, eyes):
        self._require_eye_camera()
        self.transformer._eye = eyes

    @property
    def eyes(self):
        return self.transformer._eye

"""Camera transform modules (reference: SoftRas/transform.py).  UMR uses `look_at` with
`perspective=False` and overwrites `transform.transformer._eye` (nnutils/smr.py:56,60)."""
import math

import torch.nn as nn

from . import functional as srf


class LookAt(nn.Module):
    def __init__(self, perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super().__init__()
        self.perspective = perspective
        self.viewing_angle = viewing_angle
        self.viewing_scale = viewing_scale
        self._eye = eye if eye is not None else [0, 0, -(1. / math.tan(math.radians(viewing_angle)) + 1)]

    def forward(self, vertices):
        vertices = srf.look_at(vertices, self._eye)
        if self.perspective:
            return srf.perspective(vertices, angle=self.viewing_angle)
        return srf.orthogonal(vertices, scale=self.viewing_scale)


class Transform(nn.Module):
    def __init__(self, camera_mode="projection", P=None, dist_coeffs=None, orig_size=512, perspective=True,
                 viewing_angle=30, viewing_scale=1.0, eye=None, camera_direction=(0, 0, 1)):
        super().__init__()
        self.camera_mode = camera_mode
        if camera_mode == "look_at":
            self.transformer = LookAt(perspective, viewing_angle, viewing_scale, eye)
        elif camera_mode in ("projection", "look"):
            raise NotImplementedError("camera_mode=%r is outside the UMR hot path (SURVEY.md §8f-3); "
                                      "use camera_mode='look_at'" % camera_mode)
        else:
            raise ValueError("Camera mode has to be one of projection, look or look_at")

    def forward(self, mesh):
        mesh.vertices = self.transformer(mesh.vertices)
        return mesh

    def set_eyes(self, eyes):
        self.transformer._eye = eyes

    @property
    def eyes(self):
        return self.transformer._eye

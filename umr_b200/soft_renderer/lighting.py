"""Lighting modules of the drop-in package (reference: SoftRas/lighting.py).

UMR pokes `renderer.lighting.ambient.light_intensity` and
`renderer.lighting.directionals[0].light_intensity` (nnutils/smr.py:63,70-71), so those attribute
paths are part of the API.
"""
import torch
import torch.nn as nn

from . import functional as srf


class AmbientLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1)):
        super().__init__()
        self.light_intensity = light_intensity
        self.light_color = light_color

    def forward(self, light):
        return srf.ambient_lighting(light, self.light_intensity, self.light_color)


class DirectionalLighting(nn.Module):
    def __init__(self, light_intensity=0.5, light_color=(1, 1, 1), light_direction=(0, 1, 0)):
        super().__init__()
        self.light_intensity = light_intensity
        self.light_color = light_color
        self.light_direction = light_direction

    def forward(self, light, normals):
        return srf.directional_lighting(light, normals, self.light_intensity, self.light_color,
                                        self.light_direction)


class Lighting(nn.Module):
    """textures *= ambient + sum_d directional_d(normals)  (lighting.py:50-67)."""

    def __init__(self, light_mode="surface", intensity_ambient=0.5, color_ambient=(1, 1, 1),
                 intensity_directionals=0.5, color_directionals=(1, 1, 1), directions=(0, 1, 0)):
        super().__init__()
        if light_mode not in ("surface", "vertex"):
            raise ValueError("Lighting mode only support surface and vertex")
        self.light_mode = light_mode
        self.ambient = AmbientLighting(intensity_ambient, color_ambient)
        self.directionals = nn.ModuleList([DirectionalLighting(intensity_directionals, color_directionals,
                                                               directions)])

    def _needs_normals(self):
        return any(float(d.light_intensity) != 0.0 for d in self.directionals)

    def forward(self, mesh):
        if self.light_mode == "surface":
            shape = (mesh.batch_size, mesh.num_faces, 3)
        else:
            shape = (mesh.batch_size, mesh.num_vertices, 3)
        light = torch.zeros(shape, dtype=torch.float32, device=mesh.device)
        light = self.ambient(light)
        if self._needs_normals():  # zero-intensity lights add exactly 0: skip the normal computation
            normals = mesh.surface_normals if self.light_mode == "surface" else mesh.vertex_normals
            for directional in self.directionals:
                light = directional(light, normals)
        if self.light_mode == "surface":
            mesh.textures = mesh.textures * light[:, :, None, :]   # [B,F,T2,3] * [B,F,1,3]
        else:
            mesh.textures = mesh.textures * light                  # [B,V,3] vertex colours
        return mesh

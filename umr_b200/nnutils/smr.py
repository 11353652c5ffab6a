"""`SoftRenderer` -- THE drop-in boundary (reference: nnutils/smr.py:49-87).

Same constructor, same methods (`forward`, `project_points`, `ambient_light_only`, `set_bgcolor`)
and the same attribute paths the reference pokes (`renderer.transform.transformer._eye`,
`renderer.lighting.ambient.light_intensity`, ...).  forward() returns
(images [B,4,is,is], p2f_info [B,F,2], aggrs_info [B,2,S,S]) with S = 2*is (anti_aliasing).
"""
import torch

from .. import soft_renderer as sr
from . import geom_utils


class Render(torch.nn.Module):
    """smr.py:29-44: y-flip, wrap into a Mesh, render."""

    def __init__(self, renderer):
        super().__init__()
        self.renderer = renderer

    def forward(self, vertices, faces, textures=None):
        vs = vertices
        vs[:, :, 1] *= -1  # in place, like the reference (smr.py:36)
        mesh_ = sr.Mesh(vs, faces) if textures is None else sr.Mesh(vs, faces, textures)
        return self.renderer.render_mesh(mesh_)


class SoftRenderer(torch.nn.Module):
    def __init__(self, img_size=256, render_type="softmax", background_color=[0, 0, 0], sigma_val=1e-5,
                 gamma_val=1e-4, dist_eps=1e-10, anti_aliasing=True):
        super().__init__()
        self.renderer = sr.SoftRenderer(image_size=img_size, aggr_func_rgb=render_type, camera_mode="look_at",
                                        sigma_val=sigma_val, dist_eps=dist_eps, gamma_val=gamma_val,
                                        background_color=background_color, anti_aliasing=anti_aliasing,
                                        perspective=False)
        self.renderer.transform.transformer._eye = [0, 0, -2.732]  # smr.py:60
        self.renderer.lighting.ambient.light_intensity = 0.8       # smr.py:63
        self.proj_fn = geom_utils.orthographic_proj_withz
        self.offset_z = 5.

    def ambient_light_only(self):
        self.renderer.lighting.ambient.light_intensity = 1
        self.renderer.lighting.directionals[0].light_intensity = 0

    def set_bgcolor(self, color):
        self.renderer.rasterizer.background_color = color

    def project_points(self, verts, cams):
        return self.proj_fn(verts, cams)[:, :, :2]

    def forward(self, vertices, faces, cams, textures=None):
        faces = faces.int()
        verts = self.proj_fn(vertices, cams, offset_z=self.offset_z)
        if textures is not None:
            return Render(self.renderer)(verts, faces, textures)
        return Render(self.renderer)(verts, faces)

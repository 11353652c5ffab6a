"""Vertex-side helpers of the drop-in package (pure torch, a handful of tiny ops).

References: functional/face_vertices.py:4-22, look_at.py:6-62, orthogonal.py:4-17,
perspective.py.  Only the camera modes UMR exercises (look_at + orthogonal, `smr.py:56`) plus the
perspective variant are provided; `projection`/`look` are out of scope.
"""
import math

import torch
import torch.nn.functional as F


def face_vertices(vertices, faces):
    """vertices [B,V,3], faces [B,F,3] (int) -> [B,F,3,3]: per-face corner coordinates."""
    if vertices.dim() != 3 or faces.dim() != 3 or vertices.shape[0] != faces.shape[0]:
        raise ValueError("face_vertices expects vertices [B,V,3] and faces [B,F,3]")
    B, V = vertices.shape[:2]
    idx = faces.long() + (torch.arange(B, device=vertices.device, dtype=torch.long) * V)[:, None, None]
    return vertices.reshape(B * V, vertices.shape[2])[idx]


def _as_batch(x, B, device):
    t = torch.as_tensor(x, dtype=torch.float32, device=device)
    if t.dim() == 1:
        t = t[None, :].expand(B, -1)
    return t


def look_at(vertices, eye, at=(0, 0, 0), up=(0, 1, 0)):
    """Camera frame with origin `eye` looking at `at` (look_at.py:48-60; normalise eps 1e-5)."""
    if vertices.dim() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    B, dev = vertices.shape[0], vertices.device
    eye, at, up = _as_batch(eye, B, dev), _as_batch(at, B, dev), _as_batch(up, B, dev)
    z_axis = F.normalize(at - eye, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=-1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=-1), eps=1e-5)
    rot = torch.stack((x_axis, y_axis, z_axis), dim=1)  # [B,3,3], rows = axes
    return torch.matmul(vertices - eye[:, None, :], rot.transpose(1, 2))


def orthogonal(vertices, scale):
    """x, y scaled; z kept (orthogonal.py:13-16)."""
    if vertices.dim() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    return torch.stack((vertices[:, :, 0] * scale, vertices[:, :, 1] * scale, vertices[:, :, 2]), dim=2)


def perspective(vertices, angle=30.0):
    """Pinhole projection with half field-of-view `angle` degrees (functional/perspective.py)."""
    if vertices.dim() != 3:
        raise ValueError("vertices Tensor should have 3 dimensions")
    width = math.tan(math.radians(float(angle)))
    z = vertices[:, :, 2]
    return torch.stack((vertices[:, :, 0] / z / width, vertices[:, :, 1] / z / width, z), dim=2)

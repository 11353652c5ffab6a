"""Hot functions of the reference's nnutils/geom_utils.py on the B200 path.

* `sample_textures` (geom_utils.py:41-59) -> hand-written bilinear sampler writing [B,F,T,T,C] directly.
* `orthographic_proj_withz` / `orthographic_proj` / `quat_rotate` / `hamilton_product`
  (geom_utils.py:62-91,119-165): closed-form torch (the reference builds the rotation from two
  Hamilton products through ~25 tiny stack/cat kernels; the arithmetic below is the same products
  written out, ~6 fused elementwise kernels).
"""
import torch

from .. import ops


def sample_textures(texture_flow, images):
    """texture_flow [B,F,T,T,2] in [-1,1], images [B,C,N,N] -> [B,F,T,T,C] (torch-1.1 grid_sample
    semantics == align_corners=True, zeros padding)."""
    B, nf, T = texture_flow.size(0), texture_flow.size(1), texture_flow.size(-2)
    C = images.size(1)
    out = ops.bilinear_sample(images, texture_flow.reshape(B, nf * T * T, 2))
    return out.view(B, nf, T, T, C)


def hamilton_product(qa, qb):
    """Quaternion product, last dim = (w, x, y, z)  (geom_utils.py:119-144)."""
    a0, a1, a2, a3 = qa.unbind(-1)
    b0, b1, b2, b3 = qb.unbind(-1)
    return torch.stack([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
                        a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                        a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
                        a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0], dim=-1)


def quat_rotate(X, q):
    """Rotate points X [B,N,3] by quaternions q [B,4]: (q (x) (0,X) (x) q*)_{1:4} (geom_utils.py:147-165)."""
    q = q[:, None, :].expand(X.size(0), X.size(1), 4)
    q_conj = torch.cat([q[..., :1], -q[..., 1:]], dim=-1)
    Xq = torch.cat([torch.zeros_like(X[..., :1]), X], dim=-1)
    return hamilton_product(q, hamilton_product(Xq, q_conj))[..., 1:4]


def orthographic_proj(X, cam):
    """X [B,N,3], cam [B,7] = [s, tx, ty, quat] -> [B,N,2] (geom_utils.py:62-72)."""
    X_rot = quat_rotate(X, cam[:, -4:])
    scale = cam[:, 0].contiguous().view(-1, 1, 1)
    trans = cam[:, 1:3].contiguous().view(cam.size(0), 1, -1)
    return scale * X_rot[:, :, :2] + trans


def orthographic_proj_withz(X, cam, offset_z=0.):
    """Orthographic projection keeping z (+offset_z)  (geom_utils.py:74-91)."""
    X_rot = quat_rotate(X, cam[:, -4:])
    scale = cam[:, 0].contiguous().view(-1, 1, 1)
    trans = cam[:, 1:3].contiguous().view(cam.size(0), 1, -1)
    proj = scale * X_rot
    return torch.cat((proj[:, :, :2] + trans, proj[:, :, 2, None] + offset_z), 2)

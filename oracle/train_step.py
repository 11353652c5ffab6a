"""CPU oracle of the train_s2-shaped loss step -- TEST INFRASTRUCTURE ONLY (never imported by umr_b200/).

A torch-CPU restatement of the reference's loss modules (`nnutils/loss_utils.py`) on top of
* `oracle/softras.py` oracle B for every render (wrapped in an autograd Function), and
* `oracle/losses.py` for the elementwise / sampler / chamfer / tex-cycle pieces,
so that a whole `experiments/train_s2.py:201-316` loss step (values AND gradients) can be checked
against the GPU product path.  Each function cites the reference lines it follows.

Exact-input protocol: the reference's raster output is chaotic w.r.t. 1-ulp
changes of the projected vertices, so every render accepts `fv_override` -- the face vertices the GPU
path actually rasterised (captured by the test).  The override replaces the VALUES only; gradients
still flow through the CPU vertex pipeline below (`fv + (override - fv).detach()`).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import losses as L
import softras


# ---------------------------------------------------------------------------------------------
# render = oracle B behind autograd
# ---------------------------------------------------------------------------------------------
class OracleRasterize(torch.autograd.Function):
    """SoftRasterizer.forward (rasterizer.py:42-55) + SoftRasterizeFunction (soft_rasterize.py:9-108)."""

    @staticmethod
    def forward(ctx, face_vertices, textures, image_size, rgb, kw):
        fv = face_vertices.detach().numpy().astype(np.float32)
        tx = textures.detach().numpy().astype(np.float32)
        images, fwd, cfg = softras.render(fv.reshape(fv.shape[0], -1, 9), tx, image_size, anti_aliasing=True,
                                          impl="B", aggr_func_rgb=rgb, **kw)
        ctx.fwd, ctx.cfg, ctx.shape = fwd, cfg, tuple(face_vertices.shape)
        p2f = torch.from_numpy(fwd["p2f_info"].copy())
        aggr = torch.from_numpy(fwd["aggrs_info"].copy())
        ctx.mark_non_differentiable(p2f, aggr)
        return torch.from_numpy(np.ascontiguousarray(images)), p2f, aggr

    @staticmethod
    def backward(ctx, g, _gp, _ga):
        gf, gt = softras.render_backward(ctx.fwd, ctx.cfg, g.contiguous().numpy(), anti_aliasing=True, impl="B")
        return torch.from_numpy(gf).view(ctx.shape), torch.from_numpy(gt), None, None, None


UMR_KW = dict(sigma_val=1e-5, dist_eps=1e-10, gamma_val=1e-4)


class OracleSoftRenderer:
    """nnutils/smr.py:49-87 + SURVEY.md App. A-1 (the vertex pipeline as UMR configures it)."""

    def __init__(self, img_size=256, render_type="softmax"):
        self.img_size, self.render_type = img_size, render_type
        self.ambient, self.directional = 0.8, 0.5   # smr.py:63, renderer.py:59
        self.overrides = None                        # list of captured face-vertex arrays, consumed in call order
        self.used = 0

    def ambient_light_only(self):                    # smr.py:68-71
        self.ambient, self.directional = 1.0, 0.0

    def project_points(self, verts, cams):           # smr.py:76-78
        return L.orthographic_proj_withz(verts, cams)[:, :, :2]

    def face_vertices(self, vertices, faces, cams):
        """-> (raster-space face vertices [B,F,3,3], flipped pre-transform face vertices [B,F,3,3])."""
        verts = L.orthographic_proj_withz(vertices, cams, offset_z=5.)           # smr.py:82
        verts = verts * torch.tensor([1., -1., 1.])                              # smr.py:36
        B, V = verts.shape[:2]
        idx = (faces.long() + (torch.arange(B) * V)[:, None, None]).reshape(-1)  # face_vertices.py:16-22
        pre = verts.reshape(B * V, 3)[idx].reshape(B, -1, 3, 3)
        post = (verts + torch.tensor([0., 0., 2.732])).reshape(B * V, 3)[idx].reshape(B, -1, 3, 3)  # look_at.py:48-60
        return post, pre

    def __call__(self, vertices, faces, cams, textures=None):
        fv, pre = self.face_vertices(vertices, faces, cams)
        B, nf = fv.shape[:2]
        if textures is None:
            textures = torch.ones(B, nf, 1, 3)                                   # mesh.py:46-50
        # lighting.py:50-57 (surface mode), normals from mesh.py:112-118
        light = torch.zeros(B, nf, 3) + self.ambient * torch.ones(1, 1, 3)
        if self.directional != 0.0:
            v10 = pre[:, :, 0] - pre[:, :, 1]
            v12 = pre[:, :, 2] - pre[:, :, 1]
            n = F.normalize(torch.cross(v12, v10, dim=2), p=2, dim=2, eps=1e-6)
            cosine = F.relu((n * torch.tensor([0., 1., 0.])).sum(dim=2))
            light = light + self.directional * (torch.ones(1, 1, 3) * cosine[:, :, None])
        textures = textures * light[:, :, None, :]
        if self.overrides is not None:
            ov = torch.from_numpy(np.asarray(self.overrides[self.used], np.float32)).view_as(fv)
            self.used += 1
            fv = fv + (ov - fv).detach()
        return OracleRasterize.apply(fv, textures, self.img_size, self.render_type, UMR_KW)

    forward = __call__


# ---------------------------------------------------------------------------------------------
# loss modules (reference: nnutils/loss_utils.py)
# ---------------------------------------------------------------------------------------------
def multi_mask_loss(renderer, vs, fs, cams_all_hypo, cam_probs, masks_gt, num_hypo_cams=8):
    """loss_utils.py:250-275."""
    bs, H = vs.size(0), num_hypo_cams
    pred_vs = vs.unsqueeze(1).repeat(1, H, 1, 1).view(-1, vs.size(1), 3)
    faces = fs.unsqueeze(1).repeat(1, H, 1, 1).view(-1, fs.size(1), 3)
    pred, _, _ = renderer(pred_vs, faces, cams_all_hypo.view(-1, 7))
    mask_all_hypo = pred[:, 3, :, :]
    masks = masks_gt.unsqueeze(1).repeat(1, H, 1, 1).view(-1, masks_gt.size(-2), masks_gt.size(-1))
    loss = L.neg_iou_loss(mask_all_hypo, masks, avg=False)
    loss = (loss.view(bs, -1) * cam_probs).sum(dim=1)
    return loss.mean(), mask_all_hypo


def multi_texture_loss(renderer, hard_renderer, vs, fs, cams_all_hypo, cam_probs, proj_cam, rgbs, masks_gt,
                       masks_pred, tx, tex_flow, dts_barrier, num_hypo_cams=8):
    """loss_utils.py:277-331 with the L1 texture loss (:289-292; LPIPS is outside the hot path)."""
    bs, H = vs.size(0), num_hypo_cams
    isz = rgbs.size(-1)
    pred_vs = vs.unsqueeze(1).repeat(1, H, 1, 1).view(-1, vs.size(1), 3)
    faces = fs.unsqueeze(1).repeat(1, H, 1, 1).view(-1, fs.size(1), 3)
    tex = tx.unsqueeze(1).repeat(1, H, 1, 1, 1).view(-1, tx.size(1), tx.size(2), 3)
    texture_rgba, _, _ = renderer(pred_vs.detach(), faces, cams_all_hypo.view(-1, 7), tex)
    texture_pred = texture_rgba[:, 0:3, :, :]
    imgs = rgbs.unsqueeze(1).repeat(1, H, 1, 1, 1).view(-1, 3, isz, isz)
    mgt = masks_gt.unsqueeze(1).repeat(1, H, 1, 1).view(-1, isz, isz)
    tex_loss = L.texture_loss_masks(texture_pred, imgs, mgt, masks_pred, avg=False)
    tex_loss = (tex_loss.view(bs, -1) * cam_probs).sum(dim=1).mean()
    tex_dt_loss = L.texture_dt_loss(tex_flow, dts_barrier)
    _, p2f_info, aggr_info = hard_renderer(vs.detach(), fs, proj_cam.detach())
    aggr = aggr_info[:, 1, :, :].reshape(bs, -1)
    tex_cycle_loss, _ = L.tex_cycle(tex_flow, p2f_info.detach(), aggr.detach())
    return tex_loss, tex_dt_loss, tex_cycle_loss, texture_pred


def corr_loss_chamfer(renderer, part_vertices, head_points, belly_points, neck_points, back_points, verts, cams,
                      avg=True):
    """loss_utils.py:194-248.  part_vertices = (head, belly, neck, back) index tensors."""
    head_v, belly_v, neck_v, back_v = [torch.as_tensor(p).long() for p in part_vertices]
    vert_coords = torch.cat((verts[:, head_v, :], verts[:, belly_v, :], verts[:, neck_v, :], verts[:, back_v, :]), dim=1)
    vert2d = renderer.project_points(vert_coords, cams)
    n0 = len(head_v)
    n1 = n0 + len(belly_v)
    n2 = n1 + len(neck_v)
    n3 = n2 + len(back_v)
    weights = [1, 1, 0, 0]
    head_c, _, _, _ = L.dist_chamfer(vert2d[:, :n0, :], head_points)
    belly_c, _, _, _ = L.dist_chamfer(vert2d[:, n0:n1, :], belly_points)
    neck_c, _, _, _ = L.dist_chamfer(vert2d[:, n1:n2, :], neck_points)
    back_c, _, _, _ = L.dist_chamfer(vert2d[:, n2:n3, :], back_points)
    cdist = torch.cat((head_c * weights[0], belly_c * weights[1], neck_c * weights[2], back_c * weights[3]), dim=1)
    loss = torch.mean(cdist, dim=1)
    if avg:
        return torch.mean(loss), vert2d
    return loss


def get_coordinate_tensors(x_max, y_max):
    """nnutils/scops_utils.py:12-19."""
    x_map = np.tile(np.arange(x_max), (y_max, 1)) / x_max * 2 - 1.0
    y_map = np.tile(np.arange(y_max), (x_max, 1)).T / y_max * 2 - 1.0
    return torch.from_numpy(x_map.astype(np.float32)), torch.from_numpy(y_map.astype(np.float32))


def get_center(part_map):
    """nnutils/scops_utils.py:21-35."""
    h, w = part_map.shape
    x_map, y_map = get_coordinate_tensors(h, w)
    x_center = (part_map * x_map).sum()
    y_center = (part_map * y_map).sum()
    return x_center, y_center


def batch_get_centers(pred_softmax):
    """nnutils/scops_utils.py:37-54 (Python B x C loops kept)."""
    B, C, H, W = pred_softmax.shape
    centers_list = []
    for b in range(B):
        centers = []
        for c in range(C):
            raw_pred = pred_softmax[b, c, :, :] + 1e-3
            k = raw_pred.sum()
            part_map = raw_pred / k
            x_c, y_c = get_center(part_map)
            centers.append(torch.stack((x_c, y_c), dim=0).unsqueeze(0))
        centers_list.append(torch.cat(centers, dim=0).unsqueeze(0))
    return torch.cat(centers_list, dim=0)


def part_matching_loss(renderer, stex_one_hot, verts, faces, cams, part_segs, cam_probs=None, avg=True):
    """loss_utils.py:333-440 (mse branch).  stex_one_hot [1,F,T2,5]."""
    bs = verts.size(0)
    isz = part_segs.size(-1)
    projs = []
    for k in (1, 2, 3, 4):                                                        # :385-399
        stex = stex_one_hot[:, :, :, k].unsqueeze(-1).repeat(bs, 1, 1, 3)
        p, _, _ = renderer(verts, faces, cams, stex)
        projs.append(torch.mean(p[:, 0:3, :, :], dim=1).unsqueeze(1))
    bg = torch.zeros(bs, 1, isz, isz)
    bg[:, 0, :, :] = 0.1
    proj = torch.cat([bg] + projs, dim=1)
    centers_proj = batch_get_centers(nn.Softmax(dim=1)(proj)[:, 1:, :, :])
    centers_parts = batch_get_centers(nn.Softmax(dim=1)(part_segs)[:, 1:, :, :])
    weights = torch.tensor([0, 5.0, 0.0, 0.0, 5.0]).view(1, 5, 1, 1)
    if avg:
        loss_lmeqv = F.mse_loss(centers_proj, centers_parts)
    else:
        loss_lmeqv = F.mse_loss(centers_proj, centers_parts, reduction="none")
        loss_lmeqv = torch.sum(loss_lmeqv, dim=(1, 2)) / (loss_lmeqv.size(1) * loss_lmeqv.size(2))
        loss_lmeqv = (loss_lmeqv.view(cam_probs.size()) * cam_probs).sum(dim=1).mean()
    max_proj, _ = torch.max(proj.view(bs, 5, -1), dim=2)
    max_proj = torch.where(max_proj < 1e-5, torch.full_like(max_proj, 1e-5), max_proj)   # :417-418
    proj_norm = proj / max_proj.view(bs, 5, 1, 1)
    max_part, _ = torch.max(part_segs.view(bs, 5, -1), dim=2)
    max_part = torch.where(max_part < 1e-5, torch.full_like(max_part, 1e-5), max_part)
    part_norm = part_segs / max_part.view(bs, 5, 1, 1)
    if avg:
        loss_eqv = torch.mean(nn.MSELoss(reduction="none")(proj_norm, part_norm) * weights)
    else:
        _, cs, iis, _ = part_norm.size()
        loss_eqv = nn.MSELoss(reduction="none")(proj_norm, part_norm) * weights
        loss_eqv = torch.sum(loss_eqv, dim=(1, 2, 3)) / (cs * iis * iis)
        loss_eqv = (loss_eqv.view(cam_probs.size()) * cam_probs).sum(dim=1).mean()
    return (loss_eqv + loss_lmeqv) / 4.0, projs

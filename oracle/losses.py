"""CPU oracles for the geometric losses -- TEST INFRASTRUCTURE ONLY (never imported by umr_b200/).

Each function restates one reference entry point with stock torch CPU ops, keeping the reference's
expression order.  torch 1.1.0 (`requirements.txt:8`) `grid_sample`/`affine_grid` semantics are
today's `align_corners=True`, passed explicitly here.
"""
import torch
import torch.nn.functional as F


def sample_textures(texture_flow, images):
    """nnutils/geom_utils.py:41-59.  flow [B,F,T,T,2], images [B,C,N,N] -> [B,F,T,T,C]."""
    T = texture_flow.size(-2)
    nf = texture_flow.size(1)
    C = images.size(1)
    flow_grid = texture_flow.view(-1, nf, T * T, 2)
    samples = F.grid_sample(images, flow_grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    samples = samples.view(-1, C, nf, T, T)
    return samples.permute(0, 2, 3, 4, 1)


def texture_dt_loss(texture_flow, dist_transf):
    """nnutils/loss_utils.py:50-90 (hot part :59-64, :90)."""
    T = texture_flow.size(-2)
    nf = texture_flow.size(1)
    flow_grid = texture_flow.view(-1, nf, T * T, 2)
    d = F.grid_sample(dist_transf, flow_grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return d.mean()


def neg_iou_loss(predict, target, avg=True):
    """nnutils/loss_utils.py:41-48."""
    dims = tuple(range(predict.ndimension())[1:])
    intersect = (predict * target).sum(dims)
    union = (predict + target - predict * target).sum(dims) + 1e-6
    if avg:
        return 1. - (intersect / union).sum() / intersect.nelement()
    return 1. - (intersect / union)


def texture_loss_masks(img_pred, img_gt, mask_gt, mask_pred, avg=True):
    """nnutils/loss_utils.py:103-116."""
    mask_gt = mask_gt.unsqueeze(1)
    mask_pred = mask_pred.unsqueeze(1)
    if avg:
        return torch.nn.L1Loss()(img_pred * mask_pred, img_gt * mask_gt)
    loss = torch.nn.L1Loss(reduction="none")(img_pred * mask_pred, img_gt * mask_gt)
    return torch.sum(loss, dim=(1, 2, 3)) / (loss.size(1) * loss.size(2) * loss.size(3))


def dist_chamfer(a, b):
    """nnutils/chamfer_python.py:43-64."""
    x, y = a, b
    bs, nx, _ = x.size()
    _, ny, _ = y.size()
    xx = torch.pow(x, 2).sum(2)
    yy = torch.pow(y, 2).sum(2)
    zz = torch.bmm(x, y.transpose(2, 1))
    rx = xx.unsqueeze(1).expand(bs, ny, nx)
    ry = yy.unsqueeze(1).expand(bs, nx, ny)
    P = rx.transpose(2, 1) + ry - 2 * zz
    return torch.min(P, 2)[0], torch.min(P, 1)[0], torch.min(P, 2)[1].int(), torch.min(P, 1)[1].int()


def tex_cycle(flow, prob, aggr_info):
    """nnutils/loss_utils.py:156-182.  flow [B,F,T,T,2], prob [B,F,2], aggr_info [B,P] face-id plane."""
    nb, nf = flow.size(0), flow.size(1)
    flow_grid = flow.view(nb, nf, -1, 2)
    avg_flow = torch.mean(flow_grid, dim=2)
    mask = torch.zeros(avg_flow.size())
    for cnt in range(nb):
        fids = torch.unique(aggr_info[cnt]).long()
        mask[cnt, fids, :] = 1  # -1 (background) indexes the last face, like the reference
    loss = torch.nn.MSELoss()(avg_flow * mask, prob * mask)
    return loss, avg_flow[0, 0:10, :]


def orthographic_proj_withz(X, cam, offset_z=0.):
    """nnutils/geom_utils.py:74-91 with quat_rotate :147-165 / hamilton_product :119-144."""
    quat = cam[:, -4:]
    X_rot = quat_rotate(X, quat)
    scale = cam[:, 0].contiguous().view(-1, 1, 1)
    trans = cam[:, 1:3].contiguous().view(cam.size(0), 1, -1)
    proj = scale * X_rot
    proj_xy = proj[:, :, :2] + trans
    proj_z = proj[:, :, 2, None] + offset_z
    return torch.cat((proj_xy, proj_z), 2)


def hamilton_product(qa, qb):
    qa_0, qa_1, qa_2, qa_3 = qa[:, :, 0], qa[:, :, 1], qa[:, :, 2], qa[:, :, 3]
    qb_0, qb_1, qb_2, qb_3 = qb[:, :, 0], qb[:, :, 1], qb[:, :, 2], qb[:, :, 3]
    q0 = qa_0 * qb_0 - qa_1 * qb_1 - qa_2 * qb_2 - qa_3 * qb_3
    q1 = qa_0 * qb_1 + qa_1 * qb_0 + qa_2 * qb_3 - qa_3 * qb_2
    q2 = qa_0 * qb_2 - qa_1 * qb_3 + qa_2 * qb_0 + qa_3 * qb_1
    q3 = qa_0 * qb_3 + qa_1 * qb_2 - qa_2 * qb_1 + qa_3 * qb_0
    return torch.stack([q0, q1, q2, q3], dim=-1)


def quat_rotate(X, q):
    ones_x = X[[0], :, :][:, :, [0]] * 0 + 1
    q = torch.unsqueeze(q, 1) * ones_x
    q_conj = torch.cat([q[:, :, [0]], -1 * q[:, :, 1:4]], dim=-1)
    X = torch.cat([X[:, :, [0]] * 0, X], dim=-1)
    X_rot = hamilton_product(q, hamilton_product(X, q_conj))
    return X_rot[:, :, 1:4]


def dist_chamfer_np(a, b):
    """nnutils/chamfer_python.py:43-64 with a DEFINED fp32 operation order (numpy float32, one rounding per
    product / sum, no FMA): |p|^2 = ((p0*p0 + p1*p1) + p2*p2), a.b likewise, P = (|a|^2 + |b|^2) - 2*(a.b).
    `torch.bmm` leaves the order/fusion of the K=2..3 dot product to the BLAS in use, so the index plane of
    the torch restatement above is only defined up to near-ties; this one is exact (lowest index on ties)."""
    import numpy as np
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    D = a.shape[2]

    def sq(p):
        s = p[..., 0] * p[..., 0]
        for d in range(1, D):
            s = s + p[..., d] * p[..., d]
        return s
    xx, yy = sq(a), sq(b)
    zz = a[:, :, None, 0] * b[:, None, :, 0]
    for d in range(1, D):
        zz = zz + a[:, :, None, d] * b[:, None, :, d]
    P = (xx[:, :, None] + yy[:, None, :]) - np.float32(2) * zz
    assert P.dtype == np.float32
    return P.min(2), P.min(1), P.argmin(2).astype(np.int32), P.argmin(1).astype(np.int32)

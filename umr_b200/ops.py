"""torch.autograd bindings of the geometric-loss kernels (C ABI: include/umr_b200.h).

CUDA only: there is no CPU fallback (the CPU oracles live in oracle/ and are test infrastructure).
"""
import ctypes

import torch

from . import _lib
from .raster import _ptr, _stream_ptr


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise TypeError("umr_b200 ops support only cuda Tensors")


def _batch_view(t, inner):
    """(tensor, batch stride in elements) if `t` [B, ...] is dense within each batch item, else a copy."""
    B = t.shape[0]
    if t[0].is_contiguous() and t.dtype == torch.float32:
        return t, (t.stride(0) if B > 1 else inner)
    t = t.contiguous().float()
    return t, inner


# -------------------------------------------------------------------------------------------------
# bilinear texture-flow sampler
# -------------------------------------------------------------------------------------------------
class BilinearSampleFunction(torch.autograd.Function):
    """images [B,C,H,W], flow [B,N,2] -> out [B,N,C]; bilinear, zeros padding, align_corners=True
    (the torch-1.1 semantics the reference was written for: geom_utils.py:55, loss_utils.py:64)."""

    @staticmethod
    def forward(ctx, images, flow):
        _need_cuda(images, flow)
        lib = _lib.load()
        img = images.detach().contiguous().float()
        fl = flow.detach().contiguous().float()
        B, C, H, W = img.shape
        N = fl.shape[1]
        with torch.cuda.device(img.device):
            out = torch.empty(B, N, C, device=img.device, dtype=torch.float32)
            rc = lib.umr_bilinear_sample_forward(_ptr(img), _ptr(fl), _ptr(out), B, C, H, W, N,
                                                 _stream_ptr(img.device))
        _lib.check(rc, "umr_bilinear_sample_forward")
        ctx.save_for_backward(img, fl)
        ctx.img_grad = images.requires_grad
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        img, fl = ctx.saved_tensors
        B, C, H, W = img.shape
        N = fl.shape[1]
        g = grad_out.contiguous().float()
        with torch.cuda.device(img.device):
            gflow = torch.empty_like(fl)
            gimg = torch.empty_like(img) if ctx.img_grad else None
            rc = lib.umr_bilinear_sample_backward(_ptr(img), _ptr(fl), _ptr(g), _ptr(gflow), _ptr(gimg),
                                                  B, C, H, W, N, _stream_ptr(img.device))
        _lib.check(rc, "umr_bilinear_sample_backward")
        return gimg, gflow


def bilinear_sample(images, flow):
    return BilinearSampleFunction.apply(images, flow)


# -------------------------------------------------------------------------------------------------
# silhouette IoU
# -------------------------------------------------------------------------------------------------
class NegIouFunction(torch.autograd.Function):
    """predict/target [B, ...] -> per-image loss [B] = 1 - sum(p*t) / (sum(p+t-p*t) + 1e-6)."""

    @staticmethod
    def forward(ctx, predict, target):
        _need_cuda(predict, target)
        lib = _lib.load()
        B = predict.shape[0]
        if target.shape[0] != B or target.numel() != predict.numel():
            raise ValueError("neg_iou_loss: predict %s and target %s must hold the same number of elements per image"
                             % (tuple(predict.shape), tuple(target.shape)))
        t = target.detach().contiguous().float().view(B, -1)
        N = t.shape[1]
        p, pbs = _batch_view(predict.detach(), N)  # e.g. the alpha plane of the RGBA render, read in place
        with torch.cuda.device(p.device):
            inter = torch.empty(B, device=p.device, dtype=torch.float32)
            uni = torch.empty_like(inter)
            loss = torch.empty_like(inter)
            rc = lib.umr_iou_forward(_ptr(p), pbs, _ptr(t), _ptr(inter), _ptr(uni), _ptr(loss), B, N,
                                     _stream_ptr(p.device))
        _lib.check(rc, "umr_iou_forward")
        ctx.save_for_backward(t, inter, uni)
        ctx.shape = tuple(predict.shape)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        t, inter, uni = ctx.saved_tensors
        B, N = t.shape
        g = grad_loss.contiguous().float()
        with torch.cuda.device(t.device):
            gp = torch.empty_like(t)
            rc = lib.umr_iou_backward(_ptr(t), _ptr(inter), _ptr(uni), _ptr(g), _ptr(gp), B, N,
                                      _stream_ptr(t.device))
        _lib.check(rc, "umr_iou_backward")
        return gp.view(ctx.shape), None


def neg_iou_per_image(predict, target):
    return NegIouFunction.apply(predict, target)


# -------------------------------------------------------------------------------------------------
# masked L1 texture loss
# -------------------------------------------------------------------------------------------------
class MaskedL1Function(torch.autograd.Function):
    """img_pred [B,C,H,W], img_gt [B,C,H,W], mask_gt [B,H,W], mask_pred [B,H,W] -> per-image
    mean |pred*mask_pred - gt*mask_gt| [B]  (loss_utils.py:103-116, avg=False form)."""

    @staticmethod
    def forward(ctx, img_pred, img_gt, mask_gt, mask_pred):
        _need_cuda(img_pred, img_gt, mask_gt, mask_pred)
        lib = _lib.load()
        B, C, H, W = img_pred.shape
        HW = H * W
        if tuple(img_gt.shape) != (B, C, H, W) or mask_gt.numel() != B * HW or mask_pred.numel() != B * HW:
            raise ValueError("texture_loss_masks: img_pred %s, img_gt %s, mask_gt %s, mask_pred %s do not match"
                             % (tuple(img_pred.shape), tuple(img_gt.shape), tuple(mask_gt.shape), tuple(mask_pred.shape)))
        p, pbs = _batch_view(img_pred.detach(), C * HW)
        mp, mbs = _batch_view(mask_pred.detach(), HW)
        g = img_gt.detach().contiguous().float()
        mg = mask_gt.detach().contiguous().float()
        with torch.cuda.device(p.device):
            loss = torch.empty(B, device=p.device, dtype=torch.float32)
            rc = lib.umr_masked_l1_forward(_ptr(p), pbs, _ptr(mp), mbs, _ptr(g), _ptr(mg), _ptr(loss), B, C, HW,
                                           _stream_ptr(p.device))
        _lib.check(rc, "umr_masked_l1_forward")
        ctx.save_for_backward(p, mp, g, mg)
        ctx.meta = (pbs, mbs, img_pred.requires_grad, mask_pred.requires_grad, tuple(img_pred.shape),
                    tuple(mask_pred.shape))
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        p, mp, g, mg = ctx.saved_tensors
        pbs, mbs, need_p, need_m, pshape, mshape = ctx.meta
        B, C = pshape[0], pshape[1]
        HW = pshape[2] * pshape[3]
        gl = grad_loss.contiguous().float()
        with torch.cuda.device(p.device):
            gp = torch.empty(pshape, device=p.device, dtype=torch.float32) if need_p else None
            gm = torch.empty(mshape, device=p.device, dtype=torch.float32) if need_m else None
            rc = lib.umr_masked_l1_backward(_ptr(p), pbs, _ptr(mp), mbs, _ptr(g), _ptr(mg), _ptr(gl), _ptr(gp),
                                            _ptr(gm), B, C, HW, _stream_ptr(p.device))
        _lib.check(rc, "umr_masked_l1_backward")
        return gp, None, None, gm


def masked_l1_per_image(img_pred, img_gt, mask_gt, mask_pred):
    return MaskedL1Function.apply(img_pred, img_gt, mask_gt, mask_pred)


# -------------------------------------------------------------------------------------------------
# fused loss head: w_iou * neg_iou_loss(alpha, mask) + w_tex * texture_loss_masks(rgb, gt, mask, alpha)
# -------------------------------------------------------------------------------------------------
class LossHeadFunction(torch.autograd.Function):
    """images [B,4,H,W] (one RGBA render), img_gt [B,3,H,W], mask_gt [B,H,W] -> (scalar loss, per_image [B,2]).
    One reduction + finalize forward, ONE backward kernel writing the whole [B,4,H,W] image gradient."""

    @staticmethod
    def forward(ctx, images, img_gt, mask_gt, w_iou, w_tex):
        _need_cuda(images, img_gt, mask_gt)
        lib = _lib.load()
        B, C, H, W = images.shape
        if C != 4 or tuple(img_gt.shape) != (B, 3, H, W) or mask_gt.numel() != B * H * W:
            raise ValueError("mask_texture_loss: images %s must be [B,4,H,W], img_gt %s [B,3,H,W], mask_gt %s [B,H,W]"
                             % (tuple(images.shape), tuple(img_gt.shape), tuple(mask_gt.shape)))
        x = images.detach().contiguous().float()
        g = img_gt.detach().contiguous().float()
        m = mask_gt.detach().contiguous().float()
        with torch.cuda.device(x.device):
            stats = torch.empty(B, 3, device=x.device, dtype=torch.float32)
            per_image = torch.empty(B, 2, device=x.device, dtype=torch.float32)
            loss = torch.empty((), device=x.device, dtype=torch.float32)
            rc = lib.umr_loss_head_forward(_ptr(x), _ptr(g), _ptr(m), _ptr(stats), _ptr(per_image), _ptr(loss), B, H * W,
                                           float(w_iou), float(w_tex), _stream_ptr(x.device))
        _lib.check(rc, "umr_loss_head_forward")
        ctx.save_for_backward(x, g, m, stats)
        ctx.w = (float(w_iou), float(w_tex))
        ctx.mark_non_differentiable(per_image)
        return loss, per_image

    @staticmethod
    def backward(ctx, grad_loss, _gp=None):
        lib = _lib.load()
        x, g, m, stats = ctx.saved_tensors
        B, _, H, W = x.shape
        gl = grad_loss.contiguous().float().reshape(1)
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x)
            rc = lib.umr_loss_head_backward(_ptr(x), _ptr(g), _ptr(m), _ptr(stats), _ptr(gl), _ptr(gx), B, H * W,
                                            ctx.w[0], ctx.w[1], _stream_ptr(x.device))
        _lib.check(rc, "umr_loss_head_backward")
        return gx, None, None, None, None


def mask_texture_loss(images, img_gt, mask_gt, w_iou=1.0, w_tex=1.0):
    """-> (loss, per_image[B,2] = (1 - IoU, masked L1)) with loss == w_iou * neg_iou_loss(images[:,3], mask_gt) +
    w_tex * texture_loss_masks(images[:,:3], img_gt, mask_gt, images[:,3])."""
    return LossHeadFunction.apply(images, img_gt, mask_gt, w_iou, w_tex)


# -------------------------------------------------------------------------------------------------
# chamfer
# -------------------------------------------------------------------------------------------------
class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _need_cuda(a, b)
        lib = _lib.load()
        x = a.detach().contiguous().float()
        y = b.detach().contiguous().float()
        B, N, D = x.shape
        M = y.shape[1]
        with torch.cuda.device(x.device):
            d_ab = torch.empty(B, N, device=x.device, dtype=torch.float32)
            d_ba = torch.empty(B, M, device=x.device, dtype=torch.float32)
            i_ab = torch.empty(B, N, device=x.device, dtype=torch.int32)
            i_ba = torch.empty(B, M, device=x.device, dtype=torch.int32)
            rc = lib.umr_chamfer_forward(_ptr(x), _ptr(y), _ptr(d_ab), _ptr(d_ba), _ptr(i_ab), _ptr(i_ba),
                                         B, N, M, D, _stream_ptr(x.device))
        _lib.check(rc, "umr_chamfer_forward")
        ctx.save_for_backward(x, y, i_ab, i_ba)
        ctx.mark_non_differentiable(i_ab, i_ba)
        return d_ab, d_ba, i_ab, i_ba

    @staticmethod
    def backward(ctx, g_ab, g_ba, _1=None, _2=None):
        lib = _lib.load()
        x, y, i_ab, i_ba = ctx.saved_tensors
        B, N, D = x.shape
        M = y.shape[1]
        g1 = g_ab.contiguous().float() if g_ab is not None else None
        g2 = g_ba.contiguous().float() if g_ba is not None else None
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x)
            gy = torch.empty_like(y)
            rc = lib.umr_chamfer_backward(_ptr(x), _ptr(y), _ptr(i_ab), _ptr(i_ba), _ptr(g1), _ptr(g2),
                                          _ptr(gx), _ptr(gy), B, N, M, D, _stream_ptr(x.device))
        _lib.check(rc, "umr_chamfer_backward")
        return gx, gy


def dist_chamfer(a, b):
    return ChamferFunction.apply(a, b)


# -------------------------------------------------------------------------------------------------
# texture cycle
# -------------------------------------------------------------------------------------------------
class TexCycleFunction(torch.autograd.Function):
    """flow [B,F,T2,2], prob [B,F,2], face_ids [B,P] (float plane, -1 = background) -> scalar loss."""

    @staticmethod
    def forward(ctx, flow, prob, face_ids):
        _need_cuda(flow, prob, face_ids)
        lib = _lib.load()
        fl = flow.detach().contiguous().float()
        pr = prob.detach().contiguous().float()
        ids = face_ids.detach().contiguous().float()
        B, F, T2 = fl.shape[0], fl.shape[1], fl.shape[2]
        P = ids.shape[1]
        with torch.cuda.device(fl.device):
            vis = torch.empty(B, F, device=fl.device, dtype=torch.uint8)
            loss = torch.empty(1, device=fl.device, dtype=torch.float32)
            rc = lib.umr_texcycle_forward(_ptr(fl), _ptr(pr), _ptr(ids), _ptr(vis), _ptr(loss), B, F, T2, P,
                                          _stream_ptr(fl.device))
        _lib.check(rc, "umr_texcycle_forward")
        ctx.save_for_backward(fl, pr, vis)
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        fl, pr, vis = ctx.saved_tensors
        B, F, T2 = fl.shape[0], fl.shape[1], fl.shape[2]
        g = grad_loss.contiguous().float().view(1)
        with torch.cuda.device(fl.device):
            gflow = torch.empty_like(fl)
            rc = lib.umr_texcycle_backward(_ptr(fl), _ptr(pr), _ptr(vis), _ptr(g), _ptr(gflow), B, F, T2,
                                           _stream_ptr(fl.device))
        _lib.check(rc, "umr_texcycle_backward")
        return gflow, None, None


def tex_cycle(flow, prob, face_ids):
    return TexCycleFunction.apply(flow, prob, face_ids)

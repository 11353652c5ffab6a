"""GPU parity of the loss kernels (through the C ABI bindings) vs the torch-CPU oracles in
oracle/losses.py.  Tolerance 1e-4 relative fp32 (+ small absolute floor)."""
import numpy as np
import pytest
import torch

import losses as oracle  # oracle/losses.py (test infrastructure)
from umr_b200 import ops
from umr_b200.nnutils import chamfer_python, geom_utils, loss_utils
from util import rel_report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _chk(name, got, ref, rtol=1e-4, atol=1e-6):
    ok, msg = rel_report(name, got.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol, atol)
    print(msg)
    assert ok, msg


@pytest.mark.parametrize("C,H,W", [(3, 64, 64), (1, 48, 80), (3, 256, 256)])
def test_sample_textures(C, H, W):
    g = torch.Generator().manual_seed(0)
    B, F, T = 2, 50, 6
    img = torch.rand(B, C, H, W, generator=g)
    flow = torch.rand(B, F, T, T, 2, generator=g) * 2.4 - 1.2  # includes out-of-range samples (zero padding)
    flow_ref = flow.clone().requires_grad_(True)
    img_ref = img.clone().requires_grad_(True)
    ref = oracle.sample_textures(flow_ref, img_ref)
    w = torch.rand(ref.shape, generator=g)
    (ref * w).sum().backward()
    flow_g = flow.to(DEV).requires_grad_(True)
    img_g = img.to(DEV).requires_grad_(True)
    got = geom_utils.sample_textures(flow_g, img_g)
    assert got.shape == ref.shape
    (got * w.to(DEV)).sum().backward()
    _chk("sample fwd", got, ref)
    _chk("sample dflow", flow_g.grad, flow_ref.grad, 1e-4, 1e-4 * float(flow_ref.grad.abs().max()) * 1e-2 + 1e-6)
    _chk("sample dimage", img_g.grad, img_ref.grad, 1e-4, 1e-5)


def test_texture_dt_loss():
    g = torch.Generator().manual_seed(1)
    dt = torch.rand(3, 1, 64, 64, generator=g)
    flow = (torch.rand(3, 40, 6, 6, 2, generator=g) * 2 - 1)
    fr = flow.clone().requires_grad_(True)
    ref = oracle.texture_dt_loss(fr, dt)
    ref.backward()
    fg = flow.to(DEV).requires_grad_(True)
    got = loss_utils.texture_dt_loss(fg, dt.to(DEV))
    got.backward()
    _chk("tex_dt", got, ref)
    _chk("tex_dt dflow", fg.grad, fr.grad, 1e-4, 1e-9)


@pytest.mark.parametrize("shape,avg", [((4, 64, 64), False), ((4, 64, 64), True), ((3, 37, 41), False), ((2, 256, 256), True)])
def test_neg_iou(shape, avg):
    g = torch.Generator().manual_seed(2)
    p = torch.rand(shape, generator=g)
    t = (torch.rand(shape, generator=g) > 0.5).float()
    pr = p.clone().requires_grad_(True)
    ref = oracle.neg_iou_loss(pr, t, avg=avg)
    wgt = torch.rand(ref.shape, generator=g) if not avg else torch.tensor(1.0)
    (ref * wgt).sum().backward()
    pg = p.to(DEV).requires_grad_(True)
    got = loss_utils.neg_iou_loss(pg, t.to(DEV), avg=avg)
    (got * wgt.to(DEV)).sum().backward()
    _chk("iou", got, ref)
    _chk("iou dp", pg.grad, pr.grad, 1e-4, 1e-10)


@pytest.mark.parametrize("B,N,M,D", [(4, 40, 10, 2), (3, 80, 30, 2), (1, 5000, 642, 2), (2, 33, 7, 3),
                                     (1, 20000, 642, 2)])  # last: the eval-size call of experiments/test_kp.py:180
def test_chamfer(B, N, M, D):
    g = torch.Generator().manual_seed(3)
    a = torch.rand(B, N, D, generator=g) - 0.5
    b = torch.rand(B, M, D, generator=g) - 0.5
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    r = oracle.dist_chamfer(ar, br)
    w1, w2 = torch.rand(B, N, generator=g), torch.rand(B, M, generator=g)
    ((r[0] * w1).sum() + (r[1] * w2).sum()).backward()
    ag, bg = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    o = chamfer_python.distChamfer(ag, bg)
    ((o[0] * w1.to(DEV)).sum() + (o[1] * w2.to(DEV)).sum()).backward()
    assert o[2].dtype == torch.int32 and o[3].dtype == torch.int32
    _chk("d_ab", o[0], r[0], 1e-4, 1e-6)
    _chk("d_ba", o[1], r[1], 1e-4, 1e-6)
    # index planes: BIT-EXACT (distances and argmins) against the defined-order fp32 oracle ...
    n = oracle.dist_chamfer_np(a.numpy(), b.numpy())
    for k in range(4):
        assert np.array_equal(o[k].detach().cpu().numpy(), n[k]), "chamfer output %d differs from the fp32 oracle" % k
    # ... and against torch.bmm (whose K<=3 dot-product order belongs to the BLAS) any argmin difference must be a
    # near-tie: the two candidates' float64 distances differ by no more than the fp32 rounding of the expanded form
    P = ((a.double()[:, :, None, :] - b.double()[:, None, :, :]) ** 2).sum(-1)
    for k, dim in ((2, 2), (3, 1)):
        ours, ref = o[k].cpu().long(), r[k].long()
        mism = ours != ref
        print("argmin mismatches vs torch.bmm oracle: %d / %d" % (int(mism.sum()), mism.numel()))
        if mism.any():
            d_ours = torch.gather(P, dim, ours.unsqueeze(dim)).squeeze(dim)
            d_ref = torch.gather(P, dim, ref.unsqueeze(dim)).squeeze(dim)
            assert float((d_ours - d_ref).abs()[mism].max()) <= 4 * 1.2e-7
    _chk("da", ag.grad, ar.grad, 1e-4, 1e-5)
    _chk("db", bg.grad, br.grad, 1e-4, 1e-4)


def test_chamfer_ties_lowest_index():
    a = torch.zeros(1, 3, 2)
    b = torch.tensor([[[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]]])
    o = chamfer_python.distChamfer(a.to(DEV), b.to(DEV))
    assert o[2].cpu().tolist() == [[0, 0, 0]]
    assert o[3].cpu().tolist() == [[0, 0, 0, 0]]


def test_tex_cycle():
    g = torch.Generator().manual_seed(4)
    B, F, T, P = 3, 64, 6, 32 * 32
    flow = torch.rand(B, F, T, T, 2, generator=g) * 2 - 1
    prob = torch.rand(B, F, 2, generator=g) * 2 - 1
    ids = torch.randint(-1, F // 2, (B, P), generator=g).float()   # -1 = background, upper half never visible
    ids[2] = 5.0                                                    # one sample without background
    fr = flow.clone().requires_grad_(True)
    ref, vis_ref = oracle.tex_cycle(fr, prob, ids)
    ref.backward()
    fg = flow.to(DEV).requires_grad_(True)
    got, vis = loss_utils.TexCycle()(fg, prob.to(DEV), ids.to(DEV))
    got.backward()
    _chk("texcycle", got, ref)
    _chk("texcycle vis", vis, vis_ref, 1e-5, 1e-7)
    _chk("texcycle dflow", fg.grad, fr.grad, 1e-4, 1e-9)


def test_ops_reject_cpu():
    with pytest.raises(TypeError):
        ops.bilinear_sample(torch.zeros(1, 1, 4, 4), torch.zeros(1, 3, 2))


@pytest.mark.parametrize("avg", [False, True])
def test_texture_loss_masks_fused(avg):
    g = torch.Generator().manual_seed(5)
    B, H = 3, 40
    rgba = torch.rand(B, 4, H, H, generator=g)
    gt = torch.rand(B, 3, H, H, generator=g)
    mgt = (torch.rand(B, H, H, generator=g) > 0.4).float()
    r = rgba.clone().requires_grad_(True)
    ref = oracle.texture_loss_masks(r[:, :3], gt, mgt, r[:, 3], avg=avg)
    w = torch.rand(ref.shape, generator=g) if not avg else torch.tensor(1.0)
    (ref * w).sum().backward()
    x = rgba.to(DEV).requires_grad_(True)
    got = loss_utils.texture_loss_masks(x[:, :3], gt.to(DEV), mgt.to(DEV), x[:, 3], avg=avg)  # strided views
    (got * w.to(DEV)).sum().backward()
    _chk("masked l1", got, ref)
    _chk("masked l1 drgba", x.grad, r.grad, 1e-4, 1e-9)


@pytest.mark.parametrize("B,H", [(3, 40), (16, 256), (2, 37)])
def test_fused_loss_head_equals_the_two_reference_losses(B, H):
    """loss_utils.mask_texture_loss == 2.5 * neg_iou_loss + 3.0 * texture_loss_masks (loss_utils.py:41-48, :103-116) on
    the same RGBA render, value and gradient, against the torch-CPU oracle composition."""
    g = torch.Generator().manual_seed(6)
    rgba = torch.rand(B, 4, H, H, generator=g)
    gt = torch.rand(B, 3, H, H, generator=g)
    mgt = (torch.rand(B, H, H, generator=g) > 0.4).float()
    r = rgba.clone().requires_grad_(True)
    ref = 2.5 * oracle.neg_iou_loss(r[:, 3], mgt) + 3.0 * oracle.texture_loss_masks(r[:, :3], gt, mgt, r[:, 3])
    (ref * 1.7).backward()
    x = rgba.to(DEV).requires_grad_(True)
    got = loss_utils.mask_texture_loss(x, gt.to(DEV), mgt.to(DEV), 2.5, 3.0)
    (got * 1.7).backward()
    _chk("loss head", got, ref, 1e-5, 1e-7)
    _chk("loss head drgba", x.grad, r.grad, 1e-4, 1e-9)
    # and the unfused product ops give the same number
    y = rgba.to(DEV)
    unfused = 2.5 * loss_utils.neg_iou_loss(y[:, 3], mgt.to(DEV)) + 3.0 * loss_utils.texture_loss_masks(y[:, :3], gt.to(DEV), mgt.to(DEV), y[:, 3])
    _chk("fused vs unfused", got, unfused, 1e-5, 1e-7)


@pytest.mark.parametrize("avg", [True, False])
@pytest.mark.parametrize("shared_mesh", [False, True])
def test_fused_corr_loss_chamfer_equals_the_composition(avg, shared_mesh):
    """csrc/vertex.cu k_corr_fwd / k_corr_bwd vs the module's torch composition (project_points + 4 x distChamfer + cat +
    mean, loss_utils.py:218-248 -- itself checked against the CPU oracle in test_train_step_gpu.py): same loss, projected
    vertices, and gradients for the vertices (also through an expanded one-mesh view) and the cameras."""
    import numpy as np
    from umr_b200 import synth
    from umr_b200.nnutils import loss_utils
    rng = np.random.default_rng(17)
    B, IS = 6, 64
    v, f = synth.icosphere(3)
    V = v.shape[0]
    parts = [torch.from_numpy(p) for p in synth.part_vertex_sets(rng, V, sizes=(20, 40, 20, 40))]
    pts = [torch.from_numpy(p).to(DEV) for p in synth.part_points(rng, B)]
    m = loss_utils.CorrLossChamfer(None, IS, part_vertices=parts)
    m.weights = [1, 1, 0.5, 0.25]          # exercise all four parts (the reference's [1, 1, 0, 0] zeroes two of them)
    base = torch.from_numpy(synth.bird_like(v, rng, B))
    cams0 = torch.from_numpy(synth.cameras(rng, B))
    outs = []
    for fused in (True, False):
        cams = cams0.clone().to(DEV).requires_grad_(True)
        if shared_mesh:
            leaf = base[0].clone().to(DEV).requires_grad_(True)
            verts = leaf[None].expand(B, -1, -1)
        else:
            leaf = base.clone().to(DEV).requires_grad_(True)
            verts = leaf
        if not fused:
            m.renderer.proj_fn = lambda X, cam, offset_z=0.: loss_utils.geom_utils.orthographic_proj_withz(X, cam, offset_z)
        out = m(pts[0], pts[1], pts[2], pts[3], verts, cams, avg=avg)
        if avg:
            loss, v2d = out
            (loss + 0.01 * v2d.sum()).backward()
        else:
            loss, v2d = out, None
            (loss * torch.linspace(0.5, 1.5, B, device=DEV)).sum().backward()
        outs.append((loss.detach().cpu(), None if v2d is None else v2d.detach().cpu(), leaf.grad.cpu(), cams.grad.cpu()))
    (l1, v1, gv1, gc1), (l0, v0, gv0, gc0) = outs
    assert torch.allclose(l1, l0, rtol=1e-5, atol=1e-7)
    if v1 is not None:
        assert torch.equal(v1, v0)      # same projection arithmetic, bit for bit
    assert torch.allclose(gv1, gv0, rtol=1e-4, atol=1e-6 * float(gv0.abs().max()))
    assert torch.allclose(gc1, gc0, rtol=1e-4, atol=1e-5 * float(gc0.abs().max()))

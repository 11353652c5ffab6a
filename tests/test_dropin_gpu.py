"""Drop-in boundary on the GPU: nnutils.smr.SoftRenderer / loss modules with the reference's call
signatures.  The raster-space face vertices our host glue produces are captured and
fed to oracle B, so the comparison is exact-input (the reference output is chaotic w.r.t. 1-ulp vertex
changes, App. B-15); the torch glue itself is checked against the torch-CPU restatement."""
import numpy as np
import pytest
import torch

import losses as oracle_losses
import softras
from umr_b200 import raster, synth
from umr_b200.nnutils import loss_utils, smr
from util import rel_report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B=2, subdiv=2, seed=0, tex_res=None):
    rng = np.random.default_rng(seed)
    v, f = synth.icosphere(subdiv)
    verts = torch.from_numpy(synth.bird_like(v, rng, B))
    faces = torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1)
    cams = torch.from_numpy(synth.cameras(rng, B))
    tex = None
    if tex_res:
        tex = torch.from_numpy(rng.uniform(0, 1, size=(B, f.shape[0], tex_res ** 2, 3)).astype(np.float32))
    return verts, faces, cams, tex


class Capture:
    """Records the (face_vertices, textures) our glue hands to the rasteriser."""

    def __enter__(self):
        self.calls = []
        self.orig = raster.SoftRasterizeFunction.apply
        outer = self

        def spy(fv, tex, *a):
            outer.calls.append((fv.detach().cpu().numpy().copy(), tex.detach().cpu().numpy().copy(), a))
            return outer.orig(fv, tex, *a)
        raster.SoftRasterizeFunction.apply = staticmethod(spy)
        return self

    def __exit__(self, *exc):
        raster.SoftRasterizeFunction.apply = self.orig


@pytest.mark.parametrize("render_type,tex_res", [("softmax", None), ("softmax", 3), ("hard", 2)])
def test_soft_renderer_forward_matches_oracle(render_type, tex_res):
    verts, faces, cams, tex = _inputs(tex_res=tex_res)
    r = smr.SoftRenderer(64, render_type)
    with Capture() as cap:
        images, p2f, aggr = r(verts.to(DEV), faces.to(DEV), cams.to(DEV), None if tex is None else tex.to(DEV))
    assert images.shape == (2, 4, 64, 64) and aggr.shape == (2, 2, 128, 128) and p2f.shape == (2, faces.shape[1], 2)
    fv, tx, _ = cap.calls[0]
    ref_img, fwd, _ = softras.render(fv.reshape(2, -1, 9), tx, 64, anti_aliasing=True, impl="B",
                                     aggr_func_rgb=render_type, sigma_val=1e-5, dist_eps=1e-10, gamma_val=1e-4)
    for name, got, ref in [("images", images, ref_img), ("aggrs", aggr, fwd["aggrs_info"]), ("p2f", p2f, fwd["p2f_info"])]:
        ok, msg = rel_report(name, got.detach().cpu().numpy(), ref, 1e-4, 1e-6)
        print(msg)
        assert ok, msg
    if render_type == "hard":
        assert float(p2f.abs().max()) == 0.0  # reference quirk B-4: hard renderer never accumulates p2f
    # host glue: projected vertices == torch-CPU restatement of smr.py:80-87 / App. A-1
    pv = oracle_losses.orthographic_proj_withz(verts, cams, offset_z=5.)
    pv[:, :, 1] *= -1
    pv = pv + torch.tensor([0, 0, 2.732])
    ref_fv = pv.reshape(-1, 3)[(faces + (torch.arange(2) * verts.shape[1])[:, None, None]).reshape(-1)].reshape(2, -1, 9)
    ok, msg = rel_report("face_vertices glue", fv.reshape(2, -1, 9), ref_fv.numpy(), 1e-5, 1e-6)
    print(msg)
    assert ok, msg
    # lighting (a4): default SoftRenderer = ambient 0.8 + directional 0.5 along +y (smr.py:63, renderer.py:57-60).
    # VALUE check against the closed form of lighting.py:50-57 / mesh.py:112-118 evaluated on the CPU:
    # light = 0.8 + 0.5 * relu(n_y), n = normalize(cross(v2 - v1, v0 - v1)) of the flipped, pre-transform faces
    if tex is not None:
        assert tx.shape == tex.shape
        import train_step as O
        _, pre = O.OracleSoftRenderer(64).face_vertices(verts, faces, cams)
        n = torch.nn.functional.normalize(torch.cross(pre[:, :, 2] - pre[:, :, 1], pre[:, :, 0] - pre[:, :, 1], dim=2),
                                          p=2, dim=2, eps=1e-6)
        light = 0.8 + 0.5 * torch.relu(n[:, :, 1])
        ok, msg = rel_report("lit textures (a4)", tx, (tex * light[:, :, None, None]).numpy(), 2e-5, 1e-6)
        print(msg)
        assert ok, msg
        assert float(light.min()) < 0.81 and float(light.max()) > 1.2   # both the shadowed and the lit side occur


def test_project_points_and_bgcolor():
    verts, faces, cams, _ = _inputs()
    r = smr.SoftRenderer(32)
    p = r.project_points(verts.to(DEV), cams.to(DEV)).cpu()
    ref = oracle_losses.orthographic_proj_withz(verts, cams)[:, :, :2]
    ok, msg = rel_report("project_points", p.numpy(), ref.numpy(), 1e-5, 1e-6)
    assert ok, msg
    r.set_bgcolor([0.25, 0.5, 0.75])
    r.ambient_light_only()
    img, _, _ = r(verts.to(DEV), faces.to(DEV), cams.to(DEV))
    corner = img[:, :3, 0, 0].cpu().numpy()
    assert np.allclose(corner, [[0.25, 0.5, 0.75]] * 2, atol=1e-6)


def test_multi_mask_and_texture_loss_run_and_backprop():
    """train_s2-shaped call pattern (loss_utils.py:250-331) at small size: values finite, gradients
    reach vertices / textures / texture flow, hard-render p2f quirk preserved."""
    B, H = 2, 8
    rng = np.random.default_rng(5)
    v, f = synth.icosphere(2)
    vs = torch.from_numpy(synth.bird_like(v, rng, B)).to(DEV).requires_grad_(True)
    fs = torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1).to(DEV)
    cams = torch.from_numpy(np.stack([synth.cameras(rng, H) for _ in range(B)])).to(DEV)  # [B,8,7]
    probs = torch.softmax(torch.randn(B, H, device=DEV), 1)
    masks = torch.from_numpy(synth.ellipse_masks(rng, B, 32)).to(DEV)
    mask_loss, mask_all = loss_utils.MultiMaskLoss(32, "softmax", H).to(DEV)(vs, fs, cams, probs, masks)
    assert mask_all.shape == (B * H, 32, 32)
    mask_loss.backward()
    assert torch.isfinite(vs.grad).all() and vs.grad.abs().sum() > 0

    imgs = torch.from_numpy(synth.smooth_images(rng, B, 32)).to(DEV)
    flow = torch.from_numpy(synth.texture_flow(rng, B, f.shape[0], 3)).to(DEV).requires_grad_(True)
    tx = loss_utils.geom_utils.sample_textures(flow, imgs).reshape(B, f.shape[0], 9, 3)
    dts = torch.from_numpy(np.stack([synth.dt_barrier(m) for m in masks.cpu().numpy()]))[:, None].to(DEV)
    mtl = loss_utils.MultiTextureLoss(B, H, 32, "softmax", "l1", "smr").to(DEV)
    tl, tdt, tcyc, pred = mtl(vs.detach(), fs, cams, probs, cams[:, 0], imgs, masks, mask_all.detach(), tx, flow, dts)
    assert pred.shape == (B * H, 3, 32, 32)
    (tl + tdt + tcyc).backward()
    assert torch.isfinite(flow.grad).all() and flow.grad.abs().sum() > 0


@pytest.mark.parametrize("ambient_only,with_tex", [(False, False), (False, True), (True, True), (True, False)])
def test_fused_vertex_pipeline_matches_torch_glue(ambient_only, with_tex):
    """csrc/vertex.cu (projection + flip + look_at + gather + light in one kernel) vs the generic torch-op
    path of the drop-in package: forward bit-identical, backward within fp32 summation tolerance."""
    verts, faces, cams, tex = _inputs(B=3, subdiv=2, seed=9, tex_res=2 if with_tex else None)
    outs = []
    for fused in (True, False):
        r = smr.SoftRenderer(32, "softmax")
        r.fuse_vertex_pipeline = fused
        if ambient_only:
            r.ambient_light_only()
        v = verts.clone().to(DEV).requires_grad_(True)
        c = cams.clone().to(DEV).requires_grad_(True)
        t = tex.clone().to(DEV).requires_grad_(True) if with_tex else None
        with Capture() as cap:
            img, p2f, aggr = r(v, faces.to(DEV), c, t)
        w = torch.linspace(0.5, 1.5, img.numel(), device=DEV).view_as(img)
        (img * w).sum().backward()
        outs.append(dict(fv=cap.calls[0][0], tex=cap.calls[0][1], img=img.detach().cpu().numpy(),
                         gv=v.grad.cpu().numpy(), gc=c.grad.cpu().numpy(),
                         gt=t.grad.cpu().numpy() if with_tex else None))
    a, b = outs
    assert np.array_equal(a["fv"].reshape(b["fv"].shape), b["fv"]), "fused projection is not bit-identical"
    ok, msg = rel_report("lit textures", a["tex"], b["tex"], 1e-6, 1e-7)
    assert ok, msg
    for k, at in (("gv", 1e-5), ("gc", 1e-4), ("gt", 1e-6)):
        if a[k] is None:
            continue
        scale = float(np.abs(b[k]).max()) + 1e-30
        ok, msg = rel_report(k, a[k], b[k], 2e-4, at * scale)
        print(msg)
        assert ok, msg


def test_part_matching_loss_packed_renders_match_reference_pattern():
    """part_matching_loss (loss_utils.py:333-440): ONE 4-colour-channel render gives the same
    projections as the reference's 4 separate one-hot renders; loss and gradients agree."""
    from umr_b200 import raster
    B, IS, T = 2, 32, 2
    rng = np.random.default_rng(11)
    v, f = synth.icosphere(2)
    F_ = f.shape[0]
    part = rng.integers(0, 5, size=(F_, T * T))
    one_hot = torch.zeros(1, F_, T * T, 5)
    one_hot.scatter_(3, torch.from_numpy(part)[None, :, :, None], 1.0)
    verts0 = torch.from_numpy(synth.bird_like(v, rng, B))
    faces = torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1).to(DEV)
    cams = torch.from_numpy(synth.cameras(rng, B)).to(DEV)
    part_segs = torch.rand(B, 5, IS, IS, generator=torch.Generator().manual_seed(1)).to(DEV)
    outs = []
    for pack in (True, False):
        m = loss_utils.part_matching_loss(None, None, 0, im_size=IS, batch_size=B, tex_size=T, stex_one_hot=one_hot).to(DEV)
        m.pack_parts = pack
        vv = verts0.clone().to(DEV).requires_grad_(True)
        sink = []
        raster.set_profile_sink(sink)
        try:
            loss, projs = m(vv, faces, cams, part_segs)
            loss.backward()
        finally:
            raster.set_profile_sink(None)
        kinds = [k for k, _ in sink]
        assert kinds.count("fwd") == (1 if pack else 4) and kinds.count("bwd") == (1 if pack else 4), kinds
        outs.append((loss.item(), [p.detach().cpu() for p in projs], vv.grad.cpu()))
    for pa, pb in zip(outs[0][1], outs[1][1]):
        assert torch.equal(pa, pb)
    assert abs(outs[0][0] - outs[1][0]) <= 1e-6 * max(1.0, abs(outs[1][0]))
    ok, msg = rel_report("dverts", outs[0][2].numpy(), outs[1][2].numpy(), 1e-3, 1e-5 * float(outs[1][2].abs().max()))
    assert ok, msg


def test_camera_hypotheses_are_broadcast_not_materialised():
    """SoftRenderer.forward(vs [B], fs [B], cams [B*H], tx [B]) == the reference's call with
    repeat(1, H, ...) copies (loss_utils.py:260-261, 303-305): same images, gradients summed over the hypotheses,
    and no [B*H,F,T2,3] texture tensor is ever allocated."""
    B, H, IS, T = 2, 8, 32, 6
    rng = np.random.default_rng(3)
    v, f = synth.icosphere(2)
    F_ = f.shape[0]
    vs0 = torch.from_numpy(synth.bird_like(v, rng, B))
    fs = torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1).to(DEV)
    cams = torch.from_numpy(np.stack([synth.cameras(rng, H) for _ in range(B)])).view(-1, 7).to(DEV)
    tx0 = torch.from_numpy(rng.uniform(0, 1, size=(B, F_, T * T, 3)).astype(np.float32))
    w = torch.linspace(0.5, 1.5, B * H * 4 * IS * IS, device=DEV).view(B * H, 4, IS, IS)
    outs, peaks = [], []
    for tiled in (False, True):
        r = smr.SoftRenderer(IS, "softmax")
        r.ambient_light_only()
        vs = vs0.clone().to(DEV).requires_grad_(True)
        tx = tx0.clone().to(DEV).requires_grad_(True)
        c = cams.clone().requires_grad_(True)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        if tiled:
            img, _, _ = r(loss_utils.tile_hypotheses(vs, H), loss_utils.tile_hypotheses(fs, H), c, loss_utils.tile_hypotheses(tx, H))
        else:
            img, _, _ = r(vs, fs, c, tx)
        peaks.append(torch.cuda.max_memory_allocated() - base)
        (img * w).sum().backward()
        outs.append((img.detach().cpu().numpy(), vs.grad.cpu().numpy(), tx.grad.cpu().numpy(), c.grad.cpu().numpy()))
    a, b = outs
    assert np.array_equal(a[0], b[0]), "broadcast render differs from the materialised one"
    for k, name in ((1, "dverts"), (2, "dtex"), (3, "dcams")):
        ok, msg = rel_report(name, a[k], b[k], 2e-4, 2e-5 * float(np.abs(b[k]).max()))
        print(msg)
        assert ok, msg
    tiled_tex_bytes = B * H * F_ * T * T * 3 * 4
    assert peaks[1] - peaks[0] >= 0.9 * tiled_tex_bytes, (peaks, tiled_tex_bytes)

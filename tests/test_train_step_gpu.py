"""SURVEY.md §8(b) acceptance: a train_s2-shaped synthetic loss step on the GPU product path vs the
CPU oracle composition (oracle/train_step.py = oracle-B renders + oracle/losses.py).

The four loss modules are constructed with the arguments of `experiments/train_s2.py:128-164` and
called like `:201-316`; EVERY loss value and EVERY gradient (vertices, camera hypotheses, camera
probabilities, texture flow) is compared.  The oracle rasterises the face vertices the GPU path
produced (captured), see oracle/train_step.py "exact-input protocol".
"""
import numpy as np
import pytest
import torch

import train_step as O  # oracle/train_step.py (test infrastructure)
from umr_b200 import raster, synth
from umr_b200.nnutils import geom_utils, loss_utils
from util import rel_report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, H, IS, T = 2, 8, 32, 3


class Capture:
    """Records the face vertices our glue hands to the rasteriser, one entry per raster call."""

    def __enter__(self):
        self.fv = []
        self.orig = raster.SoftRasterizeFunction.apply
        outer = self

        def spy(fv, tex, *a):
            outer.fv.append(fv.detach().cpu().numpy().copy())
            return outer.orig(fv, tex, *a)
        raster.SoftRasterizeFunction.apply = staticmethod(spy)
        # the visibility-only kernel (MultiTextureLoss's hard render) is a raster call too
        from umr_b200.soft_renderer import rasterizer as rz
        self.orig_vis = rz.visibility

        def spy_vis(fv, *a, **kw):
            outer.fv.append(fv.detach().cpu().numpy().copy())
            return outer.orig_vis(fv, *a, **kw)
        rz.visibility = spy_vis
        return self

    def __exit__(self, *exc):
        from umr_b200.soft_renderer import rasterizer as rz
        raster.SoftRasterizeFunction.apply = self.orig
        rz.visibility = self.orig_vis


def _scene(seed=21):
    rng = np.random.default_rng(seed)
    v, f = synth.icosphere(2)
    d = dict(
        vs=torch.from_numpy(synth.bird_like(v, rng, B)),
        fs=torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1),
        cams=torch.from_numpy(np.stack([synth.cameras(rng, H) for _ in range(B)])),   # [B,8,7]
        probs=torch.softmax(torch.from_numpy(rng.normal(size=(B, H)).astype(np.float32)), 1),
        masks=torch.from_numpy(synth.ellipse_masks(rng, B, IS)),
        imgs=torch.from_numpy(synth.smooth_images(rng, B, IS)),
        flow=torch.from_numpy(synth.texture_flow(rng, B, f.shape[0], T)),
        part_segs=torch.from_numpy(rng.uniform(0, 1, size=(B, 5, IS, IS)).astype(np.float32)),
    )
    d["dts"] = torch.from_numpy(np.stack([synth.dt_barrier(m) for m in d["masks"].numpy()]))[:, None]
    d["V"], d["F"] = v.shape[0], f.shape[0]
    part = rng.integers(0, 5, size=(f.shape[0], T * T))
    one_hot = torch.zeros(1, f.shape[0], T * T, 5)
    one_hot.scatter_(3, torch.from_numpy(part)[None, :, :, None], 1.0)
    d["one_hot"] = one_hot
    d["part_vertices"] = synth.part_vertex_sets(rng, v.shape[0], sizes=(20, 40, 20, 40))
    d["part_points"] = [torch.from_numpy(p) for p in synth.part_points(rng, B)]
    return d


def _leaf(t, dev=None):
    t = t.clone()
    if dev:
        t = t.to(dev)
    return t.requires_grad_(True)


def _cmp(name, got, ref, rtol=1e-4, atol_scale=1e-5):
    got = got.detach().cpu().numpy()
    ref = ref.detach().cpu().numpy()
    at = atol_scale * float(np.abs(ref).max()) + 1e-9
    ok, msg = rel_report(name, got, ref, rtol, at)
    print(msg)
    assert ok, msg


def test_multi_mask_loss_value_and_gradients():
    """loss_utils.py:250-275 as wired at train_s2.py:130-133, called like :216-218."""
    d = _scene()
    m = loss_utils.MultiMaskLoss(IS, "softmax", H).to(DEV)
    vs, cams, probs = _leaf(d["vs"], DEV), _leaf(d["cams"], DEV), _leaf(d["probs"], DEV)
    with Capture() as cap:
        loss, mask_all = m(vs, d["fs"].to(DEV), cams, probs, d["masks"].to(DEV))
    loss.backward()
    r = O.OracleSoftRenderer(IS, "softmax")
    r.overrides = cap.fv
    ovs, ocams, oprobs = _leaf(d["vs"]), _leaf(d["cams"]), _leaf(d["probs"])
    oloss, omask = O.multi_mask_loss(r, ovs, d["fs"], ocams, oprobs, d["masks"], H)
    oloss.backward()
    assert r.used == len(cap.fv) == 1
    _cmp("mask_loss", loss, oloss)
    _cmp("mask_all_hypo", mask_all, omask, 1e-4, 1e-6)
    _cmp("d/dverts", vs.grad, ovs.grad)
    _cmp("d/dcams", cams.grad, ocams.grad)
    _cmp("d/dcam_probs", probs.grad, oprobs.grad)


def test_multi_texture_loss_values_and_gradients():
    """loss_utils.py:277-331 (L1 branch), called like train_s2.py:236-249: textures sampled from the
    texture flow; gradients reach the flow through sample_textures, the dt loss and the cycle loss."""
    d = _scene(seed=22)
    fs = d["fs"].to(DEV)
    masks_pred = torch.rand(B * H, IS, IS, generator=torch.Generator().manual_seed(3))
    m = loss_utils.MultiTextureLoss(B, H, IS, "softmax", "l1", "smr").to(DEV)
    flow = _leaf(d["flow"], DEV)
    tx = geom_utils.sample_textures(flow, d["imgs"].to(DEV)).contiguous().view(B, d["F"], T * T, 3)
    with Capture() as cap:
        out = m(d["vs"].to(DEV), fs, d["cams"].to(DEV), d["probs"].to(DEV), d["cams"][:, 0].to(DEV),
                d["imgs"].to(DEV), d["masks"].to(DEV), masks_pred.to(DEV), tx, flow, d["dts"].to(DEV))
    tl, tdt, tcyc, pred = out
    (3.0 * tl + 3.0 * tdt + 1.0 * tcyc).backward()   # train_s2.py:49-59 weights
    r, rh = O.OracleSoftRenderer(IS, "softmax"), O.OracleSoftRenderer(IS, "hard")
    r.ambient_light_only()
    r.overrides, rh.overrides = cap.fv[0:1], cap.fv[1:2]
    oflow = _leaf(d["flow"])
    otx = O.L.sample_textures(oflow, d["imgs"]).contiguous().view(B, d["F"], T * T, 3)
    otl, otdt, otcyc, opred = O.multi_texture_loss(r, rh, d["vs"], d["fs"], d["cams"], d["probs"], d["cams"][:, 0],
                                                   d["imgs"], d["masks"], masks_pred, otx, oflow, d["dts"], H)
    (3.0 * otl + 3.0 * otdt + 1.0 * otcyc).backward()
    assert len(cap.fv) == 2 and r.used == 1 and rh.used == 1
    _cmp("tex_loss", tl, otl)
    _cmp("tex_dt_loss", tdt, otdt)
    _cmp("tex_cycle_loss", tcyc, otcyc)
    _cmp("texture_pred", pred, opred, 1e-4, 1e-6)
    _cmp("d/dtex_flow", flow.grad, oflow.grad)


def test_corr_loss_chamfer_value_and_gradients():
    """loss_utils.py:194-248, called like train_s2.py:300-315 (8 hypotheses, avg=False, argument order
    head, belly, BACK, NECK as the reference passes them -- SURVEY.md App. B-12)."""
    d = _scene(seed=23)
    m = loss_utils.CorrLossChamfer(None, IS, part_vertices=[torch.from_numpy(p) for p in d["part_vertices"]])
    rep = lambda t: t.unsqueeze(1).repeat(1, H, 1, 1).view(-1, t.size(1), t.size(2))
    head, belly, neck, back = d["part_points"]
    mean_shape = _leaf(d["vs"][0], DEV)
    ms_rep = mean_shape.unsqueeze(0).repeat(B, 1, 1).unsqueeze(1).repeat(1, H, 1, 1).view(-1, d["V"], 3)
    cams = _leaf(d["cams"], DEV)
    loss = m(rep(head).to(DEV), rep(belly).to(DEV), rep(back).to(DEV), rep(neck).to(DEV), ms_rep, cams.view(-1, 7),
             avg=False)
    total = (loss.view(B, H) * d["probs"].to(DEV)).sum(dim=1).mean()
    total.backward()
    oms = _leaf(d["vs"][0])
    oms_rep = oms.unsqueeze(0).repeat(B, 1, 1).unsqueeze(1).repeat(1, H, 1, 1).view(-1, d["V"], 3)
    ocams = _leaf(d["cams"])
    oloss = O.corr_loss_chamfer(O.OracleSoftRenderer(IS), d["part_vertices"], rep(head), rep(belly), rep(back),
                                rep(neck), oms_rep, ocams.view(-1, 7), avg=False)
    ototal = (oloss.view(B, H) * d["probs"]).sum(dim=1).mean()
    ototal.backward()
    _cmp("corr per-render", loss, oloss, 1e-4, 1e-6)
    _cmp("corr_loss", total, ototal)
    _cmp("d/dmean_shape", mean_shape.grad, oms.grad)
    _cmp("d/dcams", cams.grad, ocams.grad)
    # avg=True returns (scalar, projected part vertices) -- loss_utils.py:244-245
    s, v2d = m(head.to(DEV), belly.to(DEV), neck.to(DEV), back.to(DEV), d["vs"].to(DEV), d["cams"][:, 0].to(DEV))
    os_, ov2d = O.corr_loss_chamfer(O.OracleSoftRenderer(IS), d["part_vertices"], head, belly, neck, back, d["vs"],
                                    d["cams"][:, 0])
    _cmp("corr avg", s, os_)
    _cmp("vert2d", v2d, ov2d, 1e-5, 1e-6)


@pytest.mark.parametrize("avg", [True, False])
def test_part_matching_loss_value_and_gradients(avg):
    """loss_utils.py:333-440 as wired at train_s2.py:153-160 and called at :293-295: the oracle renders the
    four one-hot part maps separately like the reference; ours packs them (umr_b200 loss_utils)."""
    d = _scene(seed=24)
    m = loss_utils.part_matching_loss(None, None, 0, im_size=IS, batch_size=B, tex_size=T,
                                      stex_one_hot=d["one_hot"]).to(DEV)
    vs, cams = _leaf(d["vs"], DEV), _leaf(d["cams"][:, 0], DEV)
    probs = torch.softmax(torch.arange(B, dtype=torch.float32).view(B, 1), 0)  # cam_probs shape [B,1] (num_cam = 1)
    with Capture() as cap:
        loss, projs = m(vs, d["fs"].to(DEV), cams, d["part_segs"].to(DEV), cam_probs=None if avg else probs.to(DEV),
                        avg=avg)
    loss.backward()
    r = O.OracleSoftRenderer(IS, "softmax")
    r.ambient_light_only()
    r.overrides = [cap.fv[0]] * 4   # all part renders rasterise the same projected mesh
    ovs, ocams = _leaf(d["vs"]), _leaf(d["cams"][:, 0])
    oloss, oprojs = O.part_matching_loss(r, d["one_hot"], ovs, d["fs"], ocams, d["part_segs"],
                                         cam_probs=None if avg else probs, avg=avg)
    oloss.backward()
    assert all(np.array_equal(cap.fv[0], f) for f in cap.fv)
    _cmp("part_loss", loss, oloss)
    for k, (p, op) in enumerate(zip(projs, oprojs)):
        _cmp("proj%d" % (k + 1), p, op, 1e-4, 1e-6)
    _cmp("d/dverts", vs.grad, ovs.grad, 2e-4, 2e-5)
    _cmp("d/dcams", cams.grad, ocams.grad, 2e-4, 2e-5)

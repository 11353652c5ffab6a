"""CPU checks of the train-step oracle composition (oracle/train_step.py): it runs end to end on oracle B,
its autograd wiring reaches every input, and the host-side pieces the product re-implements differently
(vectorised soft centroids, hypothesis tiling) agree with the reference-shaped loops of the oracle."""
import numpy as np
import torch

import train_step as O
from umr_b200 import synth
from umr_b200.nnutils import loss_utils

B, H, IS, T = 2, 8, 16, 2


def _scene():
    rng = np.random.default_rng(3)
    v, f = synth.icosphere(1)
    return dict(vs=torch.from_numpy(synth.bird_like(v, rng, B)), fs=torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1),
                cams=torch.from_numpy(np.stack([synth.cameras(rng, H) for _ in range(B)])),
                probs=torch.softmax(torch.from_numpy(rng.normal(size=(B, H)).astype(np.float32)), 1),
                masks=torch.from_numpy(synth.ellipse_masks(rng, B, IS)), imgs=torch.from_numpy(synth.smooth_images(rng, B, IS)),
                flow=torch.from_numpy(synth.texture_flow(rng, B, f.shape[0], T)), F=f.shape[0], V=v.shape[0], rng=rng)


def test_oracle_step_runs_and_backprops():
    d = _scene()
    vs = d["vs"].clone().requires_grad_(True)
    cams = d["cams"].clone().requires_grad_(True)
    probs = d["probs"].clone().requires_grad_(True)
    loss, mask_all = O.multi_mask_loss(O.OracleSoftRenderer(IS), vs, d["fs"], cams, probs, d["masks"], H)
    loss.backward()
    assert mask_all.shape == (B * H, IS, IS)
    for g in (vs.grad, cams.grad, probs.grad):
        assert torch.isfinite(g).all() and g.abs().sum() > 0
    flow = d["flow"].clone().requires_grad_(True)
    tx = O.L.sample_textures(flow, d["imgs"]).contiguous().view(B, d["F"], T * T, 3)
    r, rh = O.OracleSoftRenderer(IS), O.OracleSoftRenderer(IS, "hard")
    r.ambient_light_only()
    dts = torch.rand(B, 1, IS, IS)
    tl, tdt, tcyc, pred = O.multi_texture_loss(r, rh, d["vs"], d["fs"], d["cams"], d["probs"], d["cams"][:, 0], d["imgs"],
                                               d["masks"], mask_all.detach(), tx, flow, dts, H)
    (tl + tdt + tcyc).backward()
    assert pred.shape == (B * H, 3, IS, IS) and torch.isfinite(flow.grad).all() and flow.grad.abs().sum() > 0


def test_face_vertex_override_replaces_values_but_keeps_gradients():
    d = _scene()
    r = O.OracleSoftRenderer(IS)
    fv, _ = r.face_vertices(d["vs"], d["fs"], d["cams"][:, 0])
    shifted = fv.numpy() + np.float32(0.01)
    r.overrides = [shifted]
    vs = d["vs"].clone().requires_grad_(True)
    img, _, _ = r(vs, d["fs"], d["cams"][:, 0])
    assert r.used == 1
    # alpha does not depend on the (lit) textures: rasterising the shifted vertices directly gives the same plane
    img2, _, _ = O.OracleRasterize.apply(torch.from_numpy(shifted), torch.ones(B, d["F"], 1, 3), IS, "softmax", O.UMR_KW)
    assert torch.equal(img[:, 3], img2[:, 3])
    base, _, _ = O.OracleSoftRenderer(IS)(d["vs"], d["fs"], d["cams"][:, 0])
    assert not torch.equal(img[:, 3], base[:, 3])
    img[:, 3].sum().backward()
    assert vs.grad.abs().sum() > 0


def test_vectorised_soft_centroids_match_the_reference_loops():
    x = torch.rand(3, 4, 12, 12, generator=torch.Generator().manual_seed(0))
    a = loss_utils.batch_get_centers(x)
    b = O.batch_get_centers(x)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_part_matching_oracle_matches_product_host_logic_on_cpu(monkeypatch):
    """The product's packed 2-render part_matching_loss vs the oracle's 4 separate renders, both on oracle B
    (the product's SoftRenderer is swapped for the oracle one): isolates the host logic."""
    d = _scene()
    rng = d["rng"]
    part = rng.integers(0, 5, size=(d["F"], T * T))
    one_hot = torch.zeros(1, d["F"], T * T, 5)
    one_hot.scatter_(3, torch.from_numpy(part)[None, :, :, None], 1.0)
    segs = torch.rand(B, 5, IS, IS, generator=torch.Generator().manual_seed(1))
    m = loss_utils.part_matching_loss(None, None, 0, im_size=IS, batch_size=B, tex_size=T, stex_one_hot=one_hot)
    ro = O.OracleSoftRenderer(IS)
    ro.ambient_light_only()
    del m._modules["renderer"]
    m.__dict__["renderer"] = ro
    vs = d["vs"].clone().requires_grad_(True)
    loss, _ = m(vs, d["fs"], d["cams"][:, 0], segs)
    loss.backward()
    r = O.OracleSoftRenderer(IS)
    r.ambient_light_only()
    ovs = d["vs"].clone().requires_grad_(True)
    oloss, _ = O.part_matching_loss(r, one_hot, ovs, d["fs"], d["cams"][:, 0], segs)
    oloss.backward()
    assert abs(loss.item() - oloss.item()) <= 1e-6 * max(1.0, abs(oloss.item()))
    assert torch.allclose(vs.grad, ovs.grad, rtol=1e-3, atol=1e-5 * float(ovs.grad.abs().max()))

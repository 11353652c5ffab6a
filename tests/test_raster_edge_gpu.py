"""GPU parity on the edge cases the tiled kernels must get right (vs CPU oracle B): empty / off-screen
meshes, everything inside one tile (long lists: several record chunks), more faces than one cull-box
piece (F > 2048), degenerate triangles, near/far rejection, single-sided rendering, non-default
sigma/gamma/background, hard-mode ties on a mirror-symmetric mesh."""
import numpy as np
import pytest
import torch

import softras
from umr_b200 import raster, synth
from util import rel_report, scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def both(fv, tex, isz, aa=True, grad=True, tex_grad=True, **kw):
    kw.setdefault("sigma_val", 1e-5)
    kw.setdefault("dist_eps", 1e-10)
    kw.setdefault("gamma_val", 1e-4)
    rgb = kw.get("aggr_func_rgb", "softmax")
    img, fwd, cfg = softras.render(fv, tex, isz, anti_aliasing=aa, impl="B", **kw)
    g = np.random.default_rng(3).normal(size=img.shape).astype(np.float32)
    gf, gt = softras.render_backward(fwd, cfg, g, anti_aliasing=aa, impl="B")
    tfv = torch.from_numpy(fv).to(DEV).requires_grad_(grad)
    ttex = torch.from_numpy(tex).to(DEV).requires_grad_(grad and tex_grad)
    out, p2f, aggr = raster.soft_rasterize(tfv, ttex, isz, anti_aliasing=aa, **kw)
    res = [("images", out.detach().cpu().numpy(), img), ("aggrs", aggr.cpu().numpy(), fwd["aggrs_info"]),
           ("p2f", p2f.cpu().numpy(), fwd["p2f_info"])]
    if grad:
        out.backward(torch.from_numpy(g).to(DEV))
        res.append(("grad_faces", tfv.grad.cpu().numpy(), gf))
        if tex_grad:
            res.append(("grad_tex", ttex.grad.cpu().numpy(), gt))
    ok, msgs = True, []
    for name, a, b in res:
        at = 1e-6
        if name.startswith("grad"):
            at = 1e-6 * float(np.abs(b).max() + 1e-30) + 1e-7
        o, m = rel_report(name, a, b, 1e-4, at)
        ok &= o
        msgs.append(m)
    if rgb == "hard":
        ex = np.array_equal(aggr.cpu().numpy(), fwd["aggrs_info"])
        msgs.append("hard planes bit-exact: %s" % ex)
        ok &= ex
    print("\n".join(msgs))
    assert ok, "\n" + "\n".join(msgs)
    return out


def test_offscreen_mesh_renders_background_and_zero_grads():
    fv, tex = scene(2, 1, 1, seed=1)
    fv[:, :, 0::3] += 10.0
    out = both(fv, tex, 32, background_color=(0.2, 0.4, 0.6))
    assert torch.allclose(out[:, :3].amax(dim=(2, 3)).cpu(), torch.tensor([[0.2, 0.4, 0.6]] * 2))
    assert float(out[:, 3].abs().max()) == 0.0


def test_tiny_mesh_all_faces_in_one_tile_long_lists():
    fv, tex = scene(1, 3, 2, seed=2)      # 1280 faces squeezed into a ~12-pixel blob: lists of ~1000 faces
    c = fv[:, :, 0::3].mean(), fv[:, :, 1::3].mean()
    fv[:, :, 0::3] = (fv[:, :, 0::3] - c[0]) * 0.08 + 0.03
    fv[:, :, 1::3] = (fv[:, :, 1::3] - c[1]) * 0.08 - 0.02
    both(fv, tex, 64)
    both(fv, tex, 64, aggr_func_rgb="hard")


def test_more_faces_than_one_cull_box_piece():
    fv, tex = scene(1, 4, 1, seed=3)      # 5120 faces > BOX_PIECE (2048): three TMA pieces per tile
    both(fv, tex, 96)
    both(fv, tex, 48, aa=False, aggr_func_rgb="hard")


def test_degenerate_and_sliver_triangles():
    fv, tex = scene(1, 2, 2, seed=4)
    fv[0, 0, 3:6] = fv[0, 0, 0:3]                     # two coincident corners
    fv[0, 1, :] = np.tile(fv[0, 1, 0:3], 3)           # a point
    fv[0, 2, 6:8] = (fv[0, 2, 0:2] + fv[0, 2, 3:5]) / 2 + 1e-7   # near-collinear sliver
    both(fv, tex, 48)


def test_near_far_rejection_and_single_sided():
    fv, tex = scene(2, 2, 2, seed=5)
    fv[0, :, 2::3] -= 7.2                              # part of mesh 0 in front of the near plane (z < 1)
    both(fv, tex, 40, near=1.0, far=8.0)               # and mesh parts beyond `far`
    both(fv, tex, 40, fill_back=False)
    both(fv, tex, 40, fill_back=False, aggr_func_rgb="hard")


@pytest.mark.parametrize("sigma,gamma,deps", [(1e-4, 1e-3, 1e-4), (3e-5, 1e-2, 1e-6)])
def test_other_softness_parameters(sigma, gamma, deps):
    fv, tex = scene(1, 2, 3, seed=6)
    both(fv, tex, 40, sigma_val=sigma, gamma_val=gamma, dist_eps=deps, background_color=(1.0, 0.5, 0.25))


def test_mirror_symmetric_mesh_depth_ties():
    """CUB meshes are mirror-symmetric: exact depth ties must resolve to the lowest face index (strict <)."""
    v, f = synth.icosphere(2)
    rng = np.random.default_rng(7)
    verts = (v * np.array([1.0, 0.6, 0.5]))[None].astype(np.float32)
    cams = np.array([[0.7, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]], dtype=np.float32)   # identity rotation: head-on
    fv = synth.raster_space_faces(verts, f, cams)
    tex = rng.uniform(0, 1, size=(1, f.shape[0], 4, 3)).astype(np.float32)
    both(fv, tex, 64, aggr_func_rgb="hard")
    both(fv, tex, 64)


def test_no_grad_forward_and_texture_only_grad():
    fv, tex = scene(1, 2, 2, seed=8)
    both(fv, tex, 32, grad=False)
    # gradient w.r.t. textures only (vertices detached), as MultiTextureLoss does (loss_utils.py:313)
    img, fwd, cfg = softras.render(fv, tex, 32, impl="B", sigma_val=1e-5, dist_eps=1e-10, gamma_val=1e-4)
    g = np.random.default_rng(4).normal(size=img.shape).astype(np.float32)
    _, gt = softras.render_backward(fwd, cfg, g, impl="B")
    ttex = torch.from_numpy(tex).to(DEV).requires_grad_(True)
    out, _, _ = raster.soft_rasterize(torch.from_numpy(fv).to(DEV), ttex, 32, anti_aliasing=True, sigma_val=1e-5,
                                      dist_eps=1e-10, gamma_val=1e-4)
    out.backward(torch.from_numpy(g).to(DEV))
    ok, msg = rel_report("grad_tex only", ttex.grad.cpu().numpy(), gt, 1e-4, 1e-6)
    assert ok, msg


@pytest.mark.parametrize("dist,alpha,textype,rgb", [
    ("barycentric", "sum", "surface", "softmax"),
    ("hard", "hard", "surface", "hard"),
    ("euclidean", "sum", "vertex", "softmax"),
    ("barycentric", "prod", "vertex", "hard"),
    ("hard", "prod", "surface", "softmax"),
])
def test_modes_umr_does_not_use(dist, alpha, textype, rgb):
    """the remaining soft_rasterize modes run through the generic kernel instantiations."""
    fv, tex = scene(2, 2, 2, seed=9)
    if textype == "vertex":
        tex = np.random.default_rng(10).uniform(0, 1, size=(2, fv.shape[1], 3, 3)).astype(np.float32)
    both(fv, tex, 40, dist_func=dist, aggr_func_alpha=alpha, texture_type=textype, aggr_func_rgb=rgb,
         sigma_val=1e-4, dist_eps=1e-4, gamma_val=1e-3)


def test_bad_mode_arguments_raise():
    fv, tex = scene(1, 1, 1)
    with pytest.raises(RuntimeError):   # vertex textures must be [B,F,3,3]
        raster.soft_rasterize(torch.from_numpy(fv).to(DEV), torch.from_numpy(tex).to(DEV), 16, texture_type="vertex")

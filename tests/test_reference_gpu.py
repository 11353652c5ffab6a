"""Three-way gate on the B200: our kernels vs THE REFERENCE'S OWN CUDA kernels, rebuilt
for sm_100a from /root/reference by baseline/build_ref_gpu.py (the built modules travel in
baseline/_ref/; skipped when they are absent).

* vs the -fmad=false build: the pixel planes (RGBA, softmax sum/max, hard depth/face-id) are BIT-EXACT --
  same IEEE operation sequence, same device expf; p2f / gradients differ only by float-atomics order.
* vs the default (FMA-contracted) build only the hard face-index plane is stable (App. B-15): at most a
  handful of mismatching pixels, the rest is reported by tools/ref_gpu_compare.py, not gated.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ref_gpu_compare as rc  # noqa: E402
from umr_b200 import raster  # noqa: E402

pytestmark = pytest.mark.gpu
KW = dict(sigma_val=1e-5, dist_eps=1e-10, gamma_val=1e-4, anti_aliasing=True)


def _mods():
    return rc.load("soft_rasterize_ref"), rc.load("soft_rasterize_ref_nofma")


@pytest.mark.skipif(_mods()[1] is None, reason="baseline/_ref/soft_rasterize_ref_nofma.so not built")
@pytest.mark.parametrize("rgb_name,rgb", [("softmax", 1), ("hard", 0)])
@pytest.mark.parametrize("tex_res", [1, 3])
def test_bit_exact_with_reference_cuda_kernels_built_without_fma(rgb_name, rgb, tex_res):
    mod = _mods()[1]
    IS, S = 128, 256
    fv, tex = rc.scene(2, tex_res, seed=5)
    a = fv.clone().requires_grad_(True)
    img, p2f, aggr = raster.soft_rasterize(a, tex, IS, aggr_func_rgb=rgb_name, **KW)
    g = torch.randn(img.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    img.backward(g)
    colors, rp2f, raggr, finfo = rc.ref_forward(mod, fv, tex, S, rgb)
    assert torch.equal(img.detach(), F.avg_pool2d(colors, 2, 2))
    assert torch.equal(aggr, raggr)
    assert torch.allclose(p2f, rp2f, rtol=1e-4, atol=1e-6)
    ghi = (g / 4).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    rgf, _ = rc.ref_backward(mod, fv, tex, colors, finfo, raggr, ghi, S, rgb)
    scale = float(rgf.abs().max())
    assert torch.allclose(a.grad, rgf, rtol=1e-3, atol=1e-5 * scale)


def _three_way(mod, fv, tex, IS, rgb_name, rgb, check_tex_grad):
    S = 2 * IS
    a = fv.clone().requires_grad_(True)
    t = tex.clone().requires_grad_(True)
    img, p2f, aggr = raster.soft_rasterize(a, t, IS, aggr_func_rgb=rgb_name, **KW)
    g = torch.randn(img.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    img.backward(g)
    colors, rp2f, raggr, finfo = rc.ref_forward(mod, fv, tex, S, rgb)
    assert torch.equal(img.detach(), F.avg_pool2d(colors, 2, 2)), "pooled RGBA not bit-exact"
    assert torch.equal(aggr, raggr), "aggregation planes not bit-exact"
    assert torch.allclose(p2f, rp2f, rtol=1e-4, atol=1e-6)
    ghi = (g / 4).repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    rgf, rgt = rc.ref_backward(mod, fv, tex, colors, finfo, raggr, ghi, S, rgb)
    scale = float(rgf.abs().max())
    assert torch.allclose(a.grad, rgf, rtol=1e-3, atol=1e-5 * scale)
    if check_tex_grad:  # T^2 == 1: the reference's texel-gradient UB (App. B-1) coincides with the intended semantics
        assert torch.allclose(t.grad, rgt, rtol=1e-3, atol=1e-5 * float(rgt.abs().max()))


@pytest.mark.skipif(_mods()[1] is None, reason="baseline/_ref/soft_rasterize_ref_nofma.so not built")
@pytest.mark.parametrize("tile", [16, 32])
@pytest.mark.parametrize("rgb_name,rgb", [("softmax", 1), ("hard", 0)])
def test_bit_exact_at_c2_shape(rgb_name, rgb, tile, monkeypatch):
    """BASELINE config 2 per-image shape: F=1280, 256^2 (S=512), T^2=36, B=2 -- through both forward kernels."""
    monkeypatch.setattr(raster, "FORWARD_TILE", tile)
    fv, tex = rc.scene(2, 6, seed=3)
    _three_way(_mods()[1], fv, tex, 256, rgb_name, rgb, False)


@pytest.mark.skipif(_mods()[1] is None, reason="baseline/_ref/soft_rasterize_ref_nofma.so not built")
@pytest.mark.parametrize("rgb_name,rgb,tex_res", [("softmax", 1, 1), ("hard", 0, 2)])
def test_bit_exact_at_c5_shape(rgb_name, rgb, tex_res):
    """BASELINE config 5 per-image shape: F=5120 (icosphere subdiv 4), 1024^2 (S=2048), B=1 -- the only shape
    where the cull boxes span several staging pieces and tiles carry long face lists."""
    fv, tex = rc.scene(1, tex_res, seed=9, subdiv=4)
    assert fv.shape[1] == 5120
    _three_way(_mods()[1], fv, tex, 1024, rgb_name, rgb, tex_res == 1)


@pytest.mark.skipif(_mods()[0] is None, reason="baseline/_ref/soft_rasterize_ref.so not built")
def test_face_index_plane_against_reference_as_normally_compiled():
    mod = _mods()[0]
    IS, S = 128, 256
    fv, tex = rc.scene(2, 2, seed=6)
    _, _, aggr = raster.soft_rasterize(fv, tex, IS, aggr_func_rgb="hard", **KW)
    _, _, raggr, _ = rc.ref_forward(mod, fv, tex, S, 0)
    mism = int((aggr[:, 1] != raggr[:, 1]).sum())
    print("face-id mismatches vs FMA-compiled reference: %d / %d" % (mism, aggr[:, 1].numel()))
    assert mism <= 8  # exact ties / sliver faces only (App. B-15)

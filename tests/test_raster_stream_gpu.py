"""Round-2 raster pipeline (csrc/raster_stream.cuh): the backward that STREAMS the forward's saved pair records
must agree with (a) the CPU oracle, (b) the round-1 recompute backward (no pair buffer), and (c) itself when the
pair buffer is too small and some / all tiles fall back to the recompute kernel."""
import numpy as np
import pytest
import torch

from umr_b200 import raster
from util import rel_report, scene
from test_raster_gpu import UMR, run_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(fv, tex, image_size, rgb, g, cand_per_pixel, tile=0):
    old, old_tile, old_ad = raster.PAIR_CAND_PER_PIXEL, raster.FORWARD_TILE, raster.PAIR_ADAPTIVE
    raster.PAIR_CAND_PER_PIXEL = cand_per_pixel
    raster.FORWARD_TILE = tile
    raster.PAIR_ADAPTIVE = False   # fixed budgets here; the adaptive sizing has its own test below
    try:
        tfv = torch.from_numpy(fv).to(DEV).requires_grad_(True)
        ttex = torch.from_numpy(tex).to(DEV).requires_grad_(True)
        img, p2f, aggr = raster.soft_rasterize(tfv, ttex, image_size, aggr_func_rgb=rgb, anti_aliasing=True, **UMR)
        stats = None
        saved = img.grad_fn.saved_tensors
        if len(saved) == 5:
            stats = saved[4][:8].view(torch.int32).cpu().tolist()  # [blocks wanted, tiles unsaved]
        img.backward(torch.from_numpy(g).to(DEV))
        torch.cuda.synchronize()
        return dict(images=img.detach().cpu().numpy(), grad_faces=tfv.grad.cpu().numpy(), grad_tex=ttex.grad.cpu().numpy(),
                    stats=stats)
    finally:
        raster.PAIR_CAND_PER_PIXEL, raster.FORWARD_TILE, raster.PAIR_ADAPTIVE = old, old_tile, old_ad


@pytest.mark.parametrize("tile", [16, 32])   # k_raster_fwd3 (16x16, static warps) / k_raster_fwd4 (32x32, dynamic pixel blocks)
@pytest.mark.parametrize("rgb", ["softmax", "hard"])
# (2, 1, 1, 64): 80 large faces with one texel each (205 raster pixels per texel) -> the warp-level texel pre-reduction
# (k_raster_bwd2<..., PRE>) is on
@pytest.mark.parametrize("B,subdiv,tex_res,image_size", [(2, 3, 3, 64), (1, 3, 6, 128), (1, 2, 2, 37), (2, 1, 1, 64)])
def test_streamed_backward_matches_recompute_and_oracle(tile, rgb, B, subdiv, tex_res, image_size):
    fv, tex = scene(B, subdiv, tex_res, seed=31 + image_size)
    g = np.random.default_rng(5).normal(size=(B, 4, image_size, image_size)).astype(np.float32)
    ref = run_oracle(fv, tex, image_size, True, rgb, g)
    full = _run(fv, tex, image_size, rgb, g, 32.0, tile)    # everything saved
    none = _run(fv, tex, image_size, rgb, g, 0.0, tile)     # no pair buffer: recompute backward
    part = _run(fv, tex, image_size, rgb, g, 0.7, tile)     # buffer too small: some tiles saved, the rest recomputed
    tiny = _run(fv, tex, image_size, rgb, g, 1e-4, tile)    # nothing fits
    assert full["stats"][1] == 0 and full["stats"][0] > 0, full["stats"]
    assert none["stats"] is None
    assert part["stats"][1] > 0, "the partial-capacity case must leave tiles unsaved: %s" % (part["stats"],)
    assert tiny["stats"][1] > 0
    at = 1e-6 * float(np.abs(ref["grad_faces"]).max() + 1e-30) + 1e-7
    for name, got in (("streamed", full), ("recompute", none), ("partial", part), ("tiny", tiny)):
        assert np.array_equal(got["images"], full["images"])
        for k, a in (("grad_faces", at), ("grad_tex", 1e-6)):
            ok, msg = rel_report("%s %s" % (name, k), got[k], ref[k], 1e-4, a)
            print(msg)
            assert ok, msg


def test_forward_only_render_needs_no_pair_buffer():
    fv, tex = scene(1, 2, 1, seed=2)
    img, _, _ = raster.soft_rasterize(torch.from_numpy(fv).to(DEV), torch.from_numpy(tex).to(DEV), 32, anti_aliasing=True, **UMR)
    assert img.grad_fn is None


@pytest.mark.parametrize("rgb", ["softmax", "hard"])
@pytest.mark.parametrize("cand", [32.0, 0.0])
def test_batch_shared_texture_equals_repeated_copies(rgb, cand):
    """A [1,F,T2,3] texture is a batch-shared parameter (no repeat(B) copies, loss_utils.py:305): same images as the
    materialised copies, and its gradient is the batch sum of the per-image gradients."""
    B, IS = 3, 48
    fv, tex = scene(B, 2, 3, seed=12)
    tex1 = tex[:1]
    g = np.random.default_rng(9).normal(size=(B, 4, IS, IS)).astype(np.float32)
    old = raster.PAIR_CAND_PER_PIXEL
    raster.PAIR_CAND_PER_PIXEL = cand
    try:
        outs = []
        for shared in (True, False):
            tfv = torch.from_numpy(fv).to(DEV).requires_grad_(True)
            t = torch.from_numpy(tex1).to(DEV).requires_grad_(True)
            tin = t if shared else t.expand(B, -1, -1, -1)
            img, p2f, aggr = raster.soft_rasterize(tfv, tin, IS, aggr_func_rgb=rgb, anti_aliasing=True, **UMR)
            img.backward(torch.from_numpy(g).to(DEV))
            outs.append((img.detach().cpu().numpy(), tfv.grad.cpu().numpy(), t.grad.cpu().numpy()))
    finally:
        raster.PAIR_CAND_PER_PIXEL = old
    assert np.array_equal(outs[0][0], outs[1][0])
    ok, msg = rel_report("grad_faces", outs[0][1], outs[1][1], 1e-4, 1e-6 * float(np.abs(outs[1][1]).max()) + 1e-7)
    assert ok, msg
    assert outs[0][2].shape == (1,) + tex.shape[1:]
    ok, msg = rel_report("grad_tex (batch sum)", outs[0][2], outs[1][2], 1e-4, 1e-6 * float(np.abs(outs[1][2]).max()) + 1e-7)
    assert ok, msg


@pytest.mark.parametrize("rgb", ["softmax", "hard"])
def test_dense_mesh_takes_the_windowed_slow_path(rgb):
    """5120 faces on a 64x64 raster: the single coarse bin lists every face -- longer than the 32x32-tile kernel's shared
    tile list -- so its windowed static path runs (and leaves the tile to the recompute backward); results must not change."""
    fv, tex = scene(1, 4, 1, seed=8)
    assert fv.shape[1] == 5120
    g = np.random.default_rng(6).normal(size=(1, 4, 32, 32)).astype(np.float32)
    ref = run_oracle(fv, tex, 32, True, rgb, g)
    a = _run(fv, tex, 32, rgb, g, 64.0, 32)
    b = _run(fv, tex, 32, rgb, g, 64.0, 16)
    assert a["stats"][1] > 0, "expected the dense tile to be left unsaved by the 32x32-tile kernel: %s" % (a["stats"],)
    assert np.array_equal(a["images"], b["images"])
    ok, msg = rel_report("images", a["images"], ref["images"], 1e-4, 1e-6)
    assert ok, msg
    at = 1e-6 * float(np.abs(ref["grad_faces"]).max() + 1e-30) + 1e-7
    for got in (a, b):
        ok, msg = rel_report("grad_faces", got["grad_faces"], ref["grad_faces"], 1e-4, at)
        assert ok, msg


def test_pair_buffer_sizing_adapts_to_the_measured_need(monkeypatch):
    """The buffer's own counters are read back asynchronously; later renders of the same (raster size, face count) get a
    buffer sized from the measured need instead of the fixed budget, and a render that outgrows it only recomputes tiles."""
    monkeypatch.setattr(raster, "PAIR_ADAPTIVE", True)
    monkeypatch.setattr(raster, "PAIR_CAND_PER_PIXEL", 32.0)
    monkeypatch.setattr(raster, "FORWARD_TILE", 0)
    raster._pair_need.clear()
    del raster._pair_pending[:]
    fv, tex = scene(2, 3, 2, seed=41)
    g = torch.ones(2, 4, 64, 64, device=DEV)

    def render(f):
        tfv = torch.from_numpy(f).to(DEV).requires_grad_(True)
        img, _, _ = raster.soft_rasterize(tfv, torch.from_numpy(tex).to(DEV), 64, anti_aliasing=True, **UMR)
        pairs = img.grad_fn.saved_tensors[4]
        img.backward(g)
        torch.cuda.synchronize()
        return pairs.numel(), pairs[:8].view(torch.int32).cpu().tolist(), tfv.grad.cpu().numpy()

    n0, st0, g0 = render(fv)                  # fixed budget
    n1, st1, g1 = render(fv)                  # sized from the first call's counters
    assert st0[1] == 0 and st1[1] == 0
    assert n1 < 0.5 * n0, (n0, n1)
    assert (128, fv.shape[1], 0) in raster._pair_need
    big = fv.copy()
    big[..., [0, 1, 3, 4, 6, 7]] *= 1.6       # a much larger silhouette: outgrows the adapted buffer
    n2, st2, g2 = render(big)
    n3, st3, g3 = render(big)
    assert st3[1] == 0 and n3 > n2 * 0.99     # grown after the overflow was observed
    ok, msg = rel_report("grad (overflowing vs grown buffer)", g2, g3, 1e-4, 1e-6 * float(np.abs(g3).max()) + 1e-7)
    assert ok, msg


@pytest.mark.parametrize("cand", [32.0, 0.7, 0.0])   # all pairs saved / some tiles recomputed / no pair buffer at all
@pytest.mark.parametrize("B,subdiv,tex_res,image_size,aa", [(2, 3, 3, 64, True), (1, 2, 2, 37, True), (2, 2, 2, 48, False)])
def test_four_colour_channels_equal_two_rgb_renders(cand, B, subdiv, tex_res, image_size, aa):
    """textures [B,F,T2,4] render all four part maps of part_matching_loss (loss_utils.py:385-399)
    in ONE launch.  Colour channels never interact in the rasteriser, so planes 0-3 must be BIT-identical to the
    same channels rendered through the 3-channel kernels, alpha identical, and the vertex gradient equal to the sum
    of the two 3-channel renders' gradients (each checked against the CPU oracle elsewhere)."""
    fv, tex3 = scene(B, subdiv, tex_res, seed=77 + image_size)
    rng = np.random.default_rng(9)
    extra = rng.uniform(0, 1, size=tex3.shape[:3] + (1,)).astype(np.float32)
    tex4 = np.concatenate([tex3, extra], axis=-1)
    texb = np.concatenate([extra, np.zeros_like(tex3[..., :2])], axis=-1)   # 4th channel rides in R of a second render
    g = rng.normal(size=(B, 5, image_size, image_size)).astype(np.float32)
    old, old_ad = raster.PAIR_CAND_PER_PIXEL, raster.PAIR_ADAPTIVE
    raster.PAIR_CAND_PER_PIXEL, raster.PAIR_ADAPTIVE = cand, False
    try:
        def render(tex, gimg):
            tfv = torch.from_numpy(fv).to(DEV).requires_grad_(True)
            img, p2f, aggr = raster.soft_rasterize(tfv, torch.from_numpy(tex).to(DEV), image_size, aggr_func_rgb="softmax",
                                                   anti_aliasing=aa, **UMR)
            img.backward(torch.from_numpy(gimg).to(DEV))
            return img.detach().cpu().numpy(), p2f.cpu().numpy(), aggr.cpu().numpy(), tfv.grad.cpu().numpy()
        i4, p4, a4, gf4 = render(tex4, g)
        ga = np.ascontiguousarray(g[:, [0, 1, 2, 4]])                       # RGB + alpha gradient
        gb = np.zeros_like(ga); gb[:, 0] = g[:, 3]                          # 4th channel, no alpha gradient (counted once)
        ia, pa, aa_, gfa = render(tex3, ga)
        ib, _, _, gfb = render(texb, gb)
    finally:
        raster.PAIR_CAND_PER_PIXEL, raster.PAIR_ADAPTIVE = old, old_ad
    assert i4.shape == (B, 5, image_size, image_size)
    assert np.array_equal(i4[:, 0:3], ia[:, 0:3]) and np.array_equal(i4[:, 3], ib[:, 0]) and np.array_equal(i4[:, 4], ia[:, 3])
    assert np.array_equal(a4, aa_)
    np.testing.assert_allclose(p4, pa, rtol=1e-5, atol=1e-6)   # p2f sums are float REDs: order varies run to run
    want = gfa + gfb
    ok, msg = rel_report("grad_faces 4ch", gf4, want, 2e-4, 2e-6 * float(np.abs(want).max()) + 1e-7)
    print(msg)
    assert ok, msg


def test_four_channel_mode_rejects_what_it_was_not_built_for():
    fv, tex3 = scene(1, 2, 2, seed=1)
    tex4 = np.concatenate([tex3, tex3[..., :1]], axis=-1)
    tfv = torch.from_numpy(fv).to(DEV).requires_grad_(True)
    with pytest.raises(ValueError):   # texture gradient
        raster.soft_rasterize(tfv, torch.from_numpy(tex4).to(DEV).requires_grad_(True), 32, anti_aliasing=True, **UMR)
    with pytest.raises(Exception):    # hard colour aggregation: UMR_ERR_UNSUPPORTED from the library
        raster.soft_rasterize(tfv, torch.from_numpy(tex4).to(DEV), 32, aggr_func_rgb="hard", anti_aliasing=True, **UMR)


@pytest.mark.parametrize("tile", [16, 32])
@pytest.mark.parametrize("rgb", ["softmax", "hard"])
@pytest.mark.parametrize("cand", [32.0, 0.7])
@pytest.mark.parametrize("subdiv,tex_res", [(3, 3), (1, 1)])   # (1, 1): large faces -> with the warp-level texel pre-reduction
def test_texture_only_backward_for_detached_geometry(tile, rgb, cand, subdiv, tex_res):
    """UMR's texture branch renders DETACHED vertices / cameras (experiments/train_s2.py:248): the backward then forms
    only the texel gradients (k_raster_bwd2<..., GEOM = false>) -- same values as the full backward, no grad_faces."""
    B, image_size = 2, 64
    fv, tex = scene(B, subdiv, tex_res, seed=5)
    g = np.random.default_rng(6).normal(size=(B, 4, image_size, image_size)).astype(np.float32)
    full = _run(fv, tex, image_size, rgb, g, cand, tile)
    old, old_tile, old_ad = raster.PAIR_CAND_PER_PIXEL, raster.FORWARD_TILE, raster.PAIR_ADAPTIVE
    raster.PAIR_CAND_PER_PIXEL, raster.FORWARD_TILE, raster.PAIR_ADAPTIVE = cand, tile, False
    try:
        tfv = torch.from_numpy(fv).to(DEV)                       # no grad
        ttex = torch.from_numpy(tex).to(DEV).requires_grad_(True)
        img, _, _ = raster.soft_rasterize(tfv, ttex, image_size, aggr_func_rgb=rgb, anti_aliasing=True, **UMR)
        img.backward(torch.from_numpy(g).to(DEV))
    finally:
        raster.PAIR_CAND_PER_PIXEL, raster.FORWARD_TILE, raster.PAIR_ADAPTIVE = old, old_tile, old_ad
    assert np.array_equal(img.detach().cpu().numpy(), full["images"])
    ok, msg = rel_report("grad_tex (texture-only)", ttex.grad.cpu().numpy(), full["grad_tex"], 1e-5,
                         1e-6 * float(np.abs(full["grad_tex"]).max()) + 1e-7)
    print(msg)
    assert ok, msg


@pytest.mark.parametrize("B,subdiv,tex_res,image_size,aa", [(2, 3, 3, 64, True), (1, 2, 2, 37, True), (2, 2, 2, 48, False),
                                                            (1, 4, 1, 256, True)])
def test_visibility_only_kernel_equals_the_hard_render(B, subdiv, tex_res, image_size, aa):
    """k_raster_fwd3<2>: the z-buffer winner per pixel without the image.  Both aggrs planes (depth_min, face_index_min)
    must be BIT-identical to the full hard render's -- MultiTextureLoss keeps nothing else of it (loss_utils.py:327-329)."""
    fv, tex = scene(B, subdiv, tex_res, seed=13 + image_size)
    tfv = torch.from_numpy(fv).to(DEV)
    _, _, aggr = raster.soft_rasterize(tfv, torch.from_numpy(tex).to(DEV), image_size, aggr_func_rgb="hard", anti_aliasing=aa, **UMR)
    vis = raster.visibility(tfv, image_size, anti_aliasing=aa, **UMR)
    assert vis.shape == aggr.shape
    assert torch.equal(vis, aggr)
    assert float((vis[:, 1] >= 0).float().mean()) > 0.02   # the mesh is visible somewhere


@pytest.mark.parametrize("slivers", [False, True])
@pytest.mark.parametrize("B,subdiv,image_size,aa", [(2, 3, 64, True), (1, 2, 37, True), (2, 2, 48, False), (2, 3, 512, True),
                                                    (1, 4, 256, True)])
def test_visible_face_bytes_equal_the_set_in_the_face_index_plane(B, subdiv, image_size, aa, slivers):
    """want_faces (k_visible_faces: face-parallel z-buffer per 64x64 bin): the [B,F] bytes are exactly the faces present in
    the hard render's face-index plane, with the reference's quirk that a background pixel (-1) marks face F-1
    (loss_utils.py:161-166 index with the raw plane).  `slivers` flattens every 7th face to a needle / a zero-area
    triangle: those take the whole-cull-box scan (R_FLG bit 4)."""
    fv, tex = scene(B, subdiv, 2, seed=3 + image_size)
    if slivers:
        fv = fv.copy().reshape(B, -1, 3, 3)
        t = np.float32(1e-6)
        fv[:, ::7, 2, :2] = (fv[:, ::7, 0, :2] + fv[:, ::7, 1, :2]) * np.float32(0.5) + t       # needle
        fv[:, ::21, 2] = fv[:, ::21, 1]                                                         # two coincident corners
        fv = np.ascontiguousarray(fv.reshape(B, -1, 9))
    tfv = torch.from_numpy(fv).to(DEV)
    F = fv.shape[1]
    _, _, aggr = raster.soft_rasterize(tfv, torch.from_numpy(tex).to(DEV), image_size, aggr_func_rgb="hard", anti_aliasing=aa, **UMR)
    mask = raster.visibility(tfv, image_size, anti_aliasing=aa, want_faces=True, **UMR)
    assert mask.dtype == torch.uint8 and tuple(mask.shape) == (B, F)
    want = torch.zeros(B, F, dtype=torch.uint8)
    for b in range(B):
        ids = torch.unique(aggr[b, 1].long().cpu())
        want[b, ids] = 1          # -1 -> F-1, like the reference's indexing
    diff = (mask.cpu() != want).nonzero()
    assert diff.numel() == 0, "faces differing: %s" % diff[:10].tolist()
    assert 0 < int(want.sum()) < B * F

"""CPU tests of the host-side logic: synthetic generators, the drop-in package's torch glue, the
mesh regularisers, and CPU-tensor rejection (no CPU fallback)."""
import numpy as np
import pytest
import torch

import losses as oracle_losses
from umr_b200 import soft_renderer as sr
from umr_b200 import synth
from umr_b200.dist import shard_range
from umr_b200.nnutils import geom_utils, loss_utils, smr


def test_icosphere_counts():
    for sd, (nv, nf) in {0: (12, 20), 2: (162, 320), 3: (642, 1280), 4: (2562, 5120)}.items():
        v, f = synth.icosphere(sd)
        assert v.shape == (nv, 3) and f.shape == (nf, 3)
        assert np.allclose(np.linalg.norm(v, axis=1), 1, atol=1e-6)
        assert f.min() == 0 and f.max() == nv - 1
        # closed manifold: every undirected edge is shared by exactly two faces
        e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
        _, cnt = np.unique(e, axis=0, return_counts=True)
        assert (cnt == 2).all()


def test_camera_biases_are_unit_quaternions():
    q = synth.camera_biases()
    assert q.shape == (8, 4) and np.allclose(np.linalg.norm(q, axis=1), 1, atol=2e-4)


def test_projection_glue_matches_reference_restatement():
    rng = np.random.default_rng(0)
    X = torch.from_numpy(rng.normal(size=(3, 50, 3)).astype(np.float32))
    cam = torch.from_numpy(synth.cameras(rng, 3))
    a = geom_utils.orthographic_proj_withz(X, cam, offset_z=5.)
    b = oracle_losses.orthographic_proj_withz(X, cam, offset_z=5.)
    assert torch.allclose(a, b, atol=1e-6)
    assert torch.allclose(geom_utils.orthographic_proj(X, cam), b[:, :, :2] , atol=1e-6)
    # numpy generator used by the benches agrees too
    fv = synth.raster_space_faces(X.numpy(), np.array([[0, 1, 2]], dtype=np.int32), cam.numpy())
    ref = b.clone()
    ref[:, :, 1] *= -1
    ref[:, :, 2] += 2.732
    assert np.allclose(fv[:, 0].reshape(3, 3, 3), ref[:, :3].numpy(), atol=2e-6)


def test_mesh_lighting_transform_on_cpu():
    v, f = synth.icosphere(1)
    B = 2
    verts = torch.from_numpy(np.stack([v, v * 0.5]))
    faces = torch.from_numpy(f)[None].repeat(B, 1, 1)
    tex = torch.rand(B, f.shape[0], 4, 3)
    mesh = sr.Mesh(verts, faces, tex)
    assert mesh.face_vertices.shape == (B, f.shape[0], 3, 3) and mesh.texture_res == 2
    assert torch.equal(mesh.face_vertices[1, 5, 2], verts[1, f[5, 2]])
    n = mesh.surface_normals
    assert torch.allclose(n.norm(dim=2), torch.ones(B, f.shape[0]), atol=1e-5)
    light = sr.Lighting("surface", 0.8, (1, 1, 1), 0.5, (1, 1, 1), (0, 1, 0))
    lit = light(sr.Mesh(verts, faces, tex.clone()))
    expect = tex * (0.8 + 0.5 * torch.relu(n[:, :, 1]))[:, :, None, None]
    assert torch.allclose(lit.textures, expect, atol=1e-6)
    tr = sr.Transform("look_at", perspective=False, eye=[0, 0, -2.732])
    out = tr(sr.Mesh(verts.clone(), faces)).vertices
    assert torch.allclose(out, verts + torch.tensor([0, 0, 2.732]), atol=1e-6)
    with pytest.raises(ValueError):
        sr.Transform("projection")      # needs a [B,3,4] matrix, like the reference
    no_tex = sr.Mesh(verts, faces)
    assert no_tex.textures.shape == (B, f.shape[0], 1, 3) and float(no_tex.textures.min()) == 1.0


def test_soft_renderer_attribute_paths_and_cpu_rejection():
    r = smr.SoftRenderer(32, "softmax")
    assert r.renderer.transform.transformer._eye == [0, 0, -2.732]
    assert r.renderer.lighting.ambient.light_intensity == 0.8
    r.ambient_light_only()
    assert r.renderer.lighting.directionals[0].light_intensity == 0
    r.set_bgcolor([1, 1, 1])
    assert r.renderer.rasterizer.background_color == [1, 1, 1]
    assert r.renderer.rasterizer.anti_aliasing and r.renderer.rasterizer.dist_eps == 1e-10
    v, f = synth.icosphere(0)
    with pytest.raises(TypeError):  # no CPU fallback (reference: soft_rasterize.py:117-118)
        r(torch.from_numpy(v)[None], torch.from_numpy(f.astype(np.int64))[None], torch.tensor([[1., 0, 0, 1, 0, 0, 0]]))


def test_laplacian_and_flatten_losses():
    v, f = synth.icosphere(1)
    V, Fc = torch.from_numpy(v), torch.from_numpy(f.astype(np.int64))
    lap = sr.LaplacianLoss(V, Fc)
    x = torch.from_numpy(v)[None]
    # brute-force restatement of losses.py:12-27
    L = np.zeros((len(v), len(v)), np.float32)
    for a, b, c in f:
        for i, j in ((a, b), (b, c), (c, a)):
            L[i, j] = L[j, i] = -1
    L[np.arange(len(v)), np.arange(len(v))] = -L.sum(1)
    L = L / np.diag(L)[:, None]
    ref = ((L @ v) ** 2).sum()
    assert abs(float(lap(x)[0]) - ref) < 1e-4 * max(ref, 1)
    fl = sr.FlattenLoss(Fc)
    assert fl.v0s.numel() == 3 * len(f) // 2               # one entry per edge of a closed mesh
    flat = float(fl(x)[0])
    bumpy = float(fl(x + 0.05 * torch.randn(x.shape, generator=torch.Generator().manual_seed(0)))[0])
    assert 0 <= flat < bumpy                                # smoother surface => smaller dihedral penalty
    # brute force over edges for one configuration
    xv = x[0].numpy().astype(np.float64)
    tot = 0.0
    for a, b, c, d in zip(fl.v0s.tolist(), fl.v1s.tolist(), fl.v2s.tolist(), fl.v3s.tolist()):
        e = xv[b] - xv[a]
        def perp(p):
            q = xv[p] - xv[a]
            return q - e * (q @ e) / (e @ e)
        p1, p2 = perp(c), perp(d)
        tot += (p1 @ p2 / (np.linalg.norm(p1) * np.linalg.norm(p2)) + 1) ** 2
    assert abs(flat - tot) < 1e-2 * max(tot, 1e-3) + 1e-3


def test_batch_get_centers_matches_reference_loops():
    g = torch.Generator().manual_seed(1)
    p = torch.softmax(torch.randn(2, 4, 16, 16, generator=g), 1)
    got = loss_utils.batch_get_centers(p)
    # scops_utils.py:21-54 restated with loops
    xs = np.tile(np.arange(16), (16, 1)) / 16 * 2 - 1.0
    ys = xs.T
    for b in range(2):
        for c in range(4):
            m = p[b, c].numpy() + 1e-3
            m = m / m.sum()
            assert abs(float(got[b, c, 0]) - (m * xs).sum()) < 1e-5
            assert abs(float(got[b, c, 1]) - (m * ys).sum()) < 1e-5


def test_shard_range_partitions():
    for n in (16, 17, 128, 3):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_compat_install_and_overlay(tmp_path):
    import importlib
    import sys
    from umr_b200 import compat
    compat.install()
    import soft_renderer as sr2
    assert sr2.SoftRenderer is sr.SoftRenderer and hasattr(sr2, "LaplacianLoss") and hasattr(sr2.functional, "soft_rasterize")
    # fake reference checkout
    ref = tmp_path / "ref"
    for d in ("experiments", "data", "utils", "nnutils"):
        (ref / d).mkdir(parents=True)
        (ref / d / "__init__.py").write_text("")
    (ref / "nnutils" / "scops_utils.py").write_text("MARK = 'reference'\n")
    (ref / "nnutils" / "smr.py").write_text("MARK = 'reference smr'\n")
    (ref / "nnutils" / "geom_utils.py").write_text(
        "MARK = 'reference geom'\n"
        "def sample_textures(f, i):\n    return 'reference sampler'\n"
        "def orthographic_proj_withz(X, cam, offset_z=0.):\n    return 'reference proj'\n"
        "def rotate_cam(c, a):\n    return 'reference rotate_cam'\n")
    (ref / "nnutils" / "perceptual_loss.py").write_text(
        "class PerceptualLoss(object):\n"
        "    def __call__(self, a, b):\n        return (a - b).abs().mean(dim=(1, 2, 3))\n")
    base = compat.overlay(str(ref), package="UMRT", workdir=str(tmp_path / "ov"))
    try:
        m = importlib.import_module("UMRT.nnutils.smr")
        assert m.SoftRenderer is smr.SoftRenderer            # ours
        assert importlib.import_module("UMRT.nnutils.scops_utils").MARK == "reference"   # theirs, via symlink
        assert importlib.import_module("UMRT.nnutils.chamfer_python").distChamfer is not None
        # geom_utils: the reference module with the hot functions replaced (one sampling convention)
        g = importlib.import_module("UMRT.nnutils.geom_utils")
        from umr_b200.nnutils import geom_utils as ours_geom
        assert g.MARK == "reference geom" and g.rotate_cam(None, None) == "reference rotate_cam"
        assert g.orthographic_proj_withz is ours_geom.orthographic_proj_withz and g.quat_rotate is ours_geom.quat_rotate
        assert g.sample_textures(torch.zeros(1, 1, 1, 1, 2), torch.zeros(1, 1, 2, 2)) == "reference sampler"  # CPU tensors
        # loss_utils under the overlay: default MultiTextureLoss = perceptual through the REFERENCE's LPIPS module
        lu = importlib.import_module("UMRT.nnutils.loss_utils")
        ptl = lu.PerceptualTextureLoss()
        d = ptl(torch.ones(2, 3, 4, 4), torch.zeros(2, 3, 4, 4), torch.ones(2, 4, 4), torch.ones(2, 4, 4), avg=False)
        assert d.shape == (2,) and torch.allclose(d, torch.ones(2))
        assert float(lu.entropy_loss(torch.full((3, 4), 0.25))) == pytest.approx(float(np.log(4.0)))
    finally:
        sys.path.remove(base)
        for k in [k for k in sys.modules if k.startswith("UMRT")]:
            del sys.modules[k]


def test_perceptual_texture_loss_fails_loudly_without_the_reference_module():
    """MultiTextureLoss keeps the reference default 'perceptual' and must never fall back to L1
    silently; without the reference's LPIPS module the constructor raises."""
    import inspect
    assert inspect.signature(loss_utils.MultiTextureLoss.__init__).parameters["texture_loss_type"].default == "perceptual"
    with pytest.raises(NotImplementedError, match="perceptual"):
        loss_utils.PerceptualTextureLoss()


def test_other_camera_modes_and_vertex_normals():
    """Generic torch path of the drop-in package for the modes UMR does not use."""
    v, f = synth.icosphere(1)
    verts = torch.from_numpy(v)[None].repeat(2, 1, 1)
    faces = torch.from_numpy(f)[None].repeat(2, 1, 1)
    # sphere: area-weighted vertex normals point radially outwards (up to the mesh's winding sign)
    n = sr.functional.vertex_normals(verts, faces)
    cosang = (n * verts).sum(-1)
    assert torch.allclose(cosang.abs(), torch.ones_like(cosang), atol=1e-2) and (cosang > 0).all() or (cosang < 0).all()
    # look == look_at when the direction points at the origin
    eye = [0.0, 0.0, -2.5]
    a = sr.functional.look(verts, eye, direction=[0, 0, 1], up=[0, 1, 0])
    b = sr.functional.look_at(verts, eye)
    assert torch.allclose(a, b, atol=1e-6)
    # perspective: x / z / tan(angle)
    p = sr.functional.perspective(b, angle=30.)
    assert torch.allclose(p[..., 0], b[..., 0] / b[..., 2] / np.tan(np.pi / 6), atol=1e-6)
    # projection with identity intrinsics and no distortion: (x/z, y/z) mapped from [0, size] to [-1, 1]
    P = torch.eye(3, 4)[None].repeat(2, 1, 1)
    q = sr.functional.projection(b, P, torch.zeros(2, 5), orig_size=2.0)
    assert torch.allclose(q[..., 0], 2 * (b[..., 0] / (b[..., 2] + 1e-5) - 1.0) / 2.0, atol=1e-5)
    t = sr.Transform("projection", P=P, orig_size=2.0)
    assert torch.allclose(t(sr.Mesh(b.clone(), faces)).vertices, q, atol=1e-6)
    t2 = sr.Transform("look", perspective=False, eye=eye)
    assert torch.allclose(t2(sr.Mesh(verts.clone(), faces)).vertices[..., 2], verts[..., 2] + 2.5, atol=1e-6)
    t2.set_eyes_from_angles(2.732, 0.0, 0.0)
    assert np.allclose(t2.transformer._eye, (0.0, 0.0, -2.732), atol=1e-6)
    # vertex lighting: colours scale with ambient + directional * relu(n . d)
    tex = torch.rand(2, v.shape[0], 3)
    mesh = sr.Mesh(verts, faces, tex.clone(), texture_type="vertex")
    lit = sr.Lighting("vertex", 0.5, (1, 1, 1), 0.5, (1, 1, 1), (0, 1, 0))(mesh)
    expect = tex * (0.5 + 0.5 * torch.relu(n[..., 1]))[..., None]
    assert torch.allclose(lit.textures, expect, atol=1e-6)
    assert mesh.face_textures.shape == (2, f.shape[0], 3, 3)


def test_drop_in_constructor_tables_accept_the_reference_spellings(monkeypatch):
    """The table-driven constructors accept the reference's positional order and keywords, expose every
    field as an attribute, and hand the rasteriser exactly the argument tuple soft_rasterize expects."""
    from umr_b200.soft_renderer import rasterizer as rz
    # UMR's own construction (nnutils/smr.py:56)
    r = sr.SoftRenderer(image_size=64, aggr_func_rgb="hard", camera_mode="look_at", sigma_val=1e-5, dist_eps=1e-10,
                        gamma_val=1e-4, background_color=[0, 0, 0], anti_aliasing=True, perspective=False)
    ras = r.rasterizer
    got = {k: getattr(ras, k) for k, _ in rz.FIELDS}
    assert got == dict(image_size=64, background_color=[0, 0, 0], near=1, far=100, anti_aliasing=True, fill_back=True,
                       eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-10, gamma_val=1e-4,
                       aggr_func_rgb="hard", aggr_func_alpha="prod", texture_type="surface")
    assert r.transform.camera_mode == "look_at" and r.transform.transformer.perspective is False
    assert r.transform.transformer.viewing_scale == 1.0 and r.transform.transformer.viewing_angle == 30
    assert r.lighting.ambient.light_intensity == 0.5 and r.lighting.directionals[0].light_direction == (0, 1, 0)
    # positional order of the reference: SoftRasterizer(image_size, background_color, near, far, anti_aliasing, fill_back, eps, ...)
    p = sr.SoftRasterizer(32, (1, 1, 1), 2, 50, True, True, 1e-2)
    assert (p.image_size, p.background_color, p.near, p.far, p.anti_aliasing, p.fill_back, p.eps) == (32, (1, 1, 1), 2, 50, True, True, 1e-2)
    with pytest.raises(ValueError):
        sr.SoftRasterizer(dist_func="manhattan")
    with pytest.raises(TypeError):
        sr.SoftRasterizer(image_sise=3)
    with pytest.raises(TypeError):
        sr.SoftRasterizer(32, image_size=64)
    # the kernel call receives (fv, tex, image_size, bg, near, far, fill_back, eps, sigma, dist_func, dist_eps, gamma,
    # aggr_rgb, aggr_alpha, texture_type, anti_aliasing) -- the signature of umr_b200.raster.soft_rasterize
    seen = {}
    monkeypatch.setattr(rz, "soft_rasterize", lambda *a: seen.setdefault("args", a))
    ras.rasterize("FV", "TEX")
    assert seen["args"] == ("FV", "TEX", 64, [0, 0, 0], 1, 100, True, 1e-3, 1e-5, "euclidean", 1e-10, 1e-4, "hard", "prod",
                            "surface", True)
    import inspect
    from umr_b200 import raster
    assert list(inspect.signature(raster.soft_rasterize).parameters)[2:] == list(rz.KERNEL_ARGS)
    # light / camera modules positionally, like the reference's own call sites (renderer.py:62-76)
    lt = sr.Lighting("surface", 0.8, [1, 1, 1], 0.5, [1, 1, 1], [0, 1, 0])
    assert lt.ambient.light_intensity == 0.8 and lt.directionals[0].light_intensity == 0.5
    tr = sr.Transform("look_at", None, None, 512, False, 30, 1.0, [0, 0, -2.732], [0, 0, 1])
    assert tr.transformer._eye == [0, 0, -2.732] and tr.transformer.perspective is False
    assert abs(sr.LookAt()._eye[2] + (1. / np.tan(np.radians(30)) + 1)) < 1e-12


def test_hypothesis_tiling_and_weighting_helpers():
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 5, 2, generator=g)
    ref = x.unsqueeze(1).repeat(1, 4, 1, 1).view(-1, 5, 2)          # loss_utils.py:260
    assert torch.equal(loss_utils.tile_hypotheses(x, 4), ref)
    f = torch.randint(0, 9, (3, 7, 3), generator=g)
    assert torch.equal(loss_utils.tile_hypotheses(f, 8), f.unsqueeze(1).repeat(1, 8, 1, 1).view(-1, 7, 3))
    per = torch.rand(12, generator=g)
    probs = torch.softmax(torch.rand(3, 4, generator=g), 1)
    ref = (per.view(3, -1) * probs).sum(dim=1).mean()                # loss_utils.py:271-273
    assert torch.equal(loss_utils.expected_over_hypotheses(per, probs), ref)


def test_corr_loss_chamfer_matches_reference_expression(monkeypatch):
    """CorrLossChamfer.forward (loss_utils.py:223-248) on CPU, with distChamfer swapped for the torch oracle."""
    monkeypatch.setattr(loss_utils, "distChamfer", oracle_losses.dist_chamfer)
    g = torch.Generator().manual_seed(4)
    B, V = 3, 60
    parts = [torch.arange(0, 10), torch.arange(10, 30), torch.arange(30, 35), torch.arange(35, 50)]
    m = loss_utils.CorrLossChamfer(None, 64, part_vertices=parts)
    verts = torch.rand(B, V, 3, generator=g) - 0.5
    cams = torch.from_numpy(synth.cameras(np.random.default_rng(0), B))
    pts = [torch.rand(B, n, 2, generator=g) - 0.5 for n in (10, 30, 10, 30)]
    loss, vert2d = m(pts[0], pts[1], pts[2], pts[3], verts, cams)
    # reference expression
    v2d = oracle_losses.orthographic_proj_withz(verts[:, torch.cat(parts)], cams)[:, :, :2]
    nums = [10, 30, 35, 50]
    h = oracle_losses.dist_chamfer(v2d[:, :nums[0]], pts[0])[0]
    b = oracle_losses.dist_chamfer(v2d[:, nums[0]:nums[1]], pts[1])[0]
    n = oracle_losses.dist_chamfer(v2d[:, nums[1]:nums[2]], pts[2])[0]
    k = oracle_losses.dist_chamfer(v2d[:, nums[2]:nums[3]], pts[3])[0]
    ref = torch.mean(torch.mean(torch.cat((h * 1, b * 1, n * 0, k * 0), dim=1), dim=1))
    assert torch.allclose(vert2d, v2d, atol=1e-6) and torch.allclose(loss, ref, atol=1e-6)
    per_sample = m(pts[0], pts[1], pts[2], pts[3], verts, cams, avg=False)
    assert per_sample.shape == (B,)


def test_regulariser_cpu_formulas_match_the_reference_restatement():
    """soft_renderer.LaplacianLoss / FlattenLoss (CPU branch, and the CSR table the CUDA kernels use) vs oracle/mesh_oracle.py."""
    import mesh_oracle as MO
    v, f = synth.icosphere(2)
    verts = torch.from_numpy(synth.bird_like(v, np.random.default_rng(0), 2))
    faces = torch.from_numpy(f.astype(np.int64))
    lap = sr.LaplacianLoss(torch.from_numpy(v), faces)
    assert torch.allclose(lap(verts), MO.laplacian_loss(verts, f), rtol=1e-5)
    assert torch.allclose(sr.FlattenLoss(faces)(verts), MO.flatten_loss(verts, f), rtol=1e-5)
    # CSR == off-diagonal part of the dense matrix; tcoef == transposed entries
    dense = MO.laplacian_matrix(v.shape[0], f)
    rp, col, coef, tcoef = lap.csr_rowptr.numpy(), lap.csr_col.numpy(), lap.csr_coef.numpy(), lap.csr_tcoef.numpy()
    rec = torch.eye(v.shape[0])
    for i in range(v.shape[0]):
        for e in range(rp[i], rp[i + 1]):
            rec[i, col[e]] = float(coef[e])
            assert float(tcoef[e]) == float(dense[col[e], i])
    assert torch.equal(rec, dense)
    ft = sr.FlattenLoss(faces).edge_table
    assert ft.shape[1] == 4 and ft.dtype == torch.int32 and ft.shape[0] == 3 * f.shape[0] // 2


def test_round2_extensions_reject_cpu_tensors_and_keep_the_cpu_paths():
    """The CUDA-only extensions (visibility kernels, 4-channel part maps, fused CorrLossChamfer) must fail loudly on CPU
    tensors -- there is no CPU fallback in the product -- while the modules' generic torch paths stay usable."""
    from umr_b200 import ops, raster
    fv = torch.zeros(1, 4, 9)
    with pytest.raises(TypeError):
        raster.visibility(fv, 16)
    with pytest.raises(TypeError):
        raster.visibility(fv, 16, want_faces=True)
    with pytest.raises(TypeError):
        ops.tex_cycle(torch.zeros(1, 4, 4, 2), torch.zeros(1, 4, 2), None, torch.zeros(1, 4, dtype=torch.uint8))
    with pytest.raises(TypeError):
        ops.corr_chamfer(torch.zeros(1, 8, 3), torch.zeros(1, 7), torch.zeros(4, dtype=torch.int32),
                         [torch.zeros(1, 2, 2)] * 4, (1, 2, 3, 4), (1, 1, 0, 0))
    r = smr.SoftRenderer(16, "hard")
    assert r.visible_faces(torch.zeros(1, 8, 3), torch.zeros(1, 4, 3, dtype=torch.long), torch.zeros(1, 7)) is None
    assert r.renderer.rasterizer.supports_visibility()
    assert not smr.SoftRenderer(16, "softmax").renderer.rasterizer.supports_visibility()
    # part_matching_loss keeps the reference-shaped buffers and adds ONE batch-shared 4-channel texture
    one_hot = torch.zeros(1, 6, 4, 5)
    one_hot[..., 2] = 1
    m = loss_utils.part_matching_loss(None, None, 0, im_size=16, batch_size=2, tex_size=2, stex_one_hot=one_hot)
    assert tuple(m.stex_parts.shape) == (1, 6, 4, 4) and tuple(m.stex1.shape) == (2, 6, 4, 3)
    assert float(m.stex_parts[..., 1].min()) == 1.0 and float(m.stex_parts[..., 0].max()) == 0.0
    assert "stex_parts" not in m.state_dict()          # not part of the reference's state


def test_launch_list_tool_finds_the_step_in_the_committed_log(capsys):
    """tools/launch_list.py on the committed ncu launch log of a C2 step."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import launch_list
    launch_list.main(os.path.join(root, "profiles", "r02_launches_C2_step.csv"))
    out = capsys.readouterr().out
    assert "one eager step = " in out and "k_raster_fwd3" in out and "k_raster_bwd2" in out
    share = [l for l in out.splitlines() if "k_raster_fwd3" in l and "%" in l]
    assert share and float(share[0].split("us")[1].split("%")[0]) > 40.0

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
    base = compat.overlay(str(ref), package="UMRT", workdir=str(tmp_path / "ov"))
    try:
        m = importlib.import_module("UMRT.nnutils.smr")
        assert m.SoftRenderer is smr.SoftRenderer            # ours
        assert importlib.import_module("UMRT.nnutils.scops_utils").MARK == "reference"   # theirs, via symlink
        assert importlib.import_module("UMRT.nnutils.chamfer_python").distChamfer is not None
    finally:
        sys.path.remove(base)


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

"""Seeded synthetic scenes for the benchmark configs of BASELINE.json / SURVEY.md §8(d).

The reference's real inputs cannot be had offline (CUB images, SCOPS part maps, the learned
template `mean_v.pth`, `meshzoo`), so every workload is generated here with fixed seeds:

* icosphere subdiv 3 (V=642, F=1280) / 4 (V=2562, F=5120) -- replaces `meshzoo.iso_sphere`
  used by the reference's `utils/mesh.py:37-41` (vertex order does not affect kernel parity);
* "CUB-bird" stand-in: anisotropic scale (1.0, 0.55, 0.45) + low-frequency displacement + noise;
* 7-dof cameras [s, tx, ty, qw, qx, qy, qz] built from the 8 hypothesis biases of
  `nnutils/cub_mesh.py:326-331`;
* smooth random images, texture flow U(-0.9, 0.9), masks, distance-transform barrier
  (`utils/image.py:130-141`), part points for the chamfer term.

Everything returns numpy / CPU torch; callers move tensors to the device.
"""
import math

import numpy as np


# ----------------------------------------------------------------------------------------------
# mesh
# ----------------------------------------------------------------------------------------------
def icosphere(subdiv=3):
    """Unit icosphere. subdiv=3 -> (642, 1280); subdiv=4 -> (2562, 5120). Faces are int32 CCW."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t),
         (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9),
         (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.asarray(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    faces = [tuple(x) for x in f]
    for _ in range(subdiv):
        cache = {}

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]

        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    return np.asarray(verts, dtype=np.float32), np.asarray(faces, dtype=np.int32)


def bird_like(verts, rng, batch, noise=0.02):
    """[V,3] unit sphere -> [batch,V,3] anisotropic, smoothly displaced, per-sample jittered."""
    v = verts.astype(np.float64)
    disp = 0.05 * (np.sin(3.0 * v[:, 0:1] + 0.5) * np.cos(2.0 * v[:, 1:2]) + np.sin(4.0 * v[:, 2:3]))
    base = (v * (1.0 + disp)) * np.array([1.0, 0.55, 0.45])
    out = base[None] + rng.normal(0.0, noise, size=(batch,) + base.shape)
    return out.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# cameras
# ----------------------------------------------------------------------------------------------
def hamilton(qa, qb):
    a0, a1, a2, a3 = qa
    b0, b1, b2, b3 = qb
    return np.array([a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3, a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
                     a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1, a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0])


def camera_biases(num=8):
    """The 8 camera-hypothesis quaternions of nnutils/cub_mesh.py:326-331."""
    rot = np.array([0.9239, 0.0, 0.3827, 0.0])
    q = [np.array([0.7071, 0.7071, 0.0, 0.0])]
    for i in range(1, num):
        q.append(hamilton(rot, q[i - 1]))
    return np.stack(q)


def cameras(rng, batch, hypo_index=None):
    """[batch,7] = [s, tx, ty, qw, qx, qy, qz]; s~U(.55,.85), t~U(-.1,.1), q = bias (x) small rot."""
    biases = camera_biases()
    cams = np.zeros((batch, 7), dtype=np.float64)
    for b in range(batch):
        k = (b % 8) if hypo_index is None else hypo_index[b]
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(-0.35, 0.35)
        small = np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * axis])
        q = hamilton(small, biases[k])
        q /= np.linalg.norm(q)
        cams[b, 0] = rng.uniform(0.55, 0.85)
        cams[b, 1:3] = rng.uniform(-0.1, 0.1, size=2)
        cams[b, 3:] = q
    return cams.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# images / flow / masks
# ----------------------------------------------------------------------------------------------
def smooth_images(rng, batch, size, channels=3):
    """Smooth random images in [0,1]: a few random low-frequency cosine modes per channel."""
    ys, xs = np.meshgrid(np.linspace(0, 1, size), np.linspace(0, 1, size), indexing="ij")
    out = np.zeros((batch, channels, size, size))
    for b in range(batch):
        for c in range(channels):
            acc = np.zeros((size, size))
            for _ in range(4):
                fx, fy = rng.uniform(0.5, 4.0, size=2)
                ph = rng.uniform(0, 2 * math.pi)
                acc += np.cos(2 * math.pi * (fx * xs + fy * ys) + ph)
            out[b, c] = 0.5 + 0.125 * acc
    return np.clip(out, 0, 1).astype(np.float32)


def texture_flow(rng, batch, faces, tex_size=6):
    return rng.uniform(-0.9, 0.9, size=(batch, faces, tex_size, tex_size, 2)).astype(np.float32)


def dt_barrier(mask, k=50):
    """utils/image.py:130-141 `compute_dt_barrier` (host scipy EDT)."""
    from scipy.ndimage import distance_transform_edt
    dist_out = distance_transform_edt(1 - mask)
    dist_in = distance_transform_edt(mask)
    diff = (dist_out - dist_in) / max(mask.shape)
    return (1.0 / (1 + np.exp(k * -diff))).astype(np.float32)


def ellipse_masks(rng, batch, size):
    """Cheap GT silhouettes (an off-centre ellipse per sample) in {0,1}."""
    ys, xs = np.meshgrid(np.linspace(-1, 1, size), np.linspace(-1, 1, size), indexing="ij")
    out = np.zeros((batch, size, size), dtype=np.float32)
    for b in range(batch):
        cx, cy = rng.uniform(-0.1, 0.1, size=2)
        ax, ay = rng.uniform(0.45, 0.7), rng.uniform(0.25, 0.45)
        th = rng.uniform(0, math.pi)
        xr = (xs - cx) * math.cos(th) + (ys - cy) * math.sin(th)
        yr = -(xs - cx) * math.sin(th) + (ys - cy) * math.cos(th)
        out[b] = ((xr / ax) ** 2 + (yr / ay) ** 2 <= 1.0).astype(np.float32)
    return out


def part_points(rng, batch, sizes=(10, 30, 10, 30)):
    return [rng.uniform(-0.5, 0.5, size=(batch, n, 2)).astype(np.float32) for n in sizes]


def part_vertex_sets(rng, num_verts, sizes=(40, 80, 40, 80)):
    perm = rng.permutation(num_verts)
    out, o = [], 0
    for n in sizes:
        out.append(np.sort(perm[o:o + n]).astype(np.int64))
        o += n
    return out


# ----------------------------------------------------------------------------------------------
# projected face vertices (the raster-space input of `soft_rasterize`), numpy restatement of the
# reference's host pipeline for tests that want raw [B,F,9] inputs.
# ----------------------------------------------------------------------------------------------
def quat_rotate_np(X, q):
    """nnutils/geom_utils.py:147-165 for numpy: X [B,N,3], q [B,4] -> rotated [B,N,3] (float64)."""
    X = X.astype(np.float64)
    q = q.astype(np.float64)[:, None, :]
    qw, qx, qy, qz = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    # t = X (x) conj(q)   (X as pure quaternion)
    t0 = x * qx + y * qy + z * qz
    t1 = x * qw - y * qz + z * qy
    t2 = x * qz + y * qw - z * qx
    t3 = -x * qy + y * qx + z * qw
    r1 = qw * t1 + qx * t0 + qy * t3 - qz * t2
    r2 = qw * t2 - qx * t3 + qy * t0 + qz * t1
    r3 = qw * t3 + qx * t2 - qy * t1 + qz * t0
    return np.stack([r1, r2, r3], axis=-1)


def raster_space_faces(verts, faces, cams):
    """verts [B,V,3], faces [F,3] or [B,F,3], cams [B,7] -> face_vertices [B,F,9] float32 in the
    raster frame: x = s*Xr_x+tx, y = -(s*Xr_y+ty), z = s*Xr_z + 5 + 2.732 (App. A-1)."""
    xr = quat_rotate_np(verts, cams[:, 3:7])
    s = cams[:, 0].astype(np.float64)[:, None]
    x = s * xr[..., 0] + cams[:, 1:2]
    y = -(s * xr[..., 1] + cams[:, 2:3])
    z = (s * xr[..., 2] + 5.0).astype(np.float32) + np.float32(2.732)
    pv = np.stack([x.astype(np.float32), y.astype(np.float32), z], axis=-1)
    if faces.ndim == 2:
        faces = np.broadcast_to(faces[None], (verts.shape[0],) + faces.shape)
    B = verts.shape[0]
    fv = np.stack([pv[b][faces[b]] for b in range(B)])  # [B,F,3,3]
    return np.ascontiguousarray(fv.reshape(B, faces.shape[1], 9).astype(np.float32))

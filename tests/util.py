"""Shared helpers for the parity tests: seeded scenes and tolerance reports."""
import numpy as np

from umr_b200 import synth


def scene(B=2, subdiv=3, tex_res=2, seed=0):
    """Seeded raster-space inputs: face_vertices [B,F,9] f32, textures [B,F,R*R,3] f32."""
    rng = np.random.default_rng(seed)
    v, f = synth.icosphere(subdiv)
    verts = synth.bird_like(v, rng, B)
    cams = synth.cameras(rng, B)
    fv = synth.raster_space_faces(verts, f, cams)
    tex = rng.uniform(0, 1, size=(B, f.shape[0], tex_res * tex_res, 3)).astype(np.float32)
    return fv, tex


def rel_report(name, got, ref, rtol=1e-4, atol=1e-6):
    """Returns (ok, message). ok <=> |got-ref| <= atol + rtol*max(|got|,|ref|) everywhere."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    diff = np.abs(got - ref)
    bound = atol + rtol * np.maximum(np.abs(got), np.abs(ref))
    bad = diff > bound
    nb = int(bad.sum())
    denom = np.linalg.norm(ref.ravel()) + 1e-30
    msg = "%s: max|d|=%.3e rel-L2=%.3e bad=%d/%d (%.4f%%) max|ref|=%.3e" % (
        name, diff.max() if diff.size else 0.0, np.linalg.norm(diff.ravel()) / denom, nb, diff.size,
        100.0 * nb / max(diff.size, 1), np.abs(ref).max() if ref.size else 0.0)
    if nb:
        i = np.unravel_index(np.argmax(diff - bound), diff.shape)
        msg += " worst@%s got=%.9g ref=%.9g" % (str(tuple(int(x) for x in i)), got[i], ref[i])
    return nb == 0, msg

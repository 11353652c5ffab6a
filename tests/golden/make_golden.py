"""Generate the committed golden fixtures from THE REFERENCE ITSELF run here (oracle A = the
reference's own rasteriser device code compiled for the host, oracle/ref_host_shim.cpp).

    python tests/golden/make_golden.py      # needs /root/reference (build container only)

Each .npz holds the seeded inputs and the reference's outputs, small enough to commit.  grad_textures
is stored only for T2 == 1, where the reference's undefined behaviour (kernel.cu:199-218) cannot
matter; for T2 > 1 it is pinned by finite differences in tests/test_oracle.py instead.
Deterministic: single-threaded runs (float atomics order).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (B, subdiv, tex_res, image_size, aa, rgb, seed)
    "softmax_aa_t1": (1, 2, 1, 32, True, "softmax", 11),
    "softmax_aa_t4": (2, 2, 2, 24, True, "softmax", 12),
    "hard_aa_t4": (1, 2, 2, 32, True, "hard", 13),
    "softmax_noaa_t9": (1, 1, 3, 40, False, "softmax", 14),
    "hard_noaa_t1": (1, 2, 1, 48, False, "hard", 15),
}


def main():
    import softras
    from util import scene
    assert softras.have_oracle_a(), "oracle A (reference on host) is required to (re)generate goldens"
    for name, (B, sd, tr, isz, aa, rgb, seed) in CASES.items():
        fv, tex = scene(B, sd, tr, seed)
        img, fwd, cfg = softras.render(fv, tex, isz, anti_aliasing=aa, impl="A", nthreads=1, aggr_func_rgb=rgb,
                                       sigma_val=1e-5, dist_eps=1e-10, gamma_val=1e-4)
        g = np.random.default_rng(seed + 100).normal(size=img.shape).astype(np.float32)
        gf, gt = softras.render_backward(fwd, cfg, g, anti_aliasing=aa, impl="A", nthreads=1)
        out = dict(face_vertices=fv, textures=tex, grad_images=g, images=img, aggrs_info=fwd["aggrs_info"],
                   p2f_info=fwd["p2f_info"], grad_faces=gf, image_size=isz, anti_aliasing=aa, rgb=rgb)
        if tr == 1:
            out["grad_textures"] = gt
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

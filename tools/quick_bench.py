"""Ad-hoc kernel timing (CUDA events) of the raster forward/backward at a given config."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from umr_b200 import raster, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--is", dest="isz", type=int, default=256)
    ap.add_argument("--subdiv", type=int, default=3)
    ap.add_argument("--R", type=int, default=6)
    ap.add_argument("--rgb", default="softmax")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    v, f = synth.icosphere(a.subdiv)
    verts = synth.bird_like(v, rng, a.B)
    cams = synth.cameras(rng, a.B)
    fv = torch.from_numpy(synth.raster_space_faces(verts, f, cams)).cuda().requires_grad_(True)
    tex = torch.rand(a.B, f.shape[0], a.R * a.R, 3, device="cuda").requires_grad_(True)
    g = torch.randn(a.B, 4, a.isz, a.isz, device="cuda")
    kw = dict(sigma_val=1e-5, dist_eps=1e-10, gamma_val=1e-4, aggr_func_rgb=a.rgb, anti_aliasing=True)
    for _ in range(3):
        img, _, _ = raster.soft_rasterize(fv, tex, a.isz, **kw)
        img.backward(g)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(a.iters):
        e[0].record()
        img, _, _ = raster.soft_rasterize(fv, tex, a.isz, **kw)
        e[1].record()
        img.backward(g)
        e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    tf /= a.iters; tb /= a.iters
    img, _, _ = raster.soft_rasterize(fv, tex, a.isz, **kw)
    sv = img.grad_fn.saved_tensors
    if len(sv) == 5:
        st = sv[4][:8].view(torch.int32).cpu().tolist()
        S = 2 * a.isz
        print("pair buffer: %.1f MB, blocks wanted %d (%.2f candidates-slots/pixel), tiles unsaved %d" % (
            sv[4].numel() / 1e6, st[0], st[0] * 32.0 / (a.B * S * S), st[1]))
    print("B=%d is=%d F=%d T2=%d %s: fwd %.3f ms  bwd %.3f ms  -> %.0f img/s (fwd+bwd), alpha mean %.4f" % (
        a.B, a.isz, f.shape[0], a.R * a.R, a.rgb, tf, tb, a.B / ((tf + tb) * 1e-3), img[:, 3].mean().item()))
    if a.rgb != "softmax":
        return

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        sink = []
        raster.set_profile_sink(sink)
        for _ in range(a.iters):
            fn()
        torch.cuda.synchronize()
        raster.set_profile_sink(None)
        pr = raster.collect_profile(sink)
        return sum(pr["fwd"]) / a.iters, sum(pr["bwd"]) / a.iters

    # (1) texture-only backward: detached geometry (UMR's texture branch), kernel time from the library's own events
    fvd = fv.detach()
    kf_full = timed(lambda: raster.soft_rasterize(fv, tex, a.isz, **kw)[0].backward(g))
    kf_tex = timed(lambda: raster.soft_rasterize(fvd, tex, a.isz, **kw)[0].backward(g))
    texd = tex.detach()
    kf_geo = timed(lambda: raster.soft_rasterize(fv, texd, a.isz, **kw)[0].backward(g))
    print("kernel time  full backward: fwd %.3f bwd %.3f ms | texture-only backward (detached geometry): fwd %.3f bwd %.3f ms"
          " | geometry-only backward (constant textures): fwd %.3f bwd %.3f ms" % (kf_full + kf_tex + kf_geo))
    # (2) the four part maps of part_matching_loss: one 4-channel render vs the 2 packed / 4 separate 3-channel renders
    F = f.shape[0]
    parts = torch.zeros(1, F, a.R * a.R, 5, device="cuda")
    parts.scatter_(3, torch.randint(0, 5, (1, F, a.R * a.R, 1), device="cuda"), 1.0)
    t4 = parts[..., 1:5].contiguous()
    t3 = [parts[..., k:k + 1].repeat(1, 1, 1, 3) for k in (1, 2, 3, 4)]
    g5, g4 = torch.randn(a.B, 5, a.isz, a.isz, device="cuda"), torch.randn(a.B, 4, a.isz, a.isz, device="cuda")
    one = timed(lambda: raster.soft_rasterize(fv, t4, a.isz, **kw)[0].backward(g5))

    def four():
        for t in t3:
            raster.soft_rasterize(fv, t, a.isz, **kw)[0].backward(g4)
    sep = timed(four)
    print("part maps: ONE 4-channel render fwd %.3f + bwd %.3f ms  vs  4 x 3-channel renders (reference pattern) fwd %.3f + bwd %.3f ms"
          % (one + sep))


if __name__ == "__main__":
    main()

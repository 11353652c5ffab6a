"""Times a train_s2-shaped LOSS step (experiments/train_s2.py:201-316 of the reference: camera-hypothesis mask loss,
mesh regularisers, texture losses on detached geometry, part matching, part-chamfer correspondence) through the
drop-in modules -- the widened hot path of SURVEY.md §8(f), eager (torch autograd between our kernels).  Not the
BASELINE metric (bench.py is); it answers "what does one step of the reference's loss section cost through this
library" and how much of it is raster kernels.

    python tools/train_step_bench.py [--batch 16] [--steps 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from umr_b200 import _lib, raster, synth
from umr_b200 import soft_renderer as sr
from umr_b200.nnutils import geom_utils, loss_utils


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hypotheses", type=int, default=8)
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--subdiv", type=int, default=3)        # 1280 faces
    ap.add_argument("--tex", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, H, IS, T = a.batch, a.hypotheses, a.image_size, a.tex
    rng = np.random.default_rng(0)
    v, f = synth.icosphere(a.subdiv)
    V, F = v.shape[0], f.shape[0]
    fs = torch.from_numpy(f.astype(np.int64))[None].repeat(B, 1, 1).to(dev)
    imgs = torch.from_numpy(synth.smooth_images(rng, B, IS)).to(dev)
    masks = torch.from_numpy(synth.ellipse_masks(rng, B, IS)).to(dev)
    dts = torch.from_numpy(np.stack([synth.dt_barrier(m) for m in masks.cpu().numpy()]))[:, None].to(dev)
    part_segs = torch.from_numpy(rng.uniform(0, 1, size=(B, 5, IS, IS)).astype(np.float32)).to(dev)
    part = rng.integers(0, 5, size=(F, T * T))
    one_hot = torch.zeros(1, F, T * T, 5)
    one_hot.scatter_(3, torch.from_numpy(part)[None, :, :, None], 1.0)
    part_vertices = [torch.from_numpy(p) for p in synth.part_vertex_sets(rng, V, sizes=(20, 40, 20, 40))]
    head, belly, neck, back = [torch.from_numpy(p).to(dev) for p in synth.part_points(rng, B)]
    rep = lambda t: t.unsqueeze(1).repeat(1, H, 1, 1).view(-1, t.size(1), t.size(2))

    mask_fn = loss_utils.MultiMaskLoss(IS, "softmax", H).to(dev)
    tex_fn = loss_utils.MultiTextureLoss(B, H, IS, "softmax", "l1", "smr").to(dev)
    part_fn = loss_utils.part_matching_loss(None, None, 0, im_size=IS, batch_size=B, tex_size=T, stex_one_hot=one_hot).to(dev)
    corr_fn = loss_utils.CorrLossChamfer(None, IS, part_vertices=part_vertices)
    fcpu = torch.from_numpy(f.astype(np.int64))
    lap_fn = sr.LaplacianLoss(torch.from_numpy(v), fcpu).to(dev)
    flat_fn = sr.FlattenLoss(fcpu).to(dev)

    mean_shape = torch.from_numpy(v.astype(np.float32)).to(dev).requires_grad_(True)
    delta = (0.05 * torch.from_numpy(synth.bird_like(v, rng, B) - v[None])).to(dev).requires_grad_(True)
    cams = torch.from_numpy(np.stack([synth.cameras(rng, H) for _ in range(B)])).to(dev).requires_grad_(True)   # [B,H,7]
    logits = torch.zeros(B, H, device=dev, requires_grad=True)
    flow = torch.from_numpy(synth.texture_flow(rng, B, F, T)).to(dev).requires_grad_(True)
    leaves = [mean_shape, delta, cams, logits, flow]

    def step():
        for t in leaves:
            t.grad = None
        pred_vs = mean_shape[None] + delta
        probs = torch.softmax(logits, 1)
        proj_cam = cams[:, 0].detach()
        mask_loss, mask_all = mask_fn(pred_vs, fs, cams, probs, masks)                      # train_s2.py:226
        tri = lap_fn(pred_vs).mean()
        flat = flat_fn(pred_vs).mean()
        tex = geom_utils.sample_textures(flow, imgs).contiguous().view(B, F, T * T, 3)      # :238-242
        tl, tdt, tcyc, _ = tex_fn(pred_vs.detach(), fs, cams.detach(), probs.detach(), proj_cam, imgs, masks, mask_all,
                                  tex, flow, dts)                                           # :248
        pl, _ = part_fn(pred_vs, fs, proj_cam, part_segs)                                  # :297
        ms_rep = mean_shape[None].expand(B, -1, -1).unsqueeze(1).repeat(1, H, 1, 1).view(-1, V, 3)
        corr = corr_fn(rep(head), rep(belly), rep(back), rep(neck), ms_rep, cams.view(-1, 7), avg=False)   # :306-312
        corr = (corr.view(B, H) * probs).sum(1).mean()
        total = mask_loss.mean() + 0.1 * tri + 0.005 * flat + 3.0 * tl.mean() + 3.0 * tdt.mean() + tcyc.mean() \
            + 0.1 * pl.mean() + corr
        total.backward()
        return total

    lib = _lib.load()
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    sink = []
    raster.set_profile_sink(sink)
    l0 = lib.umr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        total = step()
    e1.record()
    torch.cuda.synchronize()
    raster.set_profile_sink(None)
    ms = e0.elapsed_time(e1) / a.steps
    launches = (lib.umr_launch_count() - l0) / a.steps
    prof = raster.collect_profile(sink)
    kf, kb = sum(prof["fwd"]) / a.steps, sum(prof["bwd"]) / a.steps
    print("train_s2-shaped loss step: B=%d x %d hypotheses, is=%d, F=%d, T=%d  (loss %.5f)" % (B, H, IS, F, T, float(total)))
    print("  %.3f ms/step eager  -> %.0f samples/s, %.0f renders/s" % (ms, B / ms * 1e3, (B * H * 2 + B * 2) / ms * 1e3))
    print("  raster kernels: %d forward launches %.3f ms, %d backward launches %.3f ms  (%.0f %% of the step)"
          % (len(prof["fwd"]) // a.steps, kf, len(prof["bwd"]) // a.steps, kb, 100 * (kf + kb) / ms))
    print("  library kernel launches per step: %.0f" % launches)


if __name__ == "__main__":
    main()

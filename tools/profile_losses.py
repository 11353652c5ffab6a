"""Runs every loss / mesh-op kernel once at a bench config's sizes, for `ncu --set full` captures (profiles/):
    ncu --set full --clock-control none -k regex:"k_(sample|losshead|iou|masked|chamfer|visible|texcycle|laplacian|flatten|edt)" \
        -o gpurun_out/r02_losses python tools/profile_losses.py [C2|C3]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from umr_b200 import ops, synth
from umr_b200 import soft_renderer as sr
from umr_b200.nnutils import chamfer_python, geom_utils, loss_utils


def main(name="C2"):
    cfg = dict(bench.CONFIGS[name], name=name)
    dev = torch.device("cuda:0")
    B, IS, R = cfg["batch"], cfg["image_size"], cfg["tex_res"]
    v, f = synth.icosphere(cfg["subdiv"])
    F_ = f.shape[0]
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    imgs = rnd(B, 3, IS, IS)
    rgba = rnd(B, 4, IS, IS).requires_grad_(True)
    masks = (rnd(B, IS, IS) > 0.5).float()
    flow = (rnd(B, F_, R, R, 2) * 1.8 - 0.9).requires_grad_(True)
    tex = geom_utils.sample_textures(flow, imgs)
    (tex.sum() + loss_utils.texture_dt_loss(flow, rnd(B, 1, IS, IS))).backward()
    loss_utils.mask_texture_loss(rgba, imgs, masks, 2.5, 3.0).backward()
    (loss_utils.neg_iou_loss(rgba[:, 3], masks) + loss_utils.texture_loss_masks(rgba[:, :3], imgs, masks, rgba[:, 3])).backward()
    ids = torch.full((B, 4 * IS * IS), -1.0, device=dev)
    ids[:, IS * IS: 3 * IS * IS] = torch.arange(2 * IS * IS, device=dev).float().div(64).floor() % F_
    loss_utils.TexCycle()(flow, rnd(B, F_, 2), ids)[0].backward()
    for cb, n, m in ((128, 40, 10), (128, 80, 30), (1, 20000, 642)):
        a = (rnd(cb, n, 2) - 0.5).requires_grad_(True)
        d = chamfer_python.distChamfer(a, rnd(cb, m, 2) - 0.5)
        (d[0].sum() + d[1].sum()).backward()
    verts = torch.from_numpy(synth.bird_like(v, np.random.default_rng(0), B)).to(dev).requires_grad_(True)
    faces = torch.from_numpy(f.astype(np.int64))
    (sr.LaplacianLoss(torch.from_numpy(v), faces).to(dev)(verts).sum() + sr.FlattenLoss(faces).to(dev)(verts).sum()).backward()
    ops.dt_barrier(masks)
    torch.cuda.synchronize()
    print("profile_losses ok", name)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C2")

"""SURVEY.md §8(e) acceptance on real GPUs: with the batch sharded over 2 ranks, the all-reduced gradient of the shared
parameters equals the single-GPU gradient of the concatenated batch (fp32 summation tolerance) -- through OUR one-shot
peer-memory all-reduce kernel (csrc/collective.cu), eagerly, repeatedly (epoch logic), and replayed from a CUDA graph.
Needs 2 GPUs (skipped otherwise; the driver's multi-GPU tier and `gpurun --gpus 2` run it)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, IS, T = 4, 32, 2


def _data():
    from umr_b200 import synth
    rng = np.random.default_rng(5)
    v, f = synth.icosphere(2)
    return dict(mean=synth.bird_like(v, rng, 1, noise=0.0)[0], tex=rng.uniform(0, 1, size=(f.shape[0], T * T, 3)).astype(np.float32),
                delta=rng.normal(0, 0.02, size=(B, v.shape[0], 3)).astype(np.float32), cams=synth.cameras(rng, B),
                imgs=synth.smooth_images(rng, B, IS), masks=synth.ellipse_masks(rng, B, IS), faces=f.astype(np.int64))


def _inputs(dev, d, lo, hi):
    """Device-resident inputs of one shard (made once: nothing inside a step may touch host memory, the step is also
    captured into a CUDA graph)."""
    from umr_b200.nnutils import smr
    r = smr.SoftRenderer(IS, "softmax")
    r.ambient_light_only()
    t = lambda k: torch.from_numpy(d[k][lo:hi]).to(dev)
    return dict(r=r, faces=torch.from_numpy(d["faces"]).to(dev)[None].repeat(hi - lo, 1, 1), delta=t("delta"), cams=t("cams"),
                imgs=t("imgs"), masks=t("masks"))


def _step(x, red, mean, tex, weight):
    from umr_b200.nnutils import loss_utils
    red.zero_grads()
    images, _, _ = x["r"](mean[None] + x["delta"], x["faces"], x["cams"], tex[None])
    loss = loss_utils.mask_texture_loss(images, x["imgs"], x["masks"], 2.5, 3.0) * weight
    loss.backward()
    return red.reduce()


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from umr_b200.dist import FlatGradAllReduce, shard_range
    from umr_b200.graph import GraphedStep
    d = _data()
    mean = torch.from_numpy(d["mean"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(d["tex"]).to(dev).requires_grad_(True)
    red = FlatGradAllReduce([mean, tex], average=False, backend=os.environ.get("UMR_ALLREDUCE", "auto"))
    lo, hi = shard_range(B, rank, world)
    # per-image mean losses: weight each shard's batch-mean by its share so the sum over ranks == the full-batch mean
    w = (hi - lo) / B
    outs = []
    x = _inputs(dev, d, lo, hi)
    for _ in range(3):                                   # repeated eager calls: the epoch / flag protocol
        outs.append(_step(x, red, mean, tex, w).clone())
    if red.backend == "p2p":                             # our kernel is a plain kernel: the collective INSIDE the graph
        g = GraphedStep(lambda: _step(x, red, mean, tex, w), warmup=2)
        for _ in range(3):
            outs.append(g().clone())
    torch.cuda.synchronize()
    if rank == 0:
        ret["backend"] = red.backend
        ret["err"] = getattr(red, "p2p_error", None)
        ret["outs"] = [o.cpu() for o in outs]
        ret["g_mean"] = mean.grad.cpu()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("backend", ["auto", "nccl"])
def test_sharded_render_gradient_equals_full_batch_gradient(backend, monkeypatch):
    monkeypatch.setenv("UMR_ALLREDUCE", backend)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29631 if backend == "auto" else 29633, ret), nprocs=world, join=True)
    print("all-reduce backend:", ret["backend"], ret["err"])
    if backend == "auto":
        assert ret["backend"] == "p2p", "symmetric-memory p2p all-reduce unavailable: %s" % ret["err"]
    # single-GPU gradient of the concatenated batch
    from umr_b200.dist import FlatGradAllReduce
    dev = torch.device("cuda", 0)
    d = _data()
    mean = torch.from_numpy(d["mean"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(d["tex"]).to(dev).requires_grad_(True)
    red = FlatGradAllReduce([mean, tex], average=False)
    full = _step(_inputs(dev, d, 0, B), red, mean, tex, 1.0).cpu()
    scale = float(full.abs().max())
    for i, o in enumerate(ret["outs"]):
        assert torch.allclose(o, full, rtol=1e-4, atol=2e-6 * scale), (i, float((o - full).abs().max()), scale)
    assert torch.allclose(ret["g_mean"], mean.grad.cpu(), rtol=1e-4, atol=2e-6 * scale)

"""world_size-2 gloo test (CPU) of the N>1 path: batches shard on dim 0 and ONE flat all-reduce of the
shared-parameter gradient reproduces the single-process gradient of the whole batch."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umr_b200.dist import FlatGradAllReduce, shard_range


def _loss(shape, tex, delta, target):
    # stand-in for render+loss: any per-sample differentiable function of (shared params, sample)
    v = shape[None] + delta
    return ((v.sin() * target[:, None, None]).sum(dim=(1, 2)) + (tex[None] * target[:, None, None, None]).pow(2).sum(dim=(1, 2, 3))).sum()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    shape = torch.randn(20, 3, generator=g).requires_grad_(True)
    tex = torch.randn(12, 4, 3, generator=g).requires_grad_(True)
    delta = torch.randn(6, 20, 3, generator=g)
    target = torch.randn(6, generator=g)
    lo, hi = shard_range(6, rank, world)
    _loss(shape, tex, delta[lo:hi], target[lo:hi]).backward()
    red = FlatGradAllReduce([shape, tex], average=False)
    assert red.backend == "gloo"
    flat = red()                      # round-1 interface: pack -> all-reduce -> unpack
    if rank == 0:
        ret["flat"] = flat.clone()
        ret["g_shape"] = shape.grad.clone()
    # round-2 interface: .grad are views of the flat buffer, autograd accumulates into it, no pack / unpack
    red2 = FlatGradAllReduce([shape, tex], average=True)
    for _ in range(2):                # second pass: zero_grads() really restarts the accumulation
        red2.zero_grads()
        assert shape.grad.data_ptr() == red2.flat.data_ptr()
        _loss(shape, tex, delta[lo:hi], target[lo:hi]).backward()
        out = red2.reduce()
    if rank == 0:
        ret["avg"] = out[:red2.n].clone()
        ret["g_tex_avg"] = tex.grad.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_equals_full_batch_gradient():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29613, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    shape = torch.randn(20, 3, generator=g).requires_grad_(True)
    tex = torch.randn(12, 4, 3, generator=g).requires_grad_(True)
    delta = torch.randn(6, 20, 3, generator=g)
    target = torch.randn(6, generator=g)
    _loss(shape, tex, delta, target).backward()
    full = torch.cat([shape.grad.reshape(-1), tex.grad.reshape(-1)])
    assert torch.allclose(ret["flat"], full, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ret["g_shape"], shape.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ret["avg"], full / world, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ret["g_tex_avg"], tex.grad / world, rtol=1e-5, atol=1e-6)


def test_single_process_is_identity():
    p = torch.ones(4, requires_grad=True)
    (p * torch.arange(4.)).sum().backward()
    FlatGradAllReduce([p])()
    assert torch.equal(p.grad, torch.arange(4.))
    red = FlatGradAllReduce([p])
    assert red.backend == "none"
    red.zero_grads()
    (p * torch.arange(4.)).sum().backward()
    red.reduce()
    assert torch.equal(p.grad, torch.arange(4.)) and p.grad.data_ptr() == red.flat.data_ptr()

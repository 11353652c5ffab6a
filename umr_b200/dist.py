"""Data-parallel plumbing for the render path.

The per-image render is embarrassingly parallel: batches shard on dim 0 across one process per GPU
and the ONLY exchange per step is one all-reduce(sum) of the gradient of the SHARED parameters
(mean shape [V,3] + texture [F,T2,3], ~0.56 MB fp32) packed in one flat buffer.  The reference's
equivalent is the implicit reduce of `torch.nn.DataParallel` (experiments/train_s2.py:101,133,149,164).
Backend: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) shard of n items for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradAllReduce:
    """Packs the .grad of a fixed list of shared parameters into one flat fp32 buffer, all-reduces
    it once, optionally averages, and scatters the result back into the .grad tensors."""

    def __init__(self, params, average=True, group=None):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(self.sizes), device=dev, dtype=torch.float32)
        self.average, self.group = average, group

    def pack(self):
        o = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                self.flat[o:o + n].zero_()
            else:
                self.flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        return self.flat

    def unpack(self):
        o = 0
        for p, n in zip(self.params, self.sizes):
            g = self.flat[o:o + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += n

    def reduce(self):
        """The one collective of the step (no-op in a single process)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, group=self.group)
            if self.average:
                self.flat.mul_(1.0 / dist.get_world_size(self.group))
        return self.flat

    def __call__(self):
        self.pack()
        self.reduce()
        self.unpack()
        return self.flat

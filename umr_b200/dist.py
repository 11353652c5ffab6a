"""Data-parallel plumbing for the render path.

The per-image render is embarrassingly parallel: batches shard on dim 0 across one process per GPU
and the ONLY exchange per step is one all-reduce of the gradient of the SHARED parameters (mean
shape [V,3] + texture [F,T2,3], ~0.56 MB fp32) held in one flat buffer.  The reference's equivalent
is the implicit reduce of `torch.nn.DataParallel` (experiments/train_s2.py:101,133,149,164).

`FlatGradAllReduce`:
* the parameters' `.grad` tensors are VIEWS of one flat buffer (`attach()`): autograd accumulates
  straight into it, so there is no pack / unpack copy (round 1 issued 4 of them per step);
* backend "p2p" (NCCL process group on GPUs with peer access): the flat buffer is a symmetric-memory
  allocation and the all-reduce is OUR one-shot kernel `umr_p2p_allreduce` (csrc/collective.cu) reading
  the peers' buffers over NVLink -- one plain kernel, captured in the step's CUDA graph with
  everything else.  After it, `.grad` are views of the reduced (local) output buffer;
* backend "nccl"/"gloo": one `dist.all_reduce` on the flat buffer (NCCL: op AVG when averaging, so no
  separate scaling kernel); used when symmetric memory is unavailable and in the CPU tests.
"""
import ctypes

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) shard of n items for `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class FlatGradAllReduce:
    """One flat fp32 gradient buffer for a fixed list of shared parameters + its all-reduce.

    Typical step:   red.zero_grads(); loss.backward(); red.reduce()     # p.grad now hold the reduced gradient
    `__call__()` keeps the round-1 behaviour for callers whose `.grad` were produced elsewhere: pack -> reduce -> unpack.
    """

    def __init__(self, params, average=True, group=None, backend="auto"):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.n = sum(self.sizes)
        self.n_pad = (self.n + 3) // 4 * 4
        self.average, self.group = average, group
        self.world, self.rank = _world(group)
        dev = self.params[0].device
        self.device = dev
        self.backend = self._pick_backend(backend, dev)
        self._p2p = None
        if self.backend == "p2p":
            try:
                self._init_p2p(dev)
            except Exception as e:  # symmetric memory not available on this system: NCCL does the exchange
                self._p2p = None
                self.backend = "nccl"
                self.p2p_error = repr(e)
        if self._p2p is None:
            self.flat = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
            self.out = self.flat  # in-place collective
        self._attached = False

    # ------------------------------------------------------------------------------------------------
    def _pick_backend(self, backend, dev):
        if self.world <= 1:
            return "none"
        pg_backend = dist.get_backend(self.group)
        if backend == "auto":
            return "p2p" if (dev.type == "cuda" and "nccl" in str(pg_backend)) else str(pg_backend)
        return backend

    def _init_p2p(self, dev):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        lib = _lib.load()
        flag_bytes = int(lib.umr_p2p_allreduce_flag_bytes())
        flag_off = self.n_pad * 4
        total_floats = self.n_pad + flag_bytes // 4
        group = self.group if self.group is not None else dist.group.WORLD
        buf = symm_mem.empty(total_floats, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(buf, group)
        buf.zero_()                      # gradient area and flag words start at 0 on every rank ...
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)   # ... before any rank's first kernel signals a peer
        self.flat = buf[:self.n_pad]
        self.out = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        state = torch.zeros(4, device=dev, dtype=torch.int32)
        self._p2p = dict(buf=buf, hdl=hdl, state=state, flag_off=flag_off, peers=int(hdl.buffer_ptrs_dev), lib=lib)

    # ------------------------------------------------------------------------------------------------
    def _views(self, flat):
        out, o = [], 0
        for p, n in zip(self.params, self.sizes):
            out.append(flat[o:o + n].view_as(p))
            o += n
        return out

    def attach(self):
        """Make every parameter's .grad a view of the flat (accumulation) buffer."""
        for p, v in zip(self.params, self._views(self.flat)):
            p.grad = v
        self._attached = True

    def zero_grads(self):
        """Start of a step: zero the flat buffer (one memset) and (re-)attach the .grad views to it."""
        self.flat.zero_()
        self.attach()

    def reduce(self):
        """The one collective of the step.  Afterwards p.grad hold the (averaged) all-reduced gradient."""
        if self.world > 1:
            if self._p2p is not None:
                s = self._p2p
                scale = 1.0 / self.world if self.average else 1.0
                stream = torch.cuda.current_stream(self.device).cuda_stream
                rc = s["lib"].umr_p2p_allreduce(ctypes.c_void_p(s["peers"]), ctypes.c_void_p(self.out.data_ptr()),
                                                self.n_pad, s["flag_off"], ctypes.c_void_p(s["state"].data_ptr()),
                                                self.rank, self.world, scale, ctypes.c_void_p(stream))
                from . import _lib
                _lib.check(rc, "umr_p2p_allreduce")
            else:
                nccl = "nccl" in str(dist.get_backend(self.group))
                if self.average and nccl:
                    dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
                else:
                    dist.all_reduce(self.flat, group=self.group)
                    if self.average:
                        self.flat.mul_(1.0 / self.world)
        if self._attached and self.out is not self.flat:
            for p, v in zip(self.params, self._views(self.out)):
                p.grad = v
        return self.out

    # ---- round-1 interface (gradients produced into ordinary .grad tensors) ------------------------------
    def pack(self):
        o = 0
        for p, n in zip(self.params, self.sizes):
            if p.grad is None:
                self.flat[o:o + n].zero_()
            elif p.grad.data_ptr() != self.flat[o:o + n].data_ptr():
                self.flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        return self.flat

    def unpack(self):
        for p, v in zip(self.params, self._views(self.out)):
            if p.grad is None:
                p.grad = v.clone()
            elif p.grad.data_ptr() != v.data_ptr():
                p.grad.copy_(v)

    def __call__(self):
        self.pack()
        self.reduce()
        if not self._attached:
            self.unpack()
        return self.out[:self.n]

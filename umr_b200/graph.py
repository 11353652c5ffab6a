"""CUDA-graph capture of a whole render step (B200-first: streams and graphs instead of a tracing
compiler).  A step issues ~40 short launches (vertex kernel, prep, raster, loss kernels, a few torch
glue ops, the NCCL all-reduce); replaying them as one graph removes the launch/Python overhead that
otherwise sits between the raster kernels.

    step = GraphedStep(fn)      # fn(): no arguments, reads/writes STATIC tensors, returns tensor(s)
    out = step()                # graph replay; `out` is the static output of the capture

All umr_b200 entry points are capture-safe: they only enqueue kernels / memsets on the current stream
(no synchronisation, no per-call attribute changes, allocations go through torch's graph pool).
"""
import torch


class GraphedStep:
    def __init__(self, fn, warmup=3, pool=None):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls made by other host threads (e.g. the NCCL watchdog of an initialised process
        # group) must not invalidate this capture
        with torch.cuda.graph(self.graph, pool=pool, capture_error_mode="thread_local"):
            self.out = fn()

    def pool(self):
        return self.graph.pool()

    def __call__(self):
        self.graph.replay()
        return self.out

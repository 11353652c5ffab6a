"""`utils/image.py` pieces on the per-step path (reference: utils/image.py:120-141).

`compute_dt_barrier` is called once per image per training step on the host with scipy in the reference
(experiments/train_s2.py:196); here the whole batch is transformed by one pair of kernels (csrc/mesh_ops.cu `k_edt_*`).
"""
import numpy as np
import torch

from .. import ops


def compute_dt_barrier(mask, k=50):
    """Barrier distance transform.  mask: [H,W] or [B,H,W], numpy array or torch tensor (non-zero = object).
    Returns the same container type: numpy input -> numpy float64-compatible array (values computed on the GPU),
    cuda tensor -> cuda float32 tensor."""
    if isinstance(mask, np.ndarray):
        out = ops.dt_barrier(torch.from_numpy(np.ascontiguousarray(mask, dtype=np.float32)).cuda(), k)
        return out.cpu().numpy().astype(np.float64)
    if not mask.is_cuda:
        return ops.dt_barrier(mask.cuda(), k).to(mask.device)
    return ops.dt_barrier(mask, k)

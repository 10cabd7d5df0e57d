"""Mirror of the hot-path members of manydepth/layers.py: compute_depth_errors (:539-557), compute_depth_errors_numpy (:559-577)."""
import numpy as np
import torch

from .. import ops
from . import to_device


def compute_depth_errors(gt, pred):
    """Computation of error metrics between predicted and ground truth depths (CUDA tensors, already masked)."""
    return ops.compute_depth_errors(gt, pred)


def compute_depth_errors_numpy(gt, pred):
    """numpy arrays in, 7 numpy scalars out: abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3."""
    _, m = ops.depth_error_sums(to_device(np.asarray(gt), torch.float32), to_device(np.asarray(pred), torch.float32))
    return tuple(m.cpu().numpy())

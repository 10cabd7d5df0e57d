"""Mirror of polarisation/pol_split_and_save.py (split_pol, :10-27)."""
import numpy as np
import torch

from .. import ops
from . import to_device


def split_pol(img):
    """
    :param img: 2x2 polarised image, H x W (x C); numpy array or CUDA tensor
    :return: im00, im10, im01, im11  (same container type as the input; new arrays, not views)
    """
    if isinstance(img, torch.Tensor):
        return ops.split_pol(img)
    img = np.asarray(img)
    if img.ndim < 2 or img.shape[0] % 2 or img.shape[1] % 2:
        raise ValueError("array split does not result in an equal division")
    return tuple(q.cpu().numpy() for q in ops.split_pol(to_device(img)))

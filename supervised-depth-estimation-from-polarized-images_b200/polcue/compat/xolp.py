"""Mirror of polarisation/xolp.py (Iun_and_xolp, :8-34)."""
import numpy as np
import torch

from .. import ops
from . import to_device


def Iun_and_xolp(images, angles):
    """
    :param images: 4 concatenate images with different polarisation filters, H x W x 4
    :param angles: angles of the polarisation filters (radians)
    :return: Iun (unpolarised image), rho (DOLP), phi (AOLP) -- numpy float64 H x W for numpy input,
             CUDA float32 tensors for CUDA-tensor input
    """
    as_numpy = not isinstance(images, torch.Tensor)
    if as_numpy:
        images = np.asarray(images)
        if images.ndim != 3 or images.shape[2] != 4:
            raise ValueError(f"cannot reshape array of size {images.size} into shape ({images.shape[0] * images.shape[1]},4)")
        dev = to_device(images, None if images.dtype == np.uint8 else torch.float32)
    else:
        dev = images
    iun, xolp = ops.xolp_from_stack(dev, angles, want_iun=True)
    if as_numpy:
        iun, xolp = iun.cpu().numpy().astype(np.float64), xolp.cpu().numpy().astype(np.float64)
    return iun[0], xolp[0, 0], xolp[0, 1]

"""Mirror of polarisation/xolp.py (Iun_and_xolp, :8-34)."""
import numpy as np
import torch
import torch.utils.data

from .. import ops
from . import to_device


def Iun_and_xolp(images, angles):
    """
    :param images: 4 concatenate images with different polarisation filters, H x W x 4
    :param angles: angles of the polarisation filters (radians)
    :return: Iun (unpolarised image), rho (DOLP), phi (AOLP) -- numpy float64 H x W for numpy input,
             CUDA float32 tensors for CUDA-tensor input
    """
    worker = torch.utils.data.get_worker_info()
    if worker is not None and not torch.cuda.is_initialized() and torch.multiprocessing.get_start_method(allow_none=True) in (None, "fork"):
        raise RuntimeError(
            "polcue Iun_and_xolp was called inside a forked DataLoader worker (indoor_dataset.py:440 runs in __getitem__): CUDA "
            "cannot be used there. Use num_workers=0 or multiprocessing_context='spawn', or keep the uint8 planes in the loader "
            "and call polcue.ops.fused_planes / loader_front_end on the collated batch in the main process (INTEGRATION.md).")
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

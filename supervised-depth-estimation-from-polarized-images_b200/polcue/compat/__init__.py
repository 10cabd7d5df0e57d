"""Drop-in mirrors of the reference's hot-path functions.

Each module keeps the name, argument order/meaning, return order and error behaviour of the reference
module it replaces, so a call site changes only its import line (INTEGRATION.md lists them):

  reference module                               mirror
  polarisation/pol_split_and_save.py             polcue.compat.pol_split_and_save
  polarisation/xolp.py                           polcue.compat.xolp
  polarisation/xolp_and_normals.py               polcue.compat.xolp_and_normals
  ppp_code/physical_normals_channels.py          polcue.compat.physical_normals_channels
  manydepth/normals_vec.py                       polcue.compat.normals_vec
  manydepth/networks/pre_encoders.py (get_normals)  polcue.compat.pre_encoders
  manydepth/layers.py (compute_depth_errors*)    polcue.compat.layers
  kornia.geometry.depth (depth_to_normals)       polcue.compat.depth

numpy-signature functions take numpy arrays and return numpy float64 like the reference (the data makes one
round trip through the GPU); torch-signature functions take and return CUDA tensors.
"""
import torch


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("polcue needs a CUDA device: there is no CPU fallback for the polarization hot path")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(array, dtype=None):
    """numpy array -> CUDA tensor (contiguous)."""
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(array))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(default_device(), non_blocking=False)

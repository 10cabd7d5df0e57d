"""Mirror of ppp_code/physical_normals_channels.py (numpy float64 in / out).

PolarisationImage_channel (:15-36), rho_diffuse_channel (:39-48), rho_spec_channel (:51-72), calc_normals_channel (:75-83).
"""
import numpy as np
import torch

from .. import ops
from . import to_device
from .xolp_and_normals import rho_diffuse as rho_diffuse_channel  # noqa: F401  (same table inversion)
from .xolp_and_normals import rho_spec as rho_spec_channel        # noqa: F401


def PolarisationImage_channel(images, angles, mask):
    """s0 = I0 + I90 direct Stokes; `angles` is ignored, as in the reference.  Returns (rho, phi, Iun)."""
    rho, phi, iun = ops.stokes_channel(to_device(np.asarray(images), torch.float32), to_device(np.asarray(mask, dtype=np.uint8)))
    return tuple(t.cpu().numpy().astype(np.float64) for t in (rho, phi, iun))


def calc_normals_channel(phi, theta, mask):
    out = ops.calc_normals_channel(to_device(np.asarray(phi), torch.float32), to_device(np.asarray(theta), torch.float32),
                                   to_device(np.asarray(mask, dtype=np.uint8)))
    return out.cpu().numpy().astype(np.float64)

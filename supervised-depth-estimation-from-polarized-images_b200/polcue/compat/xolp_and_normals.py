"""Mirror of polarisation/xolp_and_normals.py: numpy float64 in / out.

Iun_and_xolp (:13-39), rho_spec (:41-67), rho_diffuse (:69-83), calc_normals (:85-98), and `process_frame`, the
chain of its `main` (:107-121) as ONE fused launch.
"""
import numpy as np
import torch

from .. import ops
from . import to_device
from .xolp import Iun_and_xolp  # noqa: F401  (identical copy in the reference, :13-39)


def rho_spec(rho, n):
    rho = np.asarray(rho)
    t1, t2 = ops.rho_spec(to_device(rho, torch.float32), n)
    return t1.cpu().numpy().astype(np.float64), t2.cpu().numpy().astype(np.float64)


def rho_diffuse(rho, n):
    rho = np.asarray(rho)
    return ops.rho_diffuse(to_device(rho, torch.float32), n).cpu().numpy().astype(np.float64)


def calc_normals(phi, theta):
    """:return: normals, shape (H, W, 3)"""
    phi, theta = np.asarray(phi), np.asarray(theta)
    out = ops.calc_normals(to_device(phi, torch.float32)[None], to_device(theta, torch.float32)[None])
    return out[0].permute(1, 2, 0).cpu().numpy().astype(np.float64)


def process_frame(img, n=1.5):
    """H x W uint8 mosaic -> (Iun, rho, phi, N_diff, N_spec1, N_spec2) exactly as `main` computes them (:107-121)."""
    out = ops.fused_mosaic(to_device(np.asarray(img, dtype=np.uint8)), n, want_iun=True)
    torch.cuda.synchronize()
    nrm = out["normals"][0].permute(1, 2, 0).cpu().numpy().astype(np.float64)
    xolp = out["xolp"][0].cpu().numpy().astype(np.float64)
    return (out["iun"][0].cpu().numpy().astype(np.float64), xolp[0], xolp[1], nrm[..., 0:3], nrm[..., 3:6], nrm[..., 6:9])

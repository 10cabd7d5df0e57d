"""Mirror of manydepth/normals_vec.py -- rho_diffuse (:11-22), rho_spec (:25-50), calc_normals (:53-60) -- and of their one
caller on the training path, ShallowNormalsEncoder.get_normals (manydepth/networks/pre_encoders.py:99-113).

Tensors stay on the GPU (the reference hops to the CPU for scipy and back); results are float32.
"""
from .. import ops


def rho_diffuse(rho, n):
    return ops.rho_diffuse(rho, n)          # B x H x W zenith angle


def rho_spec(rho, n):
    return ops.rho_spec(rho, n)             # (theta1, theta2)


def calc_normals(phi, theta):
    return ops.calc_normals(phi, theta)     # B x 3 x H x W


def get_normals(x, n=1.5):
    """input: XOLP B x 2 x H x W (CUDA); output: 3 concatenated normals B x 9 x H x W (float32): one kernel instead of the
    three table look-ups and three calc_normals calls above."""
    return ops.get_normals(x, n)


class GetNormalsMixin:
    """`class ShallowNormalsEncoder(GetNormalsMixin, ShallowEncoder)` keeps `self.get_normals(x)` working."""

    get_normals = staticmethod(get_normals)

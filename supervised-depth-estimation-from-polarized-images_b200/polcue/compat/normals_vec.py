"""Mirror of manydepth/normals_vec.py: rho_diffuse (:11-22), rho_spec (:25-50), calc_normals (:53-60).

Tensors stay on the GPU (the reference hops to the CPU for scipy and back); results are float32.
"""
from .. import ops


def rho_diffuse(rho, n):
    return ops.rho_diffuse(rho, n)          # B x H x W zenith angle


def rho_spec(rho, n):
    return ops.rho_spec(rho, n)             # (theta1, theta2)


def calc_normals(phi, theta):
    return ops.calc_normals(phi, theta)     # B x 3 x H x W

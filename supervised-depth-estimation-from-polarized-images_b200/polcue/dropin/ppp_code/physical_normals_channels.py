"""ppp_code/physical_normals_channels.py of the reference, served by polcue."""
from polcue.compat.physical_normals_channels import (PolarisationImage_channel, calc_normals_channel,  # noqa: F401
                                                     rho_diffuse_channel, rho_spec_channel)

"""polarisation/xolp_and_normals.py of the reference, served by polcue."""
from polcue.compat.xolp_and_normals import Iun_and_xolp, calc_normals, process_frame, rho_diffuse, rho_spec  # noqa: F401

"""polarisation/xolp.py of the reference, served by polcue."""
from polcue.compat.xolp import Iun_and_xolp  # noqa: F401

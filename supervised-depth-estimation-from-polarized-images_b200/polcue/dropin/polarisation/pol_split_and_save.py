"""polarisation/pol_split_and_save.py of the reference, served by polcue (split_pol only; `main` was a file-writing script)."""
from polcue.compat.pol_split_and_save import split_pol  # noqa: F401

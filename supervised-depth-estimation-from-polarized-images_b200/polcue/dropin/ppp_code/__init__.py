"""Shadow of the reference's `ppp_code` package (physical_normals_channels only; dataset.py is orphaned in the reference)."""

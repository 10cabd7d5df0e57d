"""Shadow of the reference's `polarisation` package: same module and function names, CUDA implementation."""

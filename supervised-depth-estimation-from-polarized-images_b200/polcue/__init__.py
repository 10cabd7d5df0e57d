"""polcue: the per-pixel polarization hot path of Supervised-Depth-Estimation-from-Polarized-Images on B200.

Layout
  polcue.ops      tensor-level operators (torch CUDA tensors <-> libpolcue.so through ctypes)
  polcue.compat   drop-in mirrors of the reference's Python functions (same names / argument meaning)
  polcue.dist     one-process-per-GPU sharding helpers (frames are independent; one tiny all-reduce)
  polcue.synth    seeded synthetic inputs shaped like the reference's data
  polcue._lib     the ctypes binding of include/polcue.h

There is no CPU implementation in this package: every compute call requires the CUDA library.
"""
__version__ = "0.1.0"

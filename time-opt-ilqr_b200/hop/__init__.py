"""hop -- host side of the B200-native HOP horizon-selection path.

``hop._cabi`` binds libhop_b200.so (hand-written sm_100a kernels behind a C ABI, include/hop_b200.h);
``hop.api`` exposes the batched entry points; ``hop.cases`` registers the benchmark dynamics.  PyTorch
is used for device-memory ownership, streams and torch.distributed only.  There is no CPU fallback:
every compute call raises if the CUDA library or a CUDA device is missing.
"""
from . import cases  # noqa: F401

__all__ = ["cases"]

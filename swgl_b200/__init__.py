"""swgl_b200 -- B200-native draw-call path behind the swgl GL-style C ABI.

The product is ``libswgl_b200.so`` (include/swgl.h, include/swgl_b200.h): C host layer +
hand-written CUDA kernels for sm_100a.  This package only carries the ctypes binding used by
the tests and the benchmark, the synthetic scene generators, and the multi-GPU plumbing
(one process per GPU, torch.distributed over NCCL for the stripe gather).
"""
from . import gl, scenes  # noqa: F401
from ._lib import load, swglStats  # noqa: F401

"""arrow_b200 -- B200-native execution layer for arrow::compute's ExecBatch hot path.

Everything computes on the GPU through libarrow_b200.so (C-ABI in include/arrow_b200.h).
There is no CPU fallback: importing compute kernels without the built library, or calling
them without a visible CUDA device, raises.
"""
from . import _cabi  # noqa: F401
from .device import Context, DeviceArray, DeviceBuffer, PinnedBuffer  # noqa: F401
from . import compute  # noqa: F401

__version__ = "0.1.0"

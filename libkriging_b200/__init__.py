"""libkriging_b200 -- B200-native (sm_100a) engine for libKriging's objective-evaluation hot path.

csrc/ + liblkgpu.so : hand-written CUDA kernels behind the C ABI of include/lkgpu.h
_capi               : ctypes binding of that ABI
kriging.Kriging     : host-side mirror of the reference's Kriging fit / objective / predict surface
parallel            : multistart sharding over one process per GPU (torch.distributed)
"""
import os as _os

# More hardware work queues than CUDA's default 8 for the handles that evaluate concurrently on one device (see
# csrc/engine.cu, lk_set_max_connections); it has to be in the environment before the CUDA context exists, so it is
# set on import -- a user-provided value wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .kriging import Kriging  # noqa: E402,F401

__all__ = ["Kriging"]

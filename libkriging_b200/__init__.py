"""libkriging_b200 -- B200-native (sm_100a) engine for libKriging's objective-evaluation hot path.

csrc/ + liblkgpu.so : hand-written CUDA kernels behind the C ABI of include/lkgpu.h
_capi               : ctypes binding of that ABI
kriging.Kriging     : host-side mirror of the reference's Kriging fit / objective / predict surface
parallel            : multistart sharding over one process per GPU (torch.distributed)
"""
from .kriging import Kriging  # noqa: F401

__all__ = ["Kriging"]

"""C++ host above the C ABI (lkgpu::Kriging, libkriging_b200/host/lkgpu_kriging.hpp) and a thin Python launcher
for its command-line driver (same work-directory protocol as the reference driver used by the tests)."""
from .driver import available, build, run  # noqa: F401

"""pose_refine_b200 -- B200-native (sm_100a) implementation of pose_refine's hot path.

The product is libpose_refine_b200.so (hand-written CUDA behind the C ABI declared in
include/pose_refine_b200.h).  This package is the thin Python host mirror of the reference's
C++ interface for that path (names follow cuda_renderer/renderer.h and cuda_icp/icp.h); PyTorch
is used only for device memory, streams and torch.distributed plumbing.

There is no CPU fallback: importing `pose_refine_b200.api` loads the shared library and raises
if it is missing; calling into it without an sm_100 device raises as well.
"""
from .build import build, LIB  # noqa: F401

__all__ = ["build", "LIB"]

"""icde2019-gpu-join_b200 -- B200-native radix hash-join engine (host-side Python binding).

The product is libgpujoin.so (hand-written CUDA for sm_100a behind the C ABI in
include/gpujoin.h).  This package is the thin ctypes layer over that ABI plus the
torch.distributed plumbing of the multi-GPU path; torch is used only for device memory,
streams and process groups.  There is no CPU fallback: importing works anywhere, but every
compute call raises if the CUDA library or a GPU is missing.

The directory name contains hyphens, so import it through `__graft_entry__.load_package()`
(registered in sys.modules as `icde2019_gpu_join_b200`).
"""
from .engine import (GJError, JoinEngine, JoinResult, Timings, lib, lib_path, C_ABI_SYMBOLS,  # noqa: F401
                     kernel_launch_count, bijection, payload_of_key, pcp_ctrl_bytes)
from . import generator  # noqa: F401
from . import distributed  # noqa: F401

__all__ = ["GJError", "JoinEngine", "JoinResult", "Timings", "lib", "lib_path", "generator",
           "distributed", "C_ABI_SYMBOLS", "kernel_launch_count", "bijection", "payload_of_key", "pcp_ctrl_bytes"]

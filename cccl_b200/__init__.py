"""cccl_b200 -- B200-native (sm_100a) LSD radix sort behind the reference's own interfaces.

Only what the hot path needs lives here:
  csrc/            hand-written CUDA kernels + the C ABI (include/b200rs.h) -> libb200rs.so
  radix_sort.py    mirror of cuda.compute.{radix_sort,make_radix_sort,DoubleBuffer,SortOrder}
  segmented_sort.py mirror of cuda.compute.{segmented_sort,make_segmented_sort}
  multi_gpu.py     one-process-per-GPU distributed sort over torch.distributed (NCCL / NVLink)
The C++ drop-ins (cub::DeviceRadixSort, thrust::sort) are the headers under include/.
"""
from ._native import B200RSError, LIB_PATH, lib  # noqa: F401
from .radix_sort import DoubleBuffer, SortOrder, make_radix_sort, radix_sort  # noqa: F401
from .segmented_sort import make_segmented_sort, segmented_sort  # noqa: F401

__all__ = ["DoubleBuffer", "SortOrder", "make_radix_sort", "radix_sort", "make_segmented_sort", "segmented_sort",
           "B200RSError", "LIB_PATH", "lib"]

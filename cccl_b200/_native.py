"""ctypes binding of the C ABI in include/b200rs.h (cccl_b200/libb200rs.so, built in-tree for sm_100a).

There is no fallback: if the CUDA library is missing or fails to load, importing the product path raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rs.so")

KEY_UINT, KEY_INT, KEY_FLOAT = 0, 1, 2

# every symbol include/b200rs.h declares (tests check the library exports exactly these)
EXPORTS = (
    "b200rs_version",
    "b200rs_sort",
    "b200rs_sort_inplace",
    "b200rs_sort_tuned",
    "b200rs_segmented_sort",
    "b200rs_sort_fields",
    "b200rs_topk",
    "b200rs_digit_histogram",
    "b200rs_splitter_ranks",
    "b200rs_select_histogram",
    "b200rs_bucket_ids",
    "b200rs_partition_by_splitters",
    "b200rs_partition_to_peers",
    "b200rs_multi_comm_create",
    "b200rs_multi_comm_destroy",
    "b200rs_multi_status",
    "b200rs_multi_last_launch_count",
    "b200rs_multi_timing_enable",
    "b200rs_multi_timing_read",
    "b200rs_sort_multi",
    "b200rs_last_launch_count",
    "b200rs_set_config",
    "b200rs_set_portion_items",
    "b200rs_set_force_big",
    "b200rs_set_single_tile",
    "b200rs_set_small_max",
    "b200rs_set_segmented_long_min",
    "b200rs_set_segmented_tiny_max",
    "b200rs_set_topk_small_max",
    "b200rs_describe_config",
    "b200rs_timing_enable",
    "b200rs_timing_read",
    # include/b200rs_cccl_c.h: the cccl.c.parallel names (c/parallel/include/cccl/c/radix_sort.h:63-143)
    "cccl_device_radix_sort_build",
    "cccl_device_radix_sort_build_ex",
    "cccl_device_radix_sort_compile",
    "cccl_device_radix_sort_load",
    "cccl_device_radix_sort",
    "cccl_device_radix_sort_link_ltoir",
    "cccl_device_radix_sort_serialize",
    "cccl_device_radix_sort_deserialize",
    "cccl_device_radix_sort_cleanup",
    "cccl_serialization_buffer_free",
)

# b200rs_allgather_fn: int (*)(void* ctx, const void* send, void* recv, size_t bytes_per_rank)
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

_lib = None


class B200RSError(RuntimeError):
    """A C-ABI call returned a non-zero cudaError_t."""

    def __init__(self, code: int, what: str):
        super().__init__(f"{what} failed with cudaError_t {code}")
        self.code = code


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C cccl_b200/csrc`). cccl_b200 has no CPU or PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        l.b200rs_version.restype = i32
        l.b200rs_version.argtypes = []
        l.b200rs_sort.restype = i32
        l.b200rs_sort.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, vp, vp, u64,
                                  i32, i32, i32, i32, i32, i32, i32, ctypes.POINTER(i32), vp]
        l.b200rs_sort_tuned.restype = i32
        l.b200rs_sort_tuned.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, vp, vp, u64,
                                        i32, i32, i32, i32, i32, i32, i32, ctypes.POINTER(i32), vp, vp]
        l.b200rs_sort_inplace.restype = i32
        l.b200rs_sort_inplace.argtypes = [vp, vp, u64, i32, i32, i32, i32, i32, vp]
        l.b200rs_segmented_sort.restype = i32
        l.b200rs_segmented_sort.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, vp, vp, u64, u64, vp, vp, i32,
                                            i32, i32, i32, i32, i32, i32, vp]
        l.b200rs_sort_fields.restype = i32
        l.b200rs_sort_fields.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, i32, vp, i32, vp, vp, i32, u64, i32,
                                         i32, i32, vp]
        l.b200rs_topk.restype = i32
        l.b200rs_topk.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, vp, vp, u64, u64, i32, i32, i32, i32, vp]
        l.b200rs_digit_histogram.restype = i32
        l.b200rs_digit_histogram.argtypes = [vp, u64, i32, i32, i32, i32, i32, vp, vp]
        l.b200rs_splitter_ranks.restype = i32
        l.b200rs_splitter_ranks.argtypes = [vp, u64, i32, i32, i32, vp, i32, vp, vp, vp]
        l.b200rs_select_histogram.restype = i32
        l.b200rs_select_histogram.argtypes = [vp, u64, i32, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp, u64, vp]
        l.b200rs_bucket_ids.restype = i32
        l.b200rs_bucket_ids.argtypes = [vp, u64, i32, i32, i32, ctypes.POINTER(u64), i32, vp, vp]
        l.b200rs_partition_by_splitters.restype = i32
        l.b200rs_partition_by_splitters.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, vp, vp, u64, i32, i32,
                                                    i32, i32, ctypes.POINTER(u64), i32, ctypes.POINTER(u64), vp]
        pu64 = ctypes.POINTER(u64)
        l.b200rs_partition_to_peers.restype = i32
        l.b200rs_partition_to_peers.argtypes = [vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, u64, i32, i32, i32, i32,
                                                pu64, i32, pu64, i32, pu64, pu64, pu64, vp]
        l.b200rs_multi_comm_create.restype = i32
        l.b200rs_multi_comm_create.argtypes = [ctypes.POINTER(vp), i32, i32, ctypes.c_size_t, ALLGATHER_FN, vp]
        l.b200rs_multi_comm_destroy.restype = i32
        l.b200rs_multi_comm_destroy.argtypes = [vp]
        l.b200rs_multi_status.restype = i32
        l.b200rs_multi_status.argtypes = [vp, ctypes.POINTER(i32)]
        l.b200rs_multi_last_launch_count.restype = i32
        l.b200rs_multi_last_launch_count.argtypes = [vp]
        l.b200rs_multi_timing_enable.restype = i32
        l.b200rs_multi_timing_enable.argtypes = [vp, i32]
        l.b200rs_multi_timing_read.restype = i32
        l.b200rs_multi_timing_read.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
        l.b200rs_sort_multi.restype = i32
        l.b200rs_sort_multi.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_size_t), vp, vp, vp, vp, u64, i32, i32, i32, i32, vp]
        l.b200rs_last_launch_count.restype = i32
        l.b200rs_last_launch_count.argtypes = []
        l.b200rs_set_config.restype = i32
        l.b200rs_set_config.argtypes = [i32]
        l.b200rs_set_portion_items.restype = i32
        l.b200rs_set_portion_items.argtypes = [ctypes.c_ulonglong]
        l.b200rs_set_force_big.restype = i32
        l.b200rs_set_force_big.argtypes = [i32]
        l.b200rs_set_small_max.restype = i32
        l.b200rs_set_small_max.argtypes = [ctypes.c_ulonglong]
        l.b200rs_set_segmented_long_min.restype = i32
        l.b200rs_set_segmented_long_min.argtypes = [ctypes.c_ulonglong]
        l.b200rs_set_segmented_tiny_max.restype = i32
        l.b200rs_set_segmented_tiny_max.argtypes = [ctypes.c_ulonglong]
        l.b200rs_set_topk_small_max.restype = i32
        l.b200rs_set_topk_small_max.argtypes = [ctypes.c_ulonglong]
        l.b200rs_set_single_tile.restype = i32
        l.b200rs_set_single_tile.argtypes = [i32]
        l.b200rs_describe_config.restype = i32
        l.b200rs_describe_config.argtypes = [i32, i32, i32, ctypes.c_char_p, ctypes.c_size_t]
        l.b200rs_timing_enable.restype = i32
        l.b200rs_timing_enable.argtypes = [i32]
        l.b200rs_timing_read.restype = i32
        l.b200rs_timing_read.argtypes = [ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_float), i32]
        _lib = l
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        raise B200RSError(code, what)


def sort_raw(d_temp: int, temp_bytes: int, keys_in: int, keys_out: int, vals_in: int, vals_out: int, num_items: int,
             key_kind: int, key_bytes: int, value_bytes: int, begin_bit: int, end_bit: int, descending: bool,
             is_overwrite_okay: bool, stream: int = 0):
    """Thin call of b200rs_sort.  Returns (temp_storage_bytes, selector); selector is -1 for the size query."""
    nbytes = ctypes.c_size_t(temp_bytes)
    selector = ctypes.c_int(-1)
    rc = lib().b200rs_sort(d_temp or None, ctypes.byref(nbytes), keys_in or None, keys_out or None, vals_in or None,
                           vals_out or None, num_items, key_kind, key_bytes, value_bytes, begin_bit, end_bit,
                           int(bool(descending)), int(bool(is_overwrite_okay)), ctypes.byref(selector), stream or None)
    check(rc, "b200rs_sort")
    return nbytes.value, selector.value


OP_NAMES = ("memset", "histogram", "scan", "onesweep", "copy", "single_tile", "small_sort")


def timing_read():
    """[(op name, device ms)] of the last b200rs_sort call made by this thread while timing was enabled."""
    kinds = (ctypes.c_int * 96)()
    ms = (ctypes.c_float * 96)()
    n = lib().b200rs_timing_read(kinds, ms, 96)
    if n < 0:
        raise B200RSError(-n, "b200rs_timing_read")
    return [(OP_NAMES[kinds[i]], float(ms[i])) for i in range(n)]


def describe_configs(key_bytes: int, value_bytes: int):
    n = lib().b200rs_describe_config(key_bytes, value_bytes, -1, None, 0)
    out = []
    for i in range(n):
        buf = ctypes.create_string_buffer(256)
        lib().b200rs_describe_config(key_bytes, value_bytes, i, buf, 256)
        out.append(buf.value.decode())
    return out

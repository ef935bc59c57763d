"""Host-side mirror of the reference's Python operator interface for this path:
``cuda.compute.radix_sort`` / ``make_radix_sort`` / ``DoubleBuffer`` / ``SortOrder``
(/root/reference/python/cuda_cccl/cuda/compute/algorithms/_sort/_radix_sort.py:19-300,
 .../_sort/_sort_common.py:15-30) -- same names, keyword arguments, two-phase temp-storage protocol and
selector behaviour -- bound to the ahead-of-time compiled sm_100a kernels through the C ABI
(include/b200rs.h) instead of an NVRTC build.

Arrays are anything exposing ``__cuda_array_interface__`` (torch CUDA tensors, cupy, numba) or a torch
tensor.  PyTorch is used only for device memory (temp storage) and streams.
"""
from __future__ import annotations

from enum import Enum

import numpy as np

from . import _native


class SortOrder(Enum):
    ASCENDING = 0
    DESCENDING = 1


class DoubleBuffer:
    """Pair of device arrays plus a selector, as cub::DoubleBuffer (cub/cub/util_type.cuh:749-779)."""

    def __init__(self, d_current, d_alternate):
        self.d_buffers = [d_current, d_alternate]
        self.selector = 0

    def current(self):
        return self.d_buffers[self.selector]

    def alternate(self):
        return self.d_buffers[1 - self.selector]


_TORCH_TO_NP = None


def _torch_np_dtype(t):
    global _TORCH_TO_NP
    import torch

    if _TORCH_TO_NP is None:
        _TORCH_TO_NP = {
            torch.uint8: np.uint8, torch.int8: np.int8, torch.int16: np.int16, torch.int32: np.int32,
            torch.int64: np.int64, torch.float32: np.float32, torch.float64: np.float64, torch.bool: np.bool_,
            # 16-bit floats: only (floating point, 2 bytes) matters to the kernels -- half and bfloat16 share the
            # sign-magnitude layout the key transform works on (reference: util_type.cuh:1017-1095)
            torch.float16: np.float16, torch.bfloat16: np.float16,
        }
        for name, npdt in (("uint16", np.uint16), ("uint32", np.uint32), ("uint64", np.uint64)):
            if hasattr(torch, name):
                _TORCH_TO_NP[getattr(torch, name)] = npdt
    return np.dtype(_TORCH_TO_NP[t.dtype])


def _describe(arr):
    """(device pointer, numpy dtype, number of elements) of a device array-like."""
    if arr is None:
        return 0, None, 0
    if hasattr(arr, "data_ptr") and hasattr(arr, "is_cuda"):  # torch tensor (no import needed to detect)
        if not arr.is_cuda:
            raise ValueError("cccl_b200.radix_sort needs CUDA device arrays; there is no CPU path")
        if not arr.is_contiguous():
            raise ValueError("radix sort input must be a contiguous array (pointer-like)")
        return arr.data_ptr(), _torch_np_dtype(arr), arr.numel()
    cai = getattr(arr, "__cuda_array_interface__", None)
    if cai is None:
        raise TypeError(f"not a device array: {type(arr)!r}")
    if cai.get("strides") is not None:
        raise ValueError("radix sort input must be a contiguous array (pointer-like)")
    n = 1
    for s in cai["shape"]:
        n *= s
    return cai["data"][0], np.dtype(cai["typestr"]), n


def _device_index(arr):
    """CUDA device ordinal an array lives on, or None when it cannot be told (plain __cuda_array_interface__ objects)."""
    if arr is None:
        return None
    if hasattr(arr, "is_cuda") and hasattr(arr, "device"):
        return arr.device.index
    dev = getattr(arr, "device", None)  # cupy: arr.device.id
    return getattr(dev, "id", None)


class _on_device_of:
    """The C library launches on the CURRENT device and stream (cudaGetDevice, as the reference does,
    dispatch_radix_sort.cuh:1764-1772): make the device of the arrays current for the call, take that device's current
    stream when none was given, and refuse arrays that live on different devices."""

    def __init__(self, *arrays):
        found = {d for d in (_device_index(a) for a in arrays) if d is not None}
        if len(found) > 1:
            raise ValueError(f"keys, values and temporary storage must be on one device, got devices {sorted(found)}")
        self.index = found.pop() if found else None
        self.guard = None

    def __enter__(self):
        if self.index is not None:
            import torch

            self.guard = torch.cuda.device(self.index)
            self.guard.__enter__()
        return self

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def key_kind_of(dtype: np.dtype) -> int:
    if dtype.kind == "f":
        return _native.KEY_FLOAT
    if dtype.kind == "i":
        return _native.KEY_INT
    if dtype.kind in ("u", "b"):
        return _native.KEY_UINT
    raise TypeError(f"unsupported key dtype {dtype}")


def _stream_handle(stream) -> int:
    if stream is None:
        import torch

        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    if hasattr(stream, "cuda_stream"):
        return stream.cuda_stream
    if hasattr(stream, "__cuda_stream__"):
        return stream.__cuda_stream__()[1]
    raise TypeError(f"not a stream: {type(stream)!r}")


def _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values):
    if isinstance(d_in_keys, DoubleBuffer):
        kin, kout = d_in_keys.current(), d_in_keys.alternate()
        if d_in_values is not None:
            assert isinstance(d_in_values, DoubleBuffer)
            vin, vout = d_in_values.current(), d_in_values.alternate()
        else:
            vin = vout = None
        return kin, kout, vin, vout
    return d_in_keys, d_out_keys, d_in_values, d_out_values


class _RadixSort:
    """Reusable sorter object; calling it with ``temp_storage=None`` returns the bytes required."""

    def __init__(self, d_in_keys, d_out_keys, d_in_values, d_out_values, order: SortOrder):
        kin, kout, vin, vout = _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values)
        _, kdt, _ = _describe(kin)
        self.key_dtype = kdt
        self.key_kind = key_kind_of(kdt)
        self.value_dtype = _describe(vin)[1] if vin is not None else None
        self.order = order
        _native.lib()  # fail here, loudly, if the CUDA library is absent

    def __call__(self, *, temp_storage, d_in_keys, d_out_keys, d_in_values, d_out_values, num_items: int,
                 begin_bit: int | None = None, end_bit: int | None = None, stream=None):
        kin, kout, vin, vout = _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values)
        pk_in, kdt, _ = _describe(kin)
        pk_out, kdt_out, _ = _describe(kout)
        if kdt != self.key_dtype or (kout is not None and kdt_out != kdt):
            raise TypeError("key dtype differs from the one this sorter was made for")
        pv_in, vdt, _ = _describe(vin)
        pv_out, _, _ = _describe(vout)
        vbytes = vdt.itemsize if vdt is not None else 0
        is_overwrite_okay = isinstance(d_in_keys, DoubleBuffer)
        if begin_bit is None:
            begin_bit = 0
        if end_bit is None:
            end_bit = kdt.itemsize * 8
        if temp_storage is None:
            d_temp, temp_bytes = 0, 0
        else:
            d_temp, _, _ = _describe(temp_storage)
            temp_bytes = temp_storage.numel() * temp_storage.element_size() if hasattr(temp_storage, "numel") \
                else temp_storage.nbytes
        with _on_device_of(kin, kout, vin, vout, temp_storage):
            temp_bytes, selector = _native.sort_raw(
                d_temp, temp_bytes, pk_in, pk_out, pv_in, pv_out, num_items, self.key_kind, kdt.itemsize, vbytes,
                begin_bit, end_bit, self.order is SortOrder.DESCENDING, is_overwrite_okay, _stream_handle(stream))
        if is_overwrite_okay and temp_storage is not None:
            assert selector in (0, 1)
            # the C ABI numbers buffers as (in=0, out=1) == (current, alternate) at call time
            new_sel = d_in_keys.selector ^ selector
            d_in_keys.selector = new_sel
            if d_in_values is not None:
                d_in_values.selector = new_sel
        return temp_bytes


def make_radix_sort(*, d_in_keys, d_out_keys, d_in_values, d_out_values, order: SortOrder, compute_capability=None):
    """Creates a reusable radix sort object (reference: _radix_sort.py:170-206).  `compute_capability` is accepted
    for signature compatibility; the kernels are built for sm_100a only."""
    return _RadixSort(d_in_keys, d_out_keys, d_in_values, d_out_values, order)


def radix_sort(*, d_in_keys, d_out_keys, d_in_values=None, d_out_values=None, num_items: int, order: SortOrder,
               begin_bit: int | None = None, end_bit: int | None = None, stream=None):
    """Device-wide radix sort with automatic temp-storage handling (reference: _radix_sort.py:209-300)."""
    import torch

    sorter = make_radix_sort(d_in_keys=d_in_keys, d_out_keys=d_out_keys, d_in_values=d_in_values,
                             d_out_values=d_out_values, order=order)
    kw = dict(d_in_keys=d_in_keys, d_out_keys=d_out_keys, d_in_values=d_in_values, d_out_values=d_out_values,
              num_items=num_items, begin_bit=begin_bit, end_bit=end_bit, stream=stream)
    nbytes = sorter(temp_storage=None, **kw)
    kin = _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values)[0]
    device = kin.device if hasattr(kin, "device") and hasattr(kin, "is_cuda") else torch.device("cuda")
    temp = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
    sorter(temp_storage=temp, **kw)
    return temp  # keep-alive handle: the sort is stream-ordered, not synchronised

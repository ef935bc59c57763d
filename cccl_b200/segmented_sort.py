"""Host-side mirror of the reference's Python segmented sort, ``cuda.compute.segmented_sort`` / ``make_segmented_sort``
(/root/reference/python/cuda_cccl/cuda/compute/algorithms/_sort/_segmented_sort.py:24-290) -- same names, keyword arguments,
two-phase temp-storage protocol and DoubleBuffer selector behaviour -- bound to ``b200rs_segmented_sort`` (include/b200rs.h:
a warp per tiny segment, a CTA per short one, whole-grid onesweep passes over all long ones) instead of an NVRTC build.
The sort is stable (the reference's segmented sort promises no stability; any of its tie orders is therefore matched).

Offsets are contiguous int32 / uint32 / int64 / uint64 device arrays; there is no CPU path."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native
from .radix_sort import DoubleBuffer, SortOrder, _describe, _get_arrays, _on_device_of, _stream_handle, key_kind_of


class _SegmentedSort:
    """Reusable sorter object; calling it with ``temp_storage=None`` returns the bytes required."""

    def __init__(self, d_in_keys, d_out_keys, d_in_values, d_out_values, start_offsets_in, end_offsets_in, order: SortOrder):
        kin, _, vin, _ = _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values)
        self.key_dtype = _describe(kin)[1]
        self.key_kind = key_kind_of(self.key_dtype)
        self.value_dtype = _describe(vin)[1] if vin is not None else None
        self.order = order
        _native.lib()  # fail here, loudly, if the CUDA library is absent

    def __call__(self, *, temp_storage, d_in_keys, d_out_keys, d_in_values, d_out_values, num_items: int, num_segments: int,
                 start_offsets_in, end_offsets_in, stream=None):
        if num_segments > np.iinfo(np.int32).max:
            raise RuntimeError("Segmented sort does not currently support more than 2^31-1 segments.")  # as the reference
        kin, kout, vin, vout = _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values)
        pk_in, kdt, _ = _describe(kin)
        pk_out, kdt_out, _ = _describe(kout)
        if kdt != self.key_dtype or (kout is not None and kdt_out != kdt):
            raise TypeError("key dtype differs from the one this sorter was made for")
        pv_in, vdt, _ = _describe(vin)
        pv_out, _, _ = _describe(vout)
        vbytes = vdt.itemsize if vdt is not None else 0
        pb, bdt, nb = _describe(start_offsets_in)
        pe, edt, ne = _describe(end_offsets_in)
        if bdt != edt or bdt.kind not in "iu" or bdt.itemsize not in (4, 8):
            raise TypeError("segment offsets must be two arrays of one 32- or 64-bit integer type")
        if nb < num_segments or ne < num_segments:
            raise ValueError("offset arrays are shorter than num_segments")
        if temp_storage is None:
            d_temp, temp_bytes = 0, 0
        else:
            d_temp, _, _ = _describe(temp_storage)
            temp_bytes = temp_storage.numel() * temp_storage.element_size() if hasattr(temp_storage, "numel") \
                else temp_storage.nbytes
        nbytes = ctypes.c_size_t(temp_bytes)
        with _on_device_of(kin, kout, vin, vout, temp_storage, start_offsets_in, end_offsets_in):
            rc = _native.lib().b200rs_segmented_sort(
                d_temp or None, ctypes.byref(nbytes), pk_in or None, pk_out or None, pv_in or None, pv_out or None, num_items,
                num_segments, pb or None, pe or None, bdt.itemsize, self.key_kind, kdt.itemsize, vbytes, 0, kdt.itemsize * 8,
                int(self.order is SortOrder.DESCENDING), _stream_handle(stream) or None)
        _native.check(rc, "b200rs_segmented_sort")
        if isinstance(d_in_keys, DoubleBuffer) and temp_storage is not None:
            # the C ABI sorts (current -> alternate): the result is in what was the alternate buffer
            d_in_keys.selector ^= 1
            if d_in_values is not None:
                d_in_values.selector = d_in_keys.selector
        return nbytes.value


def make_segmented_sort(*, d_in_keys, d_out_keys=None, d_in_values=None, d_out_values=None, start_offsets_in, end_offsets_in,
                        order: SortOrder, compute_capability=None):
    """Creates a reusable segmented sort object (reference: _segmented_sort.py:171-218).  `compute_capability` is accepted
    for signature compatibility; the kernels are built for sm_100a only."""
    return _SegmentedSort(d_in_keys, d_out_keys, d_in_values, d_out_values, start_offsets_in, end_offsets_in, order)


def segmented_sort(*, d_in_keys, d_out_keys=None, d_in_values=None, d_out_values=None, num_items: int, num_segments: int,
                   start_offsets_in, end_offsets_in, order: SortOrder, stream=None):
    """Device-wide segmented sort with automatic temp-storage handling (reference: _segmented_sort.py:221-290)."""
    import torch

    sorter = make_segmented_sort(d_in_keys=d_in_keys, d_out_keys=d_out_keys, d_in_values=d_in_values,
                                 d_out_values=d_out_values, start_offsets_in=start_offsets_in,
                                 end_offsets_in=end_offsets_in, order=order)
    kw = dict(d_in_keys=d_in_keys, d_out_keys=d_out_keys, d_in_values=d_in_values, d_out_values=d_out_values,
              num_items=num_items, num_segments=num_segments, start_offsets_in=start_offsets_in,
              end_offsets_in=end_offsets_in, stream=stream)
    nbytes = sorter(temp_storage=None, **kw)
    kin = _get_arrays(d_in_keys, d_out_keys, d_in_values, d_out_values)[0]
    device = kin.device if hasattr(kin, "device") and hasattr(kin, "is_cuda") else torch.device("cuda")
    temp = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
    sorter(temp_storage=temp, **kw)
    return temp  # keep-alive handle: the sort is stream-ordered, not synchronised

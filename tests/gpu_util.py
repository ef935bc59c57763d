"""Helpers for the -m gpu tests: move numpy arrays to cuda:0 as raw bytes and call the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from cccl_b200 import _native
from oracle_lib import key_kind_of


def to_dev(a: np.ndarray | None):
    if a is None:
        return None
    a = np.ascontiguousarray(a)
    if a.size == 0:
        return torch.empty(0, dtype=torch.uint8, device="cuda")
    return torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda()


def to_host(t: torch.Tensor, dtype, n) -> np.ndarray:
    return t.cpu().numpy().view(dtype)[:n].copy()


def gpu_sort(keys: np.ndarray, values: np.ndarray | None = None, *, descending=False, begin_bit=0, end_bit=None,
             api="pointer", stream=None, check_input_untouched=True, temp_misalign=0, kind=None):
    """Sort through b200rs_sort.  api='pointer' (is_overwrite_okay=0) or 'double' (DoubleBuffer semantics).
    Returns (keys_out[, values_out], info dict)."""
    n = keys.shape[0]
    kdt, kb = keys.dtype, keys.dtype.itemsize
    vb = values.dtype.itemsize if values is not None else 0
    if end_bit is None:
        end_bit = kb * 8
    kind = key_kind_of(kdt) if kind is None else kind  # override: e.g. bfloat16 bit patterns held in uint16
    d_k0 = to_dev(keys)
    d_k1 = torch.empty_like(d_k0)
    d_v0 = to_dev(values)
    d_v1 = torch.empty_like(d_v0) if values is not None else None
    ptr = lambda t: t.data_ptr() if t is not None and t.numel() > 0 else (0x100 if t is not None else 0)
    overwrite = api == "double"
    st = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
    need, _ = _native.sort_raw(0, 0, ptr(d_k0), ptr(d_k1), ptr(d_v0), ptr(d_v1), n, kind, kb, vb, begin_bit, end_bit,
                               descending, overwrite, st)
    temp = torch.empty(need + 512, dtype=torch.uint8, device="cuda")
    tptr = temp.data_ptr() + temp_misalign
    _, sel = _native.sort_raw(tptr, need, ptr(d_k0), ptr(d_k1), ptr(d_v0), ptr(d_v1), n, kind, kb, vb, begin_bit,
                              end_bit, descending, overwrite, st)
    launches = _native.lib().b200rs_last_launch_count()
    torch.cuda.synchronize()
    info = {"selector": sel, "temp_bytes": need, "launches": launches}
    if n == 0:
        out_k = np.empty(0, dtype=kdt)
        out_v = np.empty(0, dtype=values.dtype) if values is not None else None
    else:
        assert sel in (0, 1)
        if not overwrite:
            assert sel == 1
            if check_input_untouched:
                assert np.array_equal(to_host(d_k0, kdt, n).view(np.uint8), keys.view(np.uint8)), "input keys clobbered"
                if values is not None:
                    assert np.array_equal(to_host(d_v0, values.dtype, n).view(np.uint8), values.view(np.uint8))
        out_k = to_host(d_k1 if sel == 1 else d_k0, kdt, n)
        out_v = to_host(d_v1 if sel == 1 else d_v0, values.dtype, n) if values is not None else None
    return (out_k, out_v, info) if values is not None else (out_k, info)


def gpu_histogram(keys: np.ndarray, *, descending=False, begin_bit=0, end_bit=None, offset_items=0):
    kb = keys.dtype.itemsize
    if end_bit is None:
        end_bit = kb * 8
    passes = (end_bit - begin_bit + 7) // 8
    # offset_items shifts the device pointer to exercise unaligned heads
    pad = np.zeros(offset_items, dtype=keys.dtype)
    d_k = to_dev(np.concatenate([pad, keys]))
    bins = torch.zeros(max(passes, 1) * 256, dtype=torch.int64, device="cuda")
    rc = _native.lib().b200rs_digit_histogram(d_k.data_ptr() + offset_items * kb, keys.shape[0], key_kind_of(keys.dtype),
                                              kb, begin_bit, end_bit, int(descending), bins.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
    _native.check(rc, "b200rs_digit_histogram")
    torch.cuda.synchronize()
    return bins.cpu().numpy().view(np.uint64).reshape(-1, 256)[:passes]


def assert_same_bits(a: np.ndarray, b: np.ndarray, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
        bad = np.flatnonzero((a.view(np.uint8) != b.view(np.uint8)).reshape(a.shape[0], -1).any(axis=1))
        raise AssertionError(f"{what}: {bad.size} of {a.shape[0]} items differ, first at {bad[:5]}: "
                             f"got {a[bad[:5]]} want {b[bad[:5]]}")

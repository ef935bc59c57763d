"""ctypes loaders for the TEST-ONLY CPU oracle (oracle/liboracle.so) and the compiled reference
(oracle/_ref/libref_thrust_{omp,cpp}.so).  Nothing under cccl_b200/ imports this module."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

KIND_UINT, KIND_INT, KIND_FLOAT = 0, 1, 2


def key_kind_of(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return KIND_FLOAT
    if dtype.kind == "i":
        return KIND_INT
    if dtype.kind in ("u", "b"):
        return KIND_UINT
    raise TypeError(f"unsupported key dtype {dtype}")


def _build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            _build()
        lib = ctypes.CDLL(path)
        lib.oracle_radix_sort.restype = ctypes.c_int
        lib.oracle_radix_sort.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint64] + [ctypes.c_int] * 6
        lib.oracle_radix_sort_segment.restype = ctypes.c_int
        lib.oracle_radix_sort_segment.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint64] + [ctypes.c_int] * 6
        lib.oracle_digit_histogram.restype = ctypes.c_int
        lib.oracle_digit_histogram.argtypes = [ctypes.c_void_p, ctypes.c_uint64] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
        lib.oracle_counting_pass_u32.restype = ctypes.c_int
        lib.oracle_counting_pass_u32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
        _oracle = lib
    return _oracle


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def oracle_sort(keys: np.ndarray, values: np.ndarray | None = None, *, descending=False, begin_bit=0, end_bit=None,
                segment=False):
    """CPU restatement of cub/test/catch2_radix_sort_helper.cuh:270-312 (stable, bit-window aware).  segment=True: the
    order of ONE segment of cub::DeviceSegmentedRadixSort (oracle_radix_sort_segment: same, with the segmented kernel's
    -0.0 rule at every size)."""
    keys = np.ascontiguousarray(keys)
    kb = keys.dtype.itemsize
    if end_bit is None:
        end_bit = kb * 8
    kout = np.empty_like(keys)
    vout = None
    vb = 0
    if values is not None:
        values = np.ascontiguousarray(values)
        vb = values.dtype.itemsize
        assert values.shape[0] == keys.shape[0]
        vout = np.empty_like(values)
    fn = oracle().oracle_radix_sort_segment if segment else oracle().oracle_radix_sort
    rc = fn(_ptr(keys), _ptr(kout), _ptr(values), _ptr(vout), keys.shape[0], key_kind_of(keys.dtype), kb, vb,
            begin_bit, end_bit, int(descending))
    assert rc == 0
    return (kout, vout) if values is not None else kout


def oracle_histogram(keys: np.ndarray, *, descending=False, begin_bit=0, end_bit=None) -> np.ndarray:
    keys = np.ascontiguousarray(keys)
    kb = keys.dtype.itemsize
    if end_bit is None:
        end_bit = kb * 8
    passes = (end_bit - begin_bit + 7) // 8
    bins = np.zeros((max(passes, 1), 256), dtype=np.uint64)
    rc = oracle().oracle_digit_histogram(_ptr(keys), keys.shape[0], key_kind_of(keys.dtype), kb, begin_bit, end_bit,
                                         int(descending), _ptr(bins))
    assert rc == 0
    return bins[:passes]


_refs = {}


def ref_thrust(backend="omp"):
    """The reference's own thrust::sort compiled from /root/reference (None if never built)."""
    if backend not in _refs:
        path = os.path.join(ORACLE_DIR, "_ref", f"libref_thrust_{backend}.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/thrust"):
                _build()
            if not os.path.exists(path):
                _refs[backend] = None
                return None
        lib = ctypes.CDLL(path)
        lib.ref_thrust_sort.restype = ctypes.c_double
        lib.ref_thrust_sort.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64] + [ctypes.c_int] * 4
        lib.ref_thrust_max_threads.restype = ctypes.c_int
        lib.ref_thrust_set_threads.argtypes = [ctypes.c_int]
        _refs[backend] = lib
    return _refs[backend]


def ref_thrust_sort(keys: np.ndarray, values: np.ndarray | None = None, *, descending=False, backend="omp"):
    """Returns (sorted keys[, values], seconds) from the unmodified reference CPU path."""
    lib = ref_thrust(backend)
    assert lib is not None, "oracle/_ref not built"
    k = np.array(keys, copy=True, order="C")
    v = np.array(values, copy=True, order="C") if values is not None else None
    secs = lib.ref_thrust_sort(_ptr(k), _ptr(v), k.shape[0], key_kind_of(k.dtype), k.dtype.itemsize,
                               v.dtype.itemsize if v is not None else 0, int(descending))
    assert secs >= 0, "unsupported type combination in the reference wrapper"
    return (k, v, secs) if v is not None else (k, secs)


def oracle_segmented_sort(keys: np.ndarray, values, begin_offsets, end_offsets, *, descending=False, begin_bit=0,
                          end_bit=None):
    """CPU restatement of the segmented reference check (cub/test/catch2_radix_sort_helper.cuh:314-429,
    cub/cub/device/device_segmented_radix_sort.cuh): every segment [begin[s], end[s]) is sorted independently with the
    same stable, bit-window-aware order as oracle_sort; positions outside every segment keep the input.
    Groundwork for SURVEY.md 8f-1 (cub::DeviceSegmentedRadixSort); pinned by tests/golden/segmented/cubseg_*.npz."""
    kout = np.array(keys, copy=True)
    vout = np.array(values, copy=True) if values is not None else None
    for b, e in zip(np.asarray(begin_offsets).tolist(), np.asarray(end_offsets).tolist()):
        if e <= b:
            continue
        if values is None:
            kout[b:e] = oracle_sort(keys[b:e], descending=descending, begin_bit=begin_bit, end_bit=end_bit,
                                    segment=True)
        else:
            kout[b:e], vout[b:e] = oracle_sort(keys[b:e], values[b:e], descending=descending, begin_bit=begin_bit,
                                               end_bit=end_bit, segment=True)
    return (kout, vout) if values is not None else kout


def oracle_topk(keys: np.ndarray, values: np.ndarray | None, k: int, largest: bool):
    """CPU restatement of cub::DeviceTopK (cub/cub/device/device_topk.cuh:179-200): the K best items, here in sorted order
    (the reference returns them unordered and any subset of the ties of the K-th key; its own tests sort before comparing,
    cub/test/catch2_test_device_topk_api.cu:203-210).  Stable oracle sort in the requested direction, first min(k, n) items."""
    kk = min(int(k), keys.shape[0])
    if values is None:
        return oracle_sort(keys, descending=largest)[:kk]
    sk, sv = oracle_sort(keys, values, descending=largest)
    return sk[:kk], sv[:kk]

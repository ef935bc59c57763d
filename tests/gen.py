"""Seeded synthetic inputs for the radix-sort parity tests and the bench (numpy, host side).

Distributions follow the reference's benchmark/test generators:
  uniform   all bit patterns equally likely                     (nvbench_helper/nvbench_helper.cu:123-141)
  entropy   bitwise AND of k uniform words, k=1,2,3,4,5 <=> bit entropy 1.000/0.811/0.544/0.337/0.201
            (nvbench_helper/nvbench_helper.cu:370-400; cub/test/catch2_test_device_radix_sort_keys.cu:287-340)
  equal     every key the same                                  (catch2_test_device_radix_sort_keys.cu:342-374)
  few       k distinct values
  sorted / reverse
"""
from __future__ import annotations

import numpy as np


def raw_bits(rng, n, itemsize):
    udt = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[itemsize]
    return rng.integers(0, np.iinfo(udt).max, size=n, dtype=udt, endpoint=True)


def make_keys(dist: str, n: int, dtype, seed: int = 42) -> np.ndarray:
    dtype = np.dtype(dtype)
    rng = np.random.default_rng(seed)
    isz = dtype.itemsize
    if dist == "uniform":
        bits = raw_bits(rng, n, isz)
    elif dist.startswith("entropy"):
        k = int(dist[len("entropy"):] or 5)
        bits = raw_bits(rng, n, isz)
        for _ in range(k - 1):
            bits &= raw_bits(rng, n, isz)
    elif dist == "equal":
        bits = np.full(n, 4, dtype={1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[isz])
    elif dist.startswith("few"):
        k = int(dist[len("few"):] or 16)
        pool = raw_bits(rng, k, isz)
        bits = pool[rng.integers(0, k, size=n)]
    elif dist == "sorted":
        bits = np.sort(raw_bits(rng, n, isz))
    elif dist == "reverse":
        bits = np.sort(raw_bits(rng, n, isz))[::-1].copy()
    else:
        raise ValueError(dist)
    keys = bits.view(dtype) if dtype.kind != "b" else (bits & 1).astype(np.bool_)
    return np.ascontiguousarray(keys)


def make_values(n: int, dtype=np.uint32) -> np.ndarray:
    """iota values make stability observable."""
    dtype = np.dtype(dtype)
    if dtype.itemsize == 16:
        v = np.zeros(n, dtype=dtype)
        v.view(np.uint64).reshape(n, 2)[:, 0] = np.arange(n, dtype=np.uint64)
        v.view(np.uint64).reshape(n, 2)[:, 1] = ~np.arange(n, dtype=np.uint64)
        return v
    return (np.arange(n, dtype=np.uint64) & np.uint64(np.iinfo(dtype).max)).astype(dtype) if dtype.kind in "ui" \
        else np.arange(n).astype(dtype)


V16 = np.dtype([("a", np.uint64), ("b", np.uint64)])

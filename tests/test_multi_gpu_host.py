"""Host logic of the multi-GPU sort (cccl_b200/multi_gpu.py) on CPU: world_size-2 and -3 gloo process groups, with the
two device primitives replaced by the numpy oracle (checker infrastructure; the product default is CudaOps, which has
no CPU path).  Checks SURVEY.md 10.17: concatenated outputs == one stable sort of the concatenated inputs (keys and
values), per-rank output counts == input counts, with empty ranks, one item total, all-equal keys, uneven sizes."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from gen import make_keys  # noqa: E402
from oracle_lib import oracle_sort, key_kind_of as okind  # noqa: E402

TDT = {np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
       np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int16): torch.int16}
for _n in ("uint16", "uint32", "uint64"):
    if hasattr(torch, _n):
        TDT[np.dtype(_n)] = getattr(torch, _n)


def to_np(t):
    from cccl_b200.radix_sort import _torch_np_dtype

    dt = _torch_np_dtype(t)
    return t.view(torch.uint8).numpy().view(dt).copy() if t.numel() else np.empty(0, dtype=dt)


def to_torch(a):
    return torch.from_numpy(a.view(np.uint8).copy()).view(TDT[a.dtype]) if a.size else torch.empty(0, dtype=TDT[a.dtype])


def _view(bits, kind, kb, descending):
    """bit-ordered digit view, as the kernels compute it (common.cuh twiddle_in + digit_view)"""
    b = bits.astype(np.uint64)
    allm = np.uint64((1 << (8 * kb)) - 1)
    high = np.uint64(1 << (8 * kb - 1))
    if kind == 2:
        m = np.where((b & high) != 0, allm, high)
    elif kind == 1:
        m = np.full_like(b, high)
    else:
        m = np.zeros_like(b)
    t = (b ^ m) & allm
    if descending:
        t = t ^ allm
    if kind == 2:
        t = np.where(t == (allm ^ high), high, t)
    return t


class OracleOps:
    """numpy stand-ins for the device primitives (tests only)."""

    def _kv(self, keys, descending):
        k = to_np(keys)
        kb = k.dtype.itemsize
        return _view(k.view(np.dtype(f"u{kb}")), okind(k.dtype), kb, descending), kb

    def top_digit_histogram(self, keys, descending):
        return self.select_histogram(keys, torch.zeros(1, dtype=torch.int64), 0, descending)

    def select_histogram(self, keys, prefixes, rnd, descending, candidates="none"):
        prefixes = prefixes.cpu().numpy().astype(np.int64).view(np.uint64)
        v, kb = self._kv(keys, descending)
        lo = np.uint64(8 * kb - 8 * (rnd + 1))
        hi = (v >> (lo + np.uint64(8))) if int(lo) + 8 < 64 else np.zeros_like(v)
        if int(lo) + 8 >= 8 * kb:
            hi = np.zeros_like(v)
        dig = ((v >> lo) & np.uint64(255)).astype(np.int64)
        out = np.zeros((len(prefixes), 256), dtype=np.int64)
        for p, pref in enumerate(prefixes):
            out[p] = np.bincount(dig[hi == np.uint64(pref)], minlength=256)
        return torch.from_numpy(out)

    def bucket_ids(self, keys, splitters, descending):
        v, _ = self._kv(keys, descending)
        ids = np.zeros(v.shape[0], dtype=np.int64)
        for s in splitters:
            ids += (v > np.uint64(s)).astype(np.int64) + (v >= np.uint64(s)).astype(np.int64)
        return torch.from_numpy(ids.astype(np.uint8))

    def partition(self, ids, nbits, keys, values):
        order = np.argsort(ids.numpy() & ((1 << max(nbits, 1)) - 1), kind="stable")
        pk = to_torch(to_np(keys)[order])
        pv = to_torch(to_np(values)[order]) if values is not None else None
        return pk, pv

    def sort_pairs(self, keys, values, descending, preserve_input=False, out=None):
        k = to_np(keys)
        if values is None:
            return to_torch(oracle_sort(k, descending=descending)), None
        ok, ov = oracle_sort(k, to_np(values), descending=descending)
        return to_torch(ok), to_torch(ov)

    def splitter_ranks(self, sorted_keys, probes_bits, descending):
        k = to_np(sorted_keys)
        kind, kb = okind(k.dtype), k.dtype.itemsize
        ut = np.dtype(f"u{kb}")

        def view(bits):
            return _view(bits, kind, kb, descending)

        kv = view(k.view(ut))
        pv = view(probes_bits)
        assert (np.diff(kv.astype(np.float64)) >= 0).all() or kv.size < 2 or (kv[1:] >= kv[:-1]).all()
        lt = np.searchsorted(kv, pv, side="left").astype(np.int64)
        le = np.searchsorted(kv, pv, side="right").astype(np.int64)
        return torch.from_numpy(lt), torch.from_numpy(le - lt)


CASES = [
    # name, dtype, per-rank sizes (by world), distribution, descending, with values
    ("uniform_u32", np.uint32, {2: [5000, 5000], 3: [4000, 100, 7001]}, "uniform", False, True),
    ("few16_u32_desc", np.uint32, {2: [3000, 4000], 3: [1000, 2000, 3000]}, "few16", True, True),
    ("equal_u64", np.uint64, {2: [2500, 1500], 3: [700, 700, 701]}, "equal", False, True),
    ("entropy5_u64", np.uint64, {2: [4096, 4097], 3: [3000, 0, 3000]}, "entropy5", False, True),
    ("empty_rank", np.uint32, {2: [0, 3000], 3: [0, 0, 10]}, "uniform", False, True),
    ("one_item", np.int32, {2: [1, 0], 3: [0, 1, 0]}, "uniform", False, True),
    ("all_empty", np.uint32, {2: [0, 0], 3: [0, 0, 0]}, "uniform", False, False),
    ("f32_zeros", np.float32, {2: [2000, 2100], 3: [900, 1000, 1100]}, "uniform", True, True),
    ("i64_keys_only", np.int64, {2: [3000, 2000], 3: [10, 2000, 30]}, "uniform", False, False),
    ("u8_keys", np.uint8, {2: [3000, 2000], 3: [1000, 2000, 300]}, "uniform", False, True),
    ("extremes_u32", np.uint32, {2: [600, 600], 3: [400, 400, 400]}, "extremes", False, True),
]


def _worker(rank, world, port, results, protocol="partition"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cccl_b200.multi_gpu import distributed_sort

    failures = []
    for name, dtype, sizes, dist_name, desc, with_vals in CASES:
        ns = sizes[world]
        shards, vshards = [], []
        base = 0
        for r, n in enumerate(ns):
            if dist_name == "extremes":
                k = np.random.default_rng(100 + r).choice(np.array([0, 1, 2**32 - 2, 2**32 - 1], dtype=np.uint32), n)
            else:
                k = make_keys(dist_name, n, dtype, seed=31 * r + 7)
            if np.dtype(dtype).kind == "f" and n:
                k[::17] = -0.0
                k[::19] = 0.0
            shards.append(k)
            vshards.append(np.arange(base, base + n, dtype=np.uint32))
            base += n
        stats = {}
        ok, ov = distributed_sort(to_torch(shards[rank]), to_torch(vshards[rank]) if with_vals else None,
                                  descending=desc, ops=OracleOps(), stats=stats, protocol=protocol)
        allk, allv = np.concatenate(shards), np.concatenate(vshards)
        if with_vals:
            ek, ev = oracle_sort(allk, allv, descending=desc)
        else:
            ek, ev = oracle_sort(allk, descending=desc), None
        lo = int(np.sum(ns[:rank]))
        hi = lo + ns[rank]
        gk = to_np(ok)
        if gk.shape[0] != ns[rank] or not np.array_equal(gk.view(np.uint8), ek[lo:hi].view(np.uint8)):
            failures.append((name, "keys"))
        if with_vals and not np.array_equal(to_np(ov), ev[lo:hi]):
            failures.append((name, "values"))
    results[rank] = failures
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("protocol", ["partition", "sort"])
@pytest.mark.parametrize("world", [2, 3])
def test_distributed_sort_host_logic_gloo(world, protocol):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results, protocol), nprocs=world, join=True)
    for r in range(world):
        assert results[r] == [], f"rank {r}: {results[r]}"


def test_untwiddle_inverts_the_kernel_transform():
    from cccl_b200.multi_gpu import untwiddle

    rng = np.random.default_rng(3)
    for kind, kb in ((0, 1), (0, 4), (1, 2), (1, 8), (2, 4), (2, 8)):
        for desc in (False, True):
            bits = rng.integers(0, 2 ** (8 * kb), size=1000, dtype=np.uint64) if kb < 8 else \
                rng.integers(0, 2**63, size=1000, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
            allm = np.uint64(2 ** (8 * kb) - 1) if kb < 8 else np.uint64(0xFFFFFFFFFFFFFFFF)
            high = np.uint64(1 << (8 * kb - 1))
            if kind == 2:
                m = np.where((bits & high) != 0, allm, high)
            elif kind == 1:
                m = np.full_like(bits, high)
            else:
                m = np.zeros_like(bits)
            t = (bits ^ m) & allm
            if desc:
                t = t ^ allm
            assert np.array_equal(untwiddle(t, kind, kb, desc), bits), (kind, kb, desc)

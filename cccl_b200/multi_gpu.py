"""One-box multi-GPU stable radix sort: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch)
for the plumbing, the sm_100a kernels of libb200rs.so for every pass over the data.

What it computes: the concatenation of the per-rank outputs in rank order equals ONE stable
``cub::DeviceRadixSort::SortPairs`` of the concatenation of the per-rank inputs in rank order, keys and values bit
for bit; rank r's output has exactly as many items as its input.  The reference's own multi-GPU sort,
``cudax::sort`` (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/hss/execute.h:56-129), is
keys-only and unstable; its protocol (local sort -> splitters by global counting -> all-to-all -> local merge/sort,
hss/histogramming.h:522-610, hss/data_exchange.h:380-450) is what this restates for radix keys, with EXACT splitters.

Protocol
  1. local stable sort of the shard (b200rs_sort, DoubleBuffer form);
  2. exact splitters by MSD radix select over the bit-ordered key space: per round every rank binary-searches 257 bin
     boundaries per splitter in its SORTED shard (b200rs_splitter_ranks), one all-reduce sums the counts, the bin
     holding the target rank is kept; key_bytes rounds.  Ties of the splitter key are assigned by (source rank, local
     position), so partitions are exact (any duplicates, all-equal keys included) and stability is kept;
  3. key/value all-to-all-v (NCCL send/recv over NVLink; receive buffer in source-rank order);
  4. final local stable sort of the received runs (source-rank order + stable sort == global stable order).

No collective is used inside the sort passes themselves (the path shards; the exchange step is its only data-path
collective).  ``ops`` abstracts the two device primitives so the host logic can be exercised on CPU with gloo in
tests/; the default is the CUDA implementation, which raises if libb200rs.so is missing (no fallback).
"""
from __future__ import annotations

import numpy as np

from . import _native
from .radix_sort import key_kind_of, _torch_np_dtype

RADIX_BITS = 8
RADIX = 256


# ----------------------------------------------------------------------------------------------- bit-ordered keys
def untwiddle(t: np.ndarray, kind: int, key_bytes: int, descending: bool) -> np.ndarray:
    """Inverse of the kernels' key transform (cccl_b200/csrc/common.cuh twiddle_out; reference
    cub/cub/util_type.cuh:857-865, :906-914, :953-963 + radix_rank_sort_operations.cuh:545-565): bit-ordered
    unsigned value (uint64 array) -> user key bits (uint64 array, low key_bytes*8 bits)."""
    bits = key_bytes * 8
    all_ones = np.uint64((1 << bits) - 1)
    high = np.uint64(1 << (bits - 1))
    y = t.astype(np.uint64)
    if descending:
        y = y ^ all_ones
    if kind == _native.KEY_FLOAT:
        neg = (y & high) == 0  # bit-ordered values below the midpoint are negative floats (stored inverted)
        m = np.where(neg, all_ones, high)
    elif kind == _native.KEY_INT:
        m = np.full_like(y, high)
    else:
        m = np.zeros_like(y)
    return (y ^ m) & all_ones


class CudaOps:
    """The two device primitives of the protocol, through the C ABI (include/b200rs.h)."""

    def __init__(self):
        self.lib = _native.lib()  # raises if the CUDA library is absent: there is no CPU path

    def sort_pairs(self, keys, values, descending, preserve_input=False):
        import torch

        from .radix_sort import DoubleBuffer, SortOrder, radix_sort

        n = keys.numel()
        if n == 0:
            return keys, values
        order = SortOrder.DESCENDING if descending else SortOrder.ASCENDING
        if preserve_input:
            # pointer API: the caller's shard is left untouched (device_radix_sort.cuh:315)
            okeys = torch.empty_like(keys)
            ovals = torch.empty_like(values) if values is not None else None
            keep = radix_sort(d_in_keys=keys, d_out_keys=okeys, d_in_values=values, d_out_values=ovals, num_items=n,
                              order=order)
            del keep
            return okeys, ovals
        kb = DoubleBuffer(keys, torch.empty_like(keys))
        vb = DoubleBuffer(values, torch.empty_like(values)) if values is not None else None
        keep = radix_sort(d_in_keys=kb, d_out_keys=None, d_in_values=vb, d_out_values=None, num_items=n, order=order)
        del keep  # stream-ordered: the caching allocator reuses it only for later work on the same stream
        return kb.current(), (vb.current() if vb is not None else None)

    def splitter_ranks(self, sorted_keys, probes_bits: np.ndarray, descending):
        """probes_bits: uint64 numpy array of user-domain key bit patterns.  Returns (lt, eq) int64 CUDA tensors."""
        import torch

        kdt = _torch_np_dtype(sorted_keys)
        m = int(probes_bits.size)
        host = probes_bits.astype(np.dtype(f"u{kdt.itemsize}")).view(np.uint8)
        d_probes = torch.from_numpy(host.copy()).to(sorted_keys.device, non_blocking=True)
        out = torch.empty(2 * m, dtype=torch.int64, device=sorted_keys.device)
        rc = self.lib.b200rs_splitter_ranks(
            sorted_keys.data_ptr() if sorted_keys.numel() else 0, sorted_keys.numel(), key_kind_of(kdt), kdt.itemsize,
            int(bool(descending)), d_probes.data_ptr(), m, out.data_ptr(), out.data_ptr() + 8 * m,
            torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "b200rs_splitter_ranks")
        return out[:m], out[m:]


# ----------------------------------------------------------------------------------------------- splitter selection
def select_splitters(sorted_keys, targets, *, kind, key_bytes, descending, ops, group, dist, stats=None):
    """For every global target rank t (number of items that must end up on lower ranks) find the bit-ordered splitter
    value x = the (t)-th smallest key of the whole job (1-based; t == 0 gives x = 0), and return, per target, this
    rank's boundary position in its sorted shard plus the (world x targets) matrix of boundary positions of all ranks.
    """
    import torch

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nt = len(targets)
    device = sorted_keys.device
    if nt == 0:
        return np.zeros((world, 0), dtype=np.int64)
    bits = key_bytes * 8
    prefix = np.zeros(nt, dtype=np.uint64)  # high bits decided so far (value of the digits, not shifted)
    tgt = np.asarray(targets, dtype=np.int64)
    j = np.arange(RADIX + 1, dtype=np.uint64)
    rounds = 0
    for rnd in range(key_bytes):
        rem = bits - RADIX_BITS * (rnd + 1)  # bits still undecided after this round
        # boundaries of the 256 bins below the current prefix, as bit-ordered values; the 257th may be 2^bits
        cand = ((prefix[:, None] << np.uint64(RADIX_BITS)) + j[None, :])  # (nt, 257); may wrap where it overflows
        # the upper boundary of the last bin under an all-ones prefix is 2^bits: not representable, counts everything
        top_prefix = np.uint64((1 << (RADIX_BITS * rnd)) - 1)
        overflow = (j[None, :] == np.uint64(RADIX)) & (prefix[:, None] == top_prefix)
        bound = np.where(overflow, np.uint64(0), cand << np.uint64(rem))
        probes = untwiddle(bound.reshape(-1), kind, key_bytes, descending)
        lt, _ = ops.splitter_ranks(sorted_keys, probes, descending)
        lt = lt.clone()
        # the boundary 2^bits (above every key) counts the whole shard
        if overflow.any():
            lt[torch.from_numpy(overflow.reshape(-1)).to(device)] = sorted_keys.numel()
        _all_reduce(dist, lt, group)
        g = lt.cpu().numpy().reshape(nt, RADIX + 1)
        # bin b holds global ranks (g[b], g[b+1]]; keep the first bin whose upper count reaches the target
        pick = np.empty(nt, dtype=np.uint64)
        for i in range(nt):
            b = int(np.searchsorted(g[i, 1:], max(int(tgt[i]), 1), side="left"))
            pick[i] = min(b, RADIX - 1)
        prefix = (prefix << np.uint64(RADIX_BITS)) + pick
        rounds += 1
    # prefix is now the full bit-ordered splitter value; exact local counts below / equal
    probes = untwiddle(prefix, kind, key_bytes, descending)
    lt, eq = ops.splitter_ranks(sorted_keys, probes, descending)
    mine = torch.stack([lt, eq]).contiguous()
    allc = [torch.empty_like(mine) for _ in range(world)]
    _all_gather(dist, allc, mine, group)
    allc = torch.stack(allc).cpu().numpy()  # (world, 2, nt)
    lt_all, eq_all = allc[:, 0, :], allc[:, 1, :]
    need = tgt - lt_all.sum(axis=0)  # how many items EQUAL to the splitter go to lower ranks
    before = np.cumsum(eq_all, axis=0) - eq_all  # equal items held by lower source ranks
    take = np.clip(need[None, :] - before, 0, eq_all)
    if stats is not None:
        stats["splitter_rounds"] = rounds
        stats["splitters_bit_ordered"] = [int(x) for x in prefix]
    assert (take.sum(axis=0) == np.clip(need, 0, None)).all(), "splitter selection is inconsistent"
    return lt_all + take  # (world, nt) boundary positions


class _Phases:
    """CUDA-event phase timer (device time on the current stream); a no-op when stats are not requested or on CPU."""

    def __init__(self, enabled, torch):
        self.on, self.torch, self.marks = enabled, torch, []

    def mark(self, name):
        if self.on:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def result(self):
        if not self.on or len(self.marks) < 2:
            return {}
        self.marks[-1][1].synchronize()
        return {self.marks[i + 1][0]: self.marks[i][1].elapsed_time(self.marks[i + 1][1])
                for i in range(len(self.marks) - 1)}


def distributed_sort(keys, values=None, *, descending=False, group=None, ops=None, stats=None):
    """Stable distributed sort of one shard per rank; returns (keys_out, values_out) with len == len(keys).
    The caller's shard is left untouched.  ``stats`` (a dict) receives splitters, exchange counts and, on CUDA, the
    device time of each phase."""
    import torch
    import torch.distributed as dist

    ops = ops or CudaOps()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    kdt = _torch_np_dtype(keys)
    kind, key_bytes = key_kind_of(kdt), kdt.itemsize
    n_local = keys.numel()
    ph = _Phases(stats is not None and keys.is_cuda, torch)
    ph.mark("start")

    # 1. local stable sort (the caller's shard is not modified)
    skeys, svals = ops.sort_pairs(keys, values, descending, preserve_input=True)
    ph.mark("local_sort")
    if world == 1:
        if stats is not None:
            stats["phase_ms"] = ph.result()
        return skeys, svals

    # 2. exact splitters: rank r must end with the global stable positions [sum(n[:r]), sum(n[:r+1]))
    counts = torch.tensor([n_local], dtype=torch.int64, device=keys.device)
    allcounts = [torch.empty_like(counts) for _ in range(world)]
    _all_gather(dist, allcounts, counts, group)
    n_all = np.array([int(c.item()) for c in allcounts], dtype=np.int64)
    targets = np.cumsum(n_all)[:-1]
    bounds = select_splitters(skeys, targets, kind=kind, key_bytes=key_bytes, descending=descending, ops=ops,
                              group=group, dist=dist, stats=stats)
    # boundary matrix with the implicit 0 and n columns: items [edges[i, r], edges[i, r+1]) of source i go to rank r
    edges = np.concatenate([np.zeros((world, 1), dtype=np.int64), bounds, n_all[:, None]], axis=1)
    send = (edges[rank, 1:] - edges[rank, :-1]).tolist()
    recv = (edges[:, rank + 1] - edges[:, rank]).tolist()
    assert sum(recv) == n_local and min(send) >= 0 and min(recv) >= 0
    ph.mark("splitters")

    # 3. all-to-all-v, receive buffer in source-rank order
    rkeys = torch.empty(n_local, dtype=skeys.dtype, device=skeys.device)
    _all_to_all(dist, rkeys, skeys, recv, send, group)
    rvals = None
    if svals is not None:
        rvals = torch.empty(n_local, dtype=svals.dtype, device=svals.device)
        _all_to_all(dist, rvals, svals, recv, send, group)
    ph.mark("exchange")

    # 4. final local stable sort
    out = ops.sort_pairs(rkeys, rvals, descending)
    ph.mark("final_sort")
    if stats is not None:
        stats["send_counts"] = send
        stats["recv_counts"] = recv
        item = key_bytes + (svals.element_size() if svals is not None else 0)
        stats["exchange_bytes_out"] = (n_local - send[rank]) * item
        stats["exchange_bytes_in"] = (n_local - recv[rank]) * item
        stats["phase_ms"] = ph.result()
    return out


_A2A_VIEW = {1: "uint8", 2: "int16", 4: "int32", 8: "int64"}


def _staged(dist, group, t):
    """gloo moves host memory: CUDA tensors are staged through the CPU (tests run several ranks on ONE GPU this way;
    the NCCL path hands device tensors straight to the collective)."""
    return t.is_cuda and dist.get_backend(group) == "gloo"


def _all_to_all(dist, out, inp, out_splits, in_splits, group):
    """all_to_all_single on a same-width signed view (NCCL/gloo do not take every unsigned dtype)."""
    import torch

    view = getattr(torch, _A2A_VIEW[inp.element_size()])
    if _staged(dist, group, inp):
        tmp = torch.empty(out.numel(), dtype=view)
        dist.all_to_all_single(tmp, inp.view(view).cpu(), list(out_splits), list(in_splits), group=group)
        out.view(view).copy_(tmp)
        return
    dist.all_to_all_single(out.view(view), inp.view(view), list(out_splits), list(in_splits), group=group)


def _all_reduce(dist, t, group):
    if _staged(dist, group, t):
        c = t.cpu()
        dist.all_reduce(c, group=group)
        t.copy_(c)
        return
    dist.all_reduce(t, group=group)


def _all_gather(dist, outs, t, group):
    if _staged(dist, group, t):
        couts = [o.cpu() for o in outs]
        dist.all_gather(couts, t.cpu(), group=group)
        for o, c in zip(outs, couts):
            o.copy_(c)
        return
    dist.all_gather(outs, t, group=group)

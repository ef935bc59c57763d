"""One-box multi-GPU stable radix sort: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 / NVSwitch)
for the plumbing, the sm_100a kernels of libb200rs.so for every pass over the data.

What it computes: the concatenation of the per-rank outputs in rank order equals ONE stable
``cub::DeviceRadixSort::SortPairs`` of the concatenation of the per-rank inputs in rank order, keys and values bit
for bit; rank r's output has exactly as many items as its input.  The reference's own multi-GPU sort,
``cudax::sort`` (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/hss/execute.h:56-129), is
keys-only and unstable; its protocol (local sort -> splitters by global counting -> all-to-all -> local merge/sort,
hss/histogramming.h:522-610, hss/data_exchange.h:380-450) is what this restates for radix keys, with EXACT splitters.

Protocol "partition" (default; SURVEY.md 8e as written: histogram rounds -> partition by destination -> exchange ->
ONE local sort)
  1. exact splitters by MSD radix select over the UNSORTED shard: per round one 256-bin histogram of the next digit
     among the keys carrying each candidate prefix (round 0: the upsweep histogram kernel; later rounds
     b200rs_select_histogram), one all-reduce, the bin holding the target rank is kept; key_bytes rounds.  The local
     counts below / equal to each splitter fall out of the same rounds;
  2. destination buckets (b200rs_bucket_ids: keys strictly between two splitters and keys tied with a splitter get
     separate buckets) and ONE stable partition pass by bucket (the 1-byte-key SortPairs path, <= 5 bits): every
     destination's items are then contiguous, in local order.  Ties of a splitter key are split by (source rank, local
     position), so partitions are exact for any duplicates and stability is kept;
  3. key/value exchange: direct NVLink peer copies into the destination's receive buffer (torch symmetric memory)
     when available, else all-to-all-v over NCCL; receive buffer in source-rank order;
  4. ONE local stable sort of the received items (source-rank order + stable sort == global stable order).

Protocol "sort" (the first implementation, kept selectable and tested)
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
        self._scratch = {}
        self.kernel_launches = 0  # kernels of libb200rs.so launched through this object (bench.py's gpu_launches)

    def scratch(self, name, numel, dtype, device):
        """Grow-only scratch tensor kept across calls (stream-ordered reuse: one stream per CudaOps object).  Scratch
        never leaves this module; results handed to the caller are always fresh or caller-provided tensors."""
        import torch

        t = self._scratch.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != device:
            t = torch.empty(max(int(numel), 1), dtype=dtype, device=device)
            self._scratch[name] = t
        return t[:numel]

    def sort_pairs(self, keys, values, descending, preserve_input=False, out=None):
        import torch

        from .radix_sort import DoubleBuffer, SortOrder, make_radix_sort, radix_sort

        n = keys.numel()
        if n == 0:
            return keys, values
        order = SortOrder.DESCENDING if descending else SortOrder.ASCENDING
        if preserve_input:
            # pointer API: the caller's shard is left untouched (device_radix_sort.cuh:315); temp storage is scratch
            okeys = out[0] if out is not None else torch.empty_like(keys)
            ovals = (out[1] if out is not None else torch.empty_like(values)) if values is not None else None
            sorter = make_radix_sort(d_in_keys=keys, d_out_keys=okeys, d_in_values=values, d_out_values=ovals,
                                     order=order)
            kw = dict(d_in_keys=keys, d_out_keys=okeys, d_in_values=values, d_out_values=ovals, num_items=n)
            nbytes = sorter(temp_storage=None, **kw)
            sorter(temp_storage=self.scratch("sort_temp", nbytes, torch.uint8, keys.device), **kw)
            self.kernel_launches += self._sort_kernels()
            return okeys, ovals
        kb = DoubleBuffer(keys, torch.empty_like(keys))
        vb = DoubleBuffer(values, torch.empty_like(values)) if values is not None else None
        keep = radix_sort(d_in_keys=kb, d_out_keys=None, d_in_values=vb, d_out_values=None, num_items=n, order=order)
        del keep  # stream-ordered: the caching allocator reuses it only for later work on the same stream
        self.kernel_launches += self._sort_kernels()
        return kb.current(), (vb.current() if vb is not None else None)

    def _sort_kernels(self):
        # stream ops of the last b200rs_sort call minus its one memset (single-tile sorts have none)
        n = self.lib.b200rs_last_launch_count()
        return n - 1 if n > 1 else n

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


    # ---- primitives of the partition-first protocol
    def top_digit_histogram(self, keys, descending):
        """(1, 256) int64: histogram of the most significant 8-bit digit of the bit-ordered keys (upsweep kernel)."""
        import torch

        kdt = _torch_np_dtype(keys)
        out = torch.empty(RADIX, dtype=torch.int64, device=keys.device)
        bits = kdt.itemsize * 8
        rc = self.lib.b200rs_digit_histogram(
            keys.data_ptr() if keys.numel() else 0, keys.numel(), key_kind_of(kdt), kdt.itemsize, bits - RADIX_BITS,
            bits, int(bool(descending)), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "b200rs_digit_histogram")
        self.kernel_launches += 1
        return out.view(1, RADIX)

    def select_histogram(self, keys, prefixes, rnd: int, descending, candidates="none"):
        """(len(prefixes), 256) int64: per candidate prefix (int64 device tensor holding the bit pattern of the high
        digits chosen so far), histogram of digit `rnd` (MSD first) of the keys whose higher digits equal the prefix.
        candidates: "emit" also compacts the keys that carry any prefix into a scratch buffer, "use" scans that buffer
        (written by the previous "emit" over the same keys) instead of all keys, "none" does neither."""
        import torch

        kdt = _torch_np_dtype(keys)
        m = int(prefixes.numel())
        n = keys.numel()
        out = torch.empty((m, RADIX), dtype=torch.int64, device=keys.device)
        cap = n // 8 + (1 << 16)
        cin = cst_in = cout = cst_out = 0
        if candidates != "none" and n:
            cand = self.scratch("select_cand", cap, keys.dtype, keys.device)
            state = self.scratch("select_state", 2 + 1024, torch.int64, keys.device)
            if candidates == "emit":
                cout, cst_out = cand.data_ptr(), state.data_ptr()
            else:
                cin, cst_in = cand.data_ptr(), state.data_ptr()
        rc = self.lib.b200rs_select_histogram(
            keys.data_ptr() if n else 0, n, key_kind_of(kdt), kdt.itemsize, int(bool(descending)),
            prefixes.data_ptr(), m, rnd, out.data_ptr(), cin or None, cst_in or None, cout or None, cst_out or None,
            cap, torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "b200rs_select_histogram")
        self.kernel_launches += 1
        return out

    def bucket_ids(self, keys, splitters: np.ndarray, descending):
        """uint8 tensor: 2 * #{splitters below the key} + [key equals a splitter] (bit-ordered domain)."""
        import ctypes

        import torch

        kdt = _torch_np_dtype(keys)
        m = int(splitters.size)
        ids = torch.empty(keys.numel(), dtype=torch.uint8, device=keys.device)
        arr = (ctypes.c_uint64 * max(m, 1))(*[int(x) for x in splitters])
        rc = self.lib.b200rs_bucket_ids(
            keys.data_ptr() if keys.numel() else 0, keys.numel(), key_kind_of(kdt), kdt.itemsize,
            int(bool(descending)), arr, m, ids.data_ptr() if keys.numel() else 0,
            torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "b200rs_bucket_ids")
        self.kernel_launches += 1
        return ids

    def partition_by_splitters(self, keys, values, splitters: np.ndarray, sizes: np.ndarray, descending):
        """Stable partition of (keys, values) into destination-bucket order in ONE onesweep launch whose digit is the
        bucket of the key (b200rs_partition_by_splitters).  Returns None when the library has no such kernel for these
        key / value widths."""
        import ctypes

        import torch

        n = keys.numel()
        kdt = _torch_np_dtype(keys)
        vb = values.element_size() if values is not None else 0
        m = int(splitters.size)
        sp = (ctypes.c_uint64 * max(m, 1))(*[int(x) for x in splitters])
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        bo = (ctypes.c_uint64 * len(offs))(*[int(x) for x in offs])
        stream = torch.cuda.current_stream().cuda_stream
        nbytes = ctypes.c_size_t(0)
        args = (n, key_kind_of(kdt), kdt.itemsize, vb, int(bool(descending)), sp, m, bo, stream)
        rc = self.lib.b200rs_partition_by_splitters(None, ctypes.byref(nbytes), None, None, None, None, *args)
        if rc == 801:  # cudaErrorNotSupported
            return None
        _native.check(rc, "b200rs_partition_by_splitters (size query)")
        pk = self.scratch("part_keys", n, keys.dtype, keys.device)
        pv = self.scratch("part_vals", n, values.dtype, values.device) if values is not None else None
        if n == 0:
            return pk, pv
        temp = self.scratch("part_temp", nbytes.value, torch.uint8, keys.device)
        rc = self.lib.b200rs_partition_by_splitters(
            temp.data_ptr(), ctypes.byref(nbytes), keys.data_ptr(), pk.data_ptr(),
            values.data_ptr() if values is not None else None, pv.data_ptr() if pv is not None else None, *args)
        _native.check(rc, "b200rs_partition_by_splitters")
        self.kernel_launches += 1
        return pk, pv

    def partition_to_peers(self, keys, values, splitters: np.ndarray, sizes: np.ndarray, descending, seg_ends,
                           dst_keys, dst_vals):
        """The partition pass fused with the exchange (b200rs_partition_to_peers): segment r of the partitioned order
        (cut at seg_ends) is stored straight through dst_keys[r] / dst_vals[r] (biased device addresses of rank r's
        receive buffers).  Returns False when the library has no bucket-mode kernel for these widths."""
        import ctypes

        import torch

        n = keys.numel()
        kdt = _torch_np_dtype(keys)
        vb = values.element_size() if values is not None else 0
        m, nd = int(splitters.size), len(dst_keys)
        u64s = lambda xs: (ctypes.c_uint64 * max(len(xs), 1))(*[int(x) & (2**64 - 1) for x in xs])
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        args = (n, key_kind_of(kdt), kdt.itemsize, vb, int(bool(descending)), u64s(splitters), m, u64s(offs), nd,
                u64s(seg_ends), u64s(dst_keys), u64s(dst_vals if values is not None else []),
                torch.cuda.current_stream().cuda_stream)
        nbytes = ctypes.c_size_t(0)
        rc = self.lib.b200rs_partition_to_peers(None, ctypes.byref(nbytes), None, None, *args)
        if rc == 801:  # cudaErrorNotSupported
            return False
        _native.check(rc, "b200rs_partition_to_peers (size query)")
        if n == 0:
            return True
        temp = self.scratch("part_temp", nbytes.value, torch.uint8, keys.device)
        rc = self.lib.b200rs_partition_to_peers(temp.data_ptr(), ctypes.byref(nbytes), keys.data_ptr(),
                                                values.data_ptr() if values is not None else None, *args)
        _native.check(rc, "b200rs_partition_to_peers")
        self.kernel_launches += 1
        return True

    def partition_to_peers_supported(self, key_dtype, value_bytes):
        """True when the library has a bucket-mode kernel for these widths (a property of the types alone, so every
        rank answers alike)."""
        import ctypes

        kdt = np.dtype(key_dtype)
        one = (ctypes.c_uint64 * 1)(0)
        nbytes = ctypes.c_size_t(0)
        rc = self.lib.b200rs_partition_to_peers(None, ctypes.byref(nbytes), None, None, 1, key_kind_of(kdt),
                                                kdt.itemsize, int(value_bytes), 0, one, 0, one, 1, one, one, one, None)
        return rc == 0

    def partition(self, ids, nbits, keys, values):
        """Stable partition of (keys, values) by the bucket id of each item: one radix pass over the low `nbits` bits
        of the 1-byte ids with the payload as the value (b200rs_sort, pointer form)."""
        import torch

        from .radix_sort import SortOrder, radix_sort

        n = ids.numel()
        outs = []
        ids_out = torch.empty_like(ids)
        for payload in (keys, values):
            if payload is None:
                outs.append(None)
                continue
            out = torch.empty_like(payload)
            if n:
                view = getattr(torch, _A2A_VIEW[payload.element_size()])
                keep = radix_sort(d_in_keys=ids, d_out_keys=ids_out, d_in_values=payload.view(view),
                                  d_out_values=out.view(view), num_items=n, order=SortOrder.ASCENDING, begin_bit=0,
                                  end_bit=max(int(nbits), 1))
                del keep
            outs.append(out)
        return outs[0], outs[1]


# ----------------------------------------------------------------------------------------------- splitter selection
def select_splitters_unsorted(keys, targets, *, kind, key_bytes, descending, ops, group, dist, stats=None,
                              top_hist=None):
    """Exact splitters from the UNSORTED shard.  For every global target rank t finds the bit-ordered value x of the
    t-th smallest key of the whole job by MSD radix select, and returns (bounds, splitters, sizes): bounds[(world, nt)]
    = number of items of each source rank that go to ranks <= the target's, counted in the order "bucket id, then local
    position"; splitters = the distinct x values, ascending (uint64); sizes = this rank's bucket sizes in bucket-id
    order.  The rounds keep their state on the device (the bin pick is a handful of tiny tensor ops), so the host only
    waits once, at the end."""
    import torch

    world = dist.get_world_size(group)
    nt = len(targets)
    if nt == 0:
        return np.zeros((world, 0), dtype=np.int64), np.zeros(0, dtype=np.uint64), np.array([keys.numel()])
    dev = keys.device
    tgt = torch.from_numpy(np.maximum(np.asarray(targets, dtype=np.int64), 1)).to(dev)
    prefix = torch.zeros(nt, dtype=torch.int64, device=dev)    # bit pattern of the digits chosen so far
    below = torch.zeros(nt, dtype=torch.int64, device=dev)     # global count of keys whose high digits are below it
    lt_local = torch.zeros(nt, dtype=torch.int64, device=dev)  # the same, for this rank's shard only
    eq_local = torch.zeros(nt, dtype=torch.int64, device=dev)
    for rnd in range(key_bytes):
        if rnd == 0:
            h0 = top_hist if top_hist is not None else ops.top_digit_histogram(keys, descending)
            h_local = h0.expand(nt, RADIX)
        else:
            # the first full scan compacts the keys that can still matter; later rounds only look at those
            mode = "none" if key_bytes <= 2 else ("emit" if rnd == 1 else "use")
            h_local = ops.select_histogram(keys, prefix, rnd, descending, candidates=mode)
        h_glob = h_local.contiguous().clone()
        _all_reduce(dist, h_glob, group)
        cum_g = torch.cumsum(h_glob, dim=1)
        cum_l = torch.cumsum(h_local, dim=1)
        # first bin whose global running count reaches the (remaining) target rank
        b = torch.searchsorted(cum_g, (tgt - below).unsqueeze(1)).squeeze(1).clamp_(max=RADIX - 1)
        prev = (b - 1).clamp_(min=0).unsqueeze(1)
        has_prev = b > 0
        below = below + torch.where(has_prev, cum_g.gather(1, prev).squeeze(1), torch.zeros_like(below))
        lt_local = lt_local + torch.where(has_prev, cum_l.gather(1, prev).squeeze(1), torch.zeros_like(below))
        eq_local = h_local.gather(1, b.unsqueeze(1)).squeeze(1)
        prefix = prefix * RADIX + b  # wraps like the unsigned 64-bit value it stands for
    mine = torch.stack([lt_local, eq_local, prefix]).contiguous()
    allc = [torch.empty_like(mine) for _ in range(world)]
    _all_gather(dist, allc, mine, group)
    allc = torch.stack(allc).cpu().numpy()  # (world, 3, nt): the only host wait of the selection
    lt_all, eq_all = allc[:, 0, :], allc[:, 1, :]
    rank = dist.get_rank(group)
    lt_mine, eq_mine = lt_all[rank], eq_all[rank]
    prefix_np = allc[rank, 2, :].view(np.uint64) & np.uint64((1 << (8 * key_bytes)) - 1 if key_bytes < 8 else 2**64 - 1)
    need = np.asarray(targets, dtype=np.int64) - lt_all.sum(axis=0)  # items EQUAL to the splitter that go to lower ranks
    before = np.cumsum(eq_all, axis=0) - eq_all
    take = np.clip(need[None, :] - before, 0, eq_all)
    if stats is not None:
        stats["splitter_rounds"] = key_bytes
        stats["splitters_bit_ordered"] = [int(x) for x in prefix_np]
    assert (take.sum(axis=0) == np.clip(need, 0, None)).all(), "splitter selection is inconsistent"
    splitters, first = np.unique(prefix_np, return_index=True)
    # local bucket sizes in bucket-id order: (below u0), (== u0), (between u0 and u1), (== u1), ..., (above the last)
    lt_u, eq_u = lt_mine[first], eq_mine[first]
    sizes = np.zeros(2 * len(splitters) + 1, dtype=np.int64)
    prev_end = 0
    for i in range(len(splitters)):
        sizes[2 * i] = lt_u[i] - prev_end
        sizes[2 * i + 1] = eq_u[i]
        prev_end = lt_u[i] + eq_u[i]
    sizes[-1] = keys.numel() - prev_end
    assert sizes.min() >= 0, "bucket sizes are inconsistent"
    return lt_all + take, splitters, sizes


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


class _PeerBuffers:
    """Receive buffers every rank of the group can write directly over NVLink (torch symmetric memory: CUDA VMM
    allocations mapped into every peer).  One byte buffer per payload (keys, values), grown on demand, cached per
    (group, device) because the rendezvous is a collective that costs milliseconds."""

    _cache = {}

    def __init__(self, nbytes, device, group, dist):
        import torch
        import torch.distributed._symmetric_memory as symm

        self.nbytes = int(nbytes)
        self.world = dist.get_world_size(group)
        self.bufs, self.hdls, self.peers = [], [], []
        for _ in range(2):
            buf = symm.empty(self.nbytes, dtype=torch.uint8, device=device)
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            self.bufs.append(buf)
            self.hdls.append(hdl)
            self.peers.append([hdl.get_buffer(r, (self.nbytes,), torch.uint8, 0) for r in range(self.world)])
        self.ptrs = [[int(p) for p in hdl.buffer_ptrs] for hdl in self.hdls]  # [payload][rank] device addresses

    @classmethod
    def get(cls, nbytes, device, group, dist):
        key = (id(group), str(device))
        cur = cls._cache.get(key)
        if cur is None or cur.nbytes < nbytes:
            cls._cache[key] = cur = cls(max(int(nbytes), 1 << 20), device, group, dist)
        return cur


def _peer_exchange(dist, group, rank, edges, payloads, n_local, max_local):
    """Every rank copies each destination's contiguous segment straight into that destination's receive buffer
    (cudaMemcpyAsync device-to-peer over NVLink, in stream order), then a device-side barrier; the receive buffer is
    laid out in source-rank order.  Returns views of this rank's receive buffers."""
    world = edges.shape[0]
    item = max(p.element_size() for p in payloads if p is not None)
    pb = _PeerBuffers.get(int(max_local) * item, payloads[0].device, group, dist)
    counts = edges[:, 1:] - edges[:, :-1]              # (source, destination)
    dst_off = np.cumsum(counts, axis=0) - counts       # where source i's segment starts in destination r's buffer
    pb.hdls[0].barrier(channel=0)  # every peer has finished reading its receive buffers of the previous exchange
    outs = []
    for which, src in enumerate(payloads):
        if src is None:
            outs.append(None)
            continue
        es = src.element_size()
        flat = src.view(-1).view(_byte_view())
        for step in range(world):
            r = (rank + step) % world  # staggered so that the ranks do not all write to the same peer first
            cnt = int(counts[rank, r])
            if cnt == 0:
                continue
            lo, off = int(edges[rank, r]) * es, int(dst_off[rank, r]) * es
            pb.peers[which][r][off:off + cnt * es].copy_(flat[lo:lo + cnt * es], non_blocking=True)
        outs.append(pb.bufs[which][: n_local * es].view(src.dtype))
    pb.hdls[0].barrier(channel=1)  # all sources have written their segments into this rank's buffers
    return outs


def _byte_view():
    import torch

    return torch.uint8


class NativeComm:
    """The C++ multi-GPU sort (b200rs_multi_comm_* / b200rs_sort_multi, cccl_b200/csrc/multi.cu): after the collective
    creation (CUDA IPC handles all-gathered through torch.distributed, any backend) a sort is one C-ABI call that
    enqueues kernels only -- no NCCL call, no host wait.  One communicator per (group, device), grown (collectively) when
    a larger receive buffer is needed."""

    _cache = {}

    def __init__(self, receive_bytes, device, group, dist):
        import ctypes

        import torch

        self.lib = _native.lib()
        self.device, self.group = device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.receive_bytes = int(receive_bytes)
        self.kernel_launches = 0
        self._temp = None
        on_gpu = dist.get_backend(group) == "nccl"

        def allgather(_ctx, send, recv, nbytes):
            try:
                mine = torch.frombuffer(bytearray(ctypes.string_at(send, nbytes)), dtype=torch.uint8)
                mine = mine.to(device) if on_gpu else mine
                outs = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(outs, mine, group=group)
                blob = b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)
                ctypes.memmove(recv, blob, nbytes * self.world)
                return 0
            except Exception:  # noqa: BLE001  (reported to the C side as a failed all-gather)
                return 1

        self._cb = _native.ALLGATHER_FN(allgather)  # keep the callback object alive
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = self.lib.b200rs_multi_comm_create(ctypes.byref(handle), self.rank, self.world, self.receive_bytes,
                                                   self._cb, None)
        _native.check(rc, "b200rs_multi_comm_create")
        self.handle = handle

    @classmethod
    def get(cls, receive_bytes, device, group, dist):
        key = (id(group), str(device))
        cur = cls._cache.get(key)
        if cur is None or cur.receive_bytes < receive_bytes:
            if cur is not None:
                cur.close()
            cls._cache[key] = cur = cls(max(int(receive_bytes), 1 << 20), device, group, dist)
        return cur

    def close(self):
        if self.handle:
            self.lib.b200rs_multi_comm_destroy(self.handle)
            self.handle = None

    @staticmethod
    def supported(key_bytes, value_bytes):
        return key_bytes in (4, 8) and value_bytes in (0, 4, 8)

    def sort(self, keys, values, descending, out=None):
        import ctypes

        import torch

        kdt = _torch_np_dtype(keys)
        n = keys.numel()
        vb = values.element_size() if values is not None else 0
        okeys = out[0] if out is not None else torch.empty_like(keys)
        ovals = (out[1] if out is not None else torch.empty_like(values)) if values is not None else None
        stream = torch.cuda.current_stream(keys.device).cuda_stream
        p = lambda t: t.data_ptr() if t is not None and t.numel() else None
        args = (n, key_kind_of(kdt), kdt.itemsize, vb, int(bool(descending)), stream)
        nbytes = ctypes.c_size_t(0)
        with torch.cuda.device(keys.device):
            rc = self.lib.b200rs_sort_multi(self.handle, None, ctypes.byref(nbytes), None, None, None, None, *args)
            _native.check(rc, "b200rs_sort_multi (size query)")
            if self._temp is None or self._temp.numel() < nbytes.value:
                self._temp = torch.empty(nbytes.value, dtype=torch.uint8, device=keys.device)
            rc = self.lib.b200rs_sort_multi(self.handle, self._temp.data_ptr(), ctypes.byref(nbytes), p(keys), p(okeys),
                                            p(values), p(ovals), *args)
        _native.check(rc, "b200rs_sort_multi")
        self.kernel_launches += self.lib.b200rs_multi_last_launch_count(self.handle)
        return okeys, ovals

    def timing(self, on):
        _native.check(self.lib.b200rs_multi_timing_enable(self.handle, int(bool(on))), "b200rs_multi_timing_enable")

    def timing_read(self):
        """Device ms of the last sort's phases (after timing(True)): splitters, partition (+ exchange), barrier,
        final_sort."""
        import ctypes

        ms = (ctypes.c_float * 4)()
        _native.check(self.lib.b200rs_multi_timing_read(self.handle, ms), "b200rs_multi_timing_read")
        return dict(zip(("splitters", "partition", "barrier", "final_sort"), (float(x) for x in ms)))

    def status(self):
        """Waits for the device; 0 = ok (see include/b200rs.h)."""
        import ctypes

        st = ctypes.c_int(0)
        _native.check(self.lib.b200rs_multi_status(self.handle, ctypes.byref(st)), "b200rs_multi_status")
        return st.value


_DEFAULT_OPS = None


def _default_ops():
    global _DEFAULT_OPS
    if _DEFAULT_OPS is None:
        _DEFAULT_OPS = CudaOps()
    return _DEFAULT_OPS


def distributed_sort(keys, values=None, *, descending=False, group=None, ops=None, stats=None, protocol="partition",
                     exchange="auto", out=None, receive_items=None):
    """Stable distributed sort of one shard per rank; returns (keys_out, values_out) with len == len(keys).
    The caller's shard is left untouched.  ``stats`` (a dict) receives splitters, exchange counts and, on CUDA, the
    device time of each phase.  ``protocol``: "partition" (histogram select -> one partition pass -> exchange -> one
    sort) or "sort" (sort -> binary-search select -> exchange -> sort).  ``exchange``: "fused" (the partition kernel stores
    straight into the destination GPUs' receive buffers), "peer" (NVLink peer copies through symmetric memory),
    "collective" (all-to-all-v) or "auto" (fused, else peer, on CUDA + NCCL; else collective).  ``protocol="native"``:
    the same partition protocol with the host side in C++ (b200rs_sort_multi): no host wait, no NCCL call.  ``out``: optional
    (keys_out, values_out) tensors for the result (same length and dtype as the inputs)."""
    import torch
    import torch.distributed as dist

    if protocol not in ("partition", "sort", "native"):
        raise ValueError(f"unknown protocol {protocol!r}")
    world = dist.get_world_size(group)
    if protocol == "native":
        # the C++ host path: one C-ABI call, kernels only.  `receive_items` (>= the largest shard of the job, the same
        # number on every rank) sizes the receive buffers; a larger shard is reported through NativeComm.status() as a
        # capacity error, nothing is written out of bounds
        kdt0 = _torch_np_dtype(keys)
        vb0 = values.element_size() if values is not None else 0
        if not NativeComm.supported(kdt0.itemsize, vb0):
            raise ValueError("protocol='native' takes 4- and 8-byte keys with 0-, 4- or 8-byte values")
        if receive_items is None:
            # convenient default: agree on the largest shard (one small all-reduce + host wait per call); pass
            # receive_items to keep the call free of host waits
            nmax = torch.tensor([keys.numel()], dtype=torch.int64, device=keys.device)
            if _staged(dist, group, nmax):
                nmax = nmax.cpu()
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
            receive_items = int(nmax.item())
        cap_items = int(receive_items)
        comm = NativeComm.get(cap_items * max(kdt0.itemsize, vb0), keys.device, group, dist)
        ph0 = _Phases(stats is not None and keys.is_cuda, torch)
        ph0.mark("start")
        res = comm.sort(keys, values, descending, out=out)
        ph0.mark("native_sort")
        if stats is not None:
            # (asking for stats waits for the device; a call without stats enqueues kernels and returns)
            stats["protocol"], stats["exchange"] = "native", "fused"
            stats["kernel_launches_total"] = comm.kernel_launches
            stats["phase_ms"] = ph0.result()
            stats["status"] = comm.status()
            stats["comm"] = comm
        return res
    ops = ops or _default_ops()
    rank = dist.get_rank(group)
    kdt = _torch_np_dtype(keys)
    kind, key_bytes = key_kind_of(kdt), kdt.itemsize
    n_local = keys.numel()
    ph = _Phases(stats is not None and keys.is_cuda, torch)
    ph.mark("start")

    if world == 1:
        res = ops.sort_pairs(keys, values, descending, preserve_input=True, out=out)
        ph.mark("local_sort")
        if stats is not None:
            stats["phase_ms"] = ph.result()
        return res

    # the first select round only needs the shard: start it before the host waits for the other ranks' sizes
    top_hist = ops.top_digit_histogram(keys, descending) if protocol == "partition" else None
    # rank r must end with the global stable positions [sum(n[:r]), sum(n[:r+1]))
    counts = torch.tensor([n_local], dtype=torch.int64, device=keys.device)
    allcounts = [torch.empty_like(counts) for _ in range(world)]
    _all_gather(dist, allcounts, counts, group)
    n_all = torch.cat(allcounts).cpu().numpy().astype(np.int64)  # one host wait
    targets = np.cumsum(n_all)[:-1]

    if protocol == "sort":
        # 1. local stable sort (the caller's shard is not modified); 2. splitters by binary search in the sorted shard
        skeys, svals = ops.sort_pairs(keys, values, descending, preserve_input=True)
        ph.mark("local_sort")
        bounds = select_splitters(skeys, targets, kind=kind, key_bytes=key_bytes, descending=descending, ops=ops,
                                  group=group, dist=dist, stats=stats)
        ph.mark("splitters")
    else:
        # 1. splitters by radix select over the unsorted shard; 2. one stable partition pass by destination bucket
        if world - 1 > 15:
            raise ValueError("the partition protocol supports up to 16 ranks (one box)")
        bounds, splitters, sizes = select_splitters_unsorted(keys, targets, kind=kind, key_bytes=key_bytes,
                                                             descending=descending, ops=ops, group=group, dist=dist,
                                                             stats=stats, top_hist=top_hist)
        ph.mark("splitters")
    # boundary matrix with the implicit 0 and n columns: items [edges[i, r], edges[i, r+1]) of source i go to rank r
    edges = np.concatenate([np.zeros((world, 1), dtype=np.int64), bounds, n_all[:, None]], axis=1)
    send = (edges[rank, 1:] - edges[rank, :-1]).tolist()
    recv = (edges[:, rank + 1] - edges[:, rank]).tolist()
    assert sum(recv) == n_local and min(send) >= 0 and min(recv) >= 0
    use_peer = exchange in ("peer", "fused") or (exchange == "auto" and keys.is_cuda
                                                 and dist.get_backend(group) == "nccl")
    peer_done = fused_done = False

    if protocol == "partition":
        # 2 + 3 fused: ONE kernel partitions by destination and stores every item straight into the receive buffer of
        # its destination GPU (peer-mapped symmetric memory) -- the compute step and the collective that follows it
        # Eligibility is decided from GLOBALLY known facts only (the largest shard, the key / value widths), so that
        # every rank takes the same path: the code below contains cross-GPU barriers.
        es_k = keys.element_size()
        es_v = values.element_size() if values is not None else 0
        fused_ok = (use_peer and exchange != "peer" and 0 < int(n_all.max()) < (1 << 30)
                    and hasattr(ops, "partition_to_peers") and ops.partition_to_peers_supported(kdt, es_v))
        pb = None
        if fused_ok:
            try:
                pb = _PeerBuffers.get(int(n_all.max()) * max(es_k, es_v), keys.device, group, dist)
            except (RuntimeError, ImportError, AttributeError) as ex:  # symmetric memory not available (every rank alike)
                if exchange == "fused":
                    raise
                if stats is not None:
                    stats["peer_exchange_unavailable"] = repr(ex)
        if pb is not None:
            counts = edges[:, 1:] - edges[:, :-1]
            dst_off = np.cumsum(counts, axis=0) - counts
            bias = dst_off[rank] - edges[rank, :-1]  # destination index of partitioned index 0, per rank
            dst_k = [pb.ptrs[0][r] + int(bias[r]) * es_k for r in range(world)]
            dst_v = [pb.ptrs[1][r] + int(bias[r]) * es_v for r in range(world)]
            # no exception is swallowed between the two barriers: a rank that fails here fails the job
            pb.hdls[0].barrier(channel=0)  # every peer has finished reading its receive buffers
            ok_ = ops.partition_to_peers(keys, values, splitters, sizes, descending, edges[rank, 1:-1].tolist(), dst_k,
                                         dst_v)
            pb.hdls[0].barrier(channel=1)  # every source's stores have landed (needed after the first one, too)
            if not ok_:
                raise RuntimeError("b200rs_partition_to_peers refused a problem the size query accepted")
            fused_done = peer_done = True
            rkeys = pb.bufs[0][: n_local * es_k].view(keys.dtype)
            rvals = pb.bufs[1][: n_local * es_v].view(values.dtype) if values is not None else None
            ph.mark("partition")
        if not fused_done:
            fused = getattr(ops, "partition_by_splitters", None)
            part = fused(keys, values, splitters, sizes, descending) if fused is not None else None
            if part is None:
                # key / value widths the fused pass is not compiled for: bucket ids + one radix pass over the ids
                ids = ops.bucket_ids(keys, splitters, descending)
                part = ops.partition(ids, int(2 * len(splitters)).bit_length(), keys, values)
            skeys, svals = part
            ph.mark("partition")

    # 3. exchange, receive buffer in source-rank order
    if not fused_done and use_peer and int(n_all.max()) > 0:
        try:
            rkeys, rvals = _peer_exchange(dist, group, rank, edges, [skeys, svals], n_local, int(n_all.max()))
            peer_done = True
        except (RuntimeError, ImportError, AttributeError) as ex:  # symmetric memory not available on this system
            if exchange == "peer":
                raise
            if stats is not None:
                stats["peer_exchange_unavailable"] = repr(ex)
    if not peer_done:
        rkeys = torch.empty(n_local, dtype=skeys.dtype, device=skeys.device)
        _all_to_all(dist, rkeys, skeys, recv, send, group)
        rvals = None
        if svals is not None:
            rvals = torch.empty(n_local, dtype=svals.dtype, device=svals.device)
            _all_to_all(dist, rvals, svals, recv, send, group)
    ph.mark("exchange")

    # 4. final local stable sort (peer mode: the receive buffers are shared scratch, so the result goes elsewhere)
    if peer_done or out is not None:
        res = ops.sort_pairs(rkeys, rvals, descending, preserve_input=True, out=out)
    else:
        res = ops.sort_pairs(rkeys, rvals, descending)
    ph.mark("final_sort")
    if stats is not None:
        stats["protocol"] = protocol
        stats["kernel_launches_total"] = getattr(ops, "kernel_launches", None)
        stats["exchange"] = "fused" if fused_done else ("peer" if peer_done else "collective")
        stats["send_counts"] = send
        stats["recv_counts"] = recv
        item = key_bytes + (values.element_size() if values is not None else 0)
        stats["exchange_bytes_out"] = (n_local - send[rank]) * item
        stats["exchange_bytes_in"] = (n_local - recv[rank]) * item
        stats["phase_ms"] = ph.result()
    return res


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

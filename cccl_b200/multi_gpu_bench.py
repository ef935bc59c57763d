"""bench.py's N > 1 arm: distributed SortPairs of 2^k (u32 key, u32 value) pairs PER GPU (weak scaling), one process
per GPU under torchrun, exchange over NCCL/NVLink.  Prints ONE JSON line on rank 0 (the contract of bench.py)."""
from __future__ import annotations

import json
import os


def run(args, metric, workload_name, ClockSampler, measured_peaks):
    import numpy as np
    import torch
    import torch.distributed as dist

    from . import _native
    from .multi_gpu import distributed_sort

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus} "
                         f"(WORLD_SIZE is {world})")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _native.lib()  # raises if the CUDA library is missing: there is no fallback

    n = 1 << args.log2_per_gpu
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    keys = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
    vals = (torch.arange(n, dtype=torch.int32, device="cuda") + (rank * n)).view(torch.uint32)
    h_keys = torch.empty(n, dtype=torch.int32).pin_memory()
    h_vals = torch.empty(n, dtype=torch.int32).pin_memory()
    h_keys.copy_(keys.view(torch.int32))
    h_vals.copy_(vals.view(torch.int32))
    h_ok = torch.empty(n, dtype=torch.int32).pin_memory()
    h_ov = torch.empty(n, dtype=torch.int32).pin_memory()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    out_buf = (torch.empty_like(keys), torch.empty_like(vals))  # results land here every step (no allocation)
    proto = getattr(args, "multi_protocol", "native")
    native = proto == "native"
    _dsort = distributed_sort

    def distributed_sort(k, v, **kw):  # noqa: F811  (the protocol under test; equal shards: receive_items = n)
        return _dsort(k, v, protocol=proto, receive_items=n if native else None, **kw)

    stats = {}
    for _ in range(args.warmup):
        distributed_sort(keys, vals, stats=stats, out=out_buf)
    barrier()
    from .multi_gpu import _default_ops

    counter = stats["comm"] if native else _default_ops()
    before = counter.kernel_launches
    distributed_sort(keys, vals, out=out_buf)
    launches = counter.kernel_launches - before  # kernels of libb200rs.so per step on this GPU
    barrier()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {}
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            if native:
                # no stats in the timed loop: the call enqueues kernels and returns (asking for stats would wait)
                ok, ov = distributed_sort(keys, vals, out=out_buf)
                continue
            st = {}
            ok, ov = distributed_sort(keys, vals, stats=st, out=out_buf)
            for k, v in st.get("phase_ms", {}).items():
                phase.setdefault(k, []).append(v)
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if native:
        # phases of a few more (untimed) steps, from events the C++ host records at its phase boundaries
        comm = stats["comm"]
        comm.timing(True)
        for _ in range(5):
            distributed_sort(keys, vals, out=out_buf)
            for k, v in comm.timing_read().items():
                phase.setdefault(k, []).append(v)
        comm.timing(False)
        st = {"protocol": "native", "exchange": "fused", "status": comm.status()}
        assert st["status"] == 0, f"b200rs_sort_multi device-side status {st['status']}"
        # bytes this rank stores into OTHER GPUs: everything except its own 1/world of a uniform shard (exact count from
        # the verified result below is not needed for a bandwidth figure)
        st["exchange_bytes_out"] = int(n * 8 * (world - 1) / world)
        barrier()

    # per-kernel device times of the final local sort of one more (untimed) step
    lib.b200rs_timing_enable(1)
    distributed_sort(keys, vals, out=out_buf)
    final_ops = {}
    for opname, t in _native.timing_read():
        final_ops.setdefault(opname, []).append(round(t, 4))
    lib.b200rs_timing_enable(0)
    barrier()

    # end to end: pinned host shard -> device, sort, sorted shard -> pinned host, every step.  Two input and two output
    # device buffers: step i+1's H2D (copy stream) and step i's D2H (another copy stream) run while step i / i+1 sort
    # on the main stream -- PCIe is full duplex and the copy engines are idle during the sort.
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    in_bufs = [(keys, vals), (torch.empty_like(keys), torch.empty_like(vals))]
    out_bufs = [out_buf, (torch.empty_like(keys), torch.empty_like(vals))]
    in_ready = [torch.cuda.Event(), torch.cuda.Event()]
    in_free = [torch.cuda.Event(), torch.cuda.Event()]
    out_ready = [torch.cuda.Event(), torch.cuda.Event()]
    out_free = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream()

    def upload(i):
        kb, vb = in_bufs[i % 2]
        with torch.cuda.stream(s_in):
            s_in.wait_event(in_free[i % 2])  # the sort that last read this buffer pair has finished
            kb.view(torch.int32).copy_(h_keys, non_blocking=True)
            vb.view(torch.int32).copy_(h_vals, non_blocking=True)
            in_ready[i % 2].record(s_in)

    def e2e_run(steps):
        for ev in in_free + out_free:
            ev.record(main)
        upload(0)
        for i in range(steps):
            if i + 1 < steps:
                upload(i + 1)
            main.wait_event(in_ready[i % 2])
            main.wait_event(out_free[i % 2])  # the D2H that last read this output pair has finished
            k, v = distributed_sort(*in_bufs[i % 2], out=out_bufs[i % 2])
            in_free[i % 2].record(main)
            out_ready[i % 2].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(out_ready[i % 2])
                h_ok.copy_(k.view(torch.int32), non_blocking=True)
                h_ov.copy_(v.view(torch.int32), non_blocking=True)
                out_free[i % 2].record(s_out)
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    e2e_steps = max(2, min(args.steps, 6))
    e2e_run(2)
    barrier()
    ev0.record()
    e2e_run(e2e_steps)
    ev1.record()
    barrier()
    e2e_ms = torch.tensor([ev0.elapsed_time(ev1) / e2e_steps], device="cuda")
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    # the box's host-copy ceiling: the same H2D and D2H copies of every step, overlapped the same way, on every rank at
    # the same time, with NO sort in between -- what the PCIe / host-memory side alone allows at this N
    def copies_only(steps):
        for i in range(steps):
            kb, vb = in_bufs[i % 2]
            ko, vo = out_bufs[i % 2]
            with torch.cuda.stream(s_in):
                kb.view(torch.int32).copy_(h_keys, non_blocking=True)
                vb.view(torch.int32).copy_(h_vals, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_ok.copy_(ko.view(torch.int32), non_blocking=True)
                h_ov.copy_(vo.view(torch.int32), non_blocking=True)
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    copies_only(1)
    barrier()
    ev0.record()
    copies_only(e2e_steps)
    ev1.record()
    barrier()
    copy_ms = torch.tensor([ev0.elapsed_time(ev1) / e2e_steps], device="cuda")
    dist.all_reduce(copy_ms, op=dist.ReduceOp.MAX)
    copy_ms = float(copy_ms.item())
    # restore the shard (in_bufs[0] is keys/vals and holds it already; uploads wrote the same data)
    ok, ov = distributed_sort(keys, vals, out=out_buf)  # keys/vals still hold the shard (uploaded from h_keys/h_vals)

    # ---- verification of the last result, on the device, at every N (asserts; the line carries "verified": true).
    # Mirrors /root/reference/cudax/test/multi_gpu/algorithms/sort/sort_common.cuh:128-155 (global order of the
    # concatenated per-rank outputs) and adds what a STABLE pair sort must also satisfy.  Values are the global input
    # positions (rank * n + i), so: (1) every rank's output is ordered by key; (2) inside a run of equal keys the
    # values increase (stability: earlier rank, then earlier local position, first); (3) the same two properties hold
    # across every rank boundary; (4) pairing: the key stored at global input position vals_out[j] is keys_out[j]
    # (every source rank's shard is regenerated from its seed); (5) the multiset of values is exactly 0 .. total-1
    # (sum and sum of squares modulo 2^64, counts per rank).
    total = n * world
    k64 = ok.to(torch.int64)
    v64 = ov.to(torch.int64)
    assert ok.numel() == n and ov.numel() == n, "bench: per-rank output count differs from the input count"
    assert bool((k64[1:] >= k64[:-1]).all()), "bench: local output not sorted"
    ties = k64[1:] == k64[:-1]
    assert bool((v64[1:][ties] > v64[:-1][ties]).all()), "bench: equal keys are not in input order (stability)"
    edge = torch.stack([k64[0], v64[0], k64[-1], v64[-1]])
    edges = [torch.empty_like(edge) for _ in range(world)]
    dist.all_gather(edges, edge)
    for a, b in zip(edges[:-1], edges[1:]):
        la, fb = (int(a[2]), int(a[3])), (int(b[0]), int(b[1]))
        assert la < fb, "bench: ranks are not globally ordered / stable across a rank boundary"
    src_of = v64 // n
    for src in range(world):
        gs = torch.Generator(device="cuda").manual_seed(42 + src)
        ks = torch.randint(-(2**31), 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=gs)
        m = src_of == src
        assert bool((ks[(v64[m] - src * n)] == ok.view(torch.int32)[m]).all()), "bench: key/value pairing broken"
        del ks, m
    sums = torch.stack([v64.sum(), (v64 * v64).sum()])  # int64 arithmetic wraps modulo 2^64
    dist.all_reduce(sums)
    want1 = (total * (total - 1) // 2) % (1 << 64)
    want2 = ((total - 1) * total * (2 * total - 1) // 6) % (1 << 64)
    got1, got2 = int(sums[0]) % (1 << 64), int(sums[1]) % (1 << 64)
    assert (got1, got2) == (want1, want2), "bench: the output values are not a permutation of the input positions"
    verified = True
    del k64, v64, src_of, ties

    # ---- like-for-like anchor: the local (u32, u32) sort of one shard on this GPU, in the same run (every rank runs it
    # at the same time, rank 0 reports): efficiency_like_for_like = per-GPU throughput of the job / this
    from .multi_gpu import _default_ops as _ops

    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        _ops().sort_pairs(keys, vals, False, preserve_input=True, out=out_buf)
    a0.record()
    for _ in range(5):
        _ops().sort_pairs(keys, vals, False, preserve_input=True, out=out_buf)
    a1.record()
    torch.cuda.synchronize()
    anchor_ms = a0.elapsed_time(a1) / 5
    barrier()

    phase_ms = {k: sum(v) / len(v) for k, v in phase.items()}
    PH = ("splitters", "partition", "barrier", "exchange", "final_sort", "local_sort")
    ph = torch.tensor([phase_ms.get(k, 0.0) for k in PH], device="cuda")
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    xbytes = torch.tensor([float(st.get("exchange_bytes_out", 0))], device="cuda")
    dist.all_reduce(xbytes, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = measured_peaks()
        fused = st.get("exchange") == "fused"
        ex_ms = float(ph[1]) if fused else float(ph[3])  # fused: the partition kernel IS the exchange
        one = 2.0 * n * 8  # bytes one onesweep pass moves per GPU
        sort_ms = float(ph[4])
        line = {
            "metric": metric,
            "value": total / (ms * 1e-3) / 1e9,
            "unit": "Gkeys/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": workload_name, "keys": "uint32", "values": "uint32", "pairs_per_gpu": n,
                       "total_pairs": total, "distribution": "uniform", "parallelism": f"range-partition x{world}",
                       "protocol": st.get("protocol"),
                       "host": ("C++ (b200rs_sort_multi): kernels only, select-round all-reduces over peer-mapped memory, "
                                "no NCCL call and no host wait per sort") if native else
                               "Python over torch.distributed (NCCL all-reduces, symmetric memory)",
                       "exchange": {"fused": "fused into the partition kernel: direct stores into the destination GPUs' "
                                             "symmetric-memory receive buffers over NVLink 5 / NVSwitch",
                                    "peer": "direct peer copies into symmetric-memory receive buffers over NVLink 5 / "
                                            "NVSwitch"}.get(st.get("exchange"),
                                                            "torch.distributed all_to_all_single (NCCL over NVLink 5 / "
                                                            "NVSwitch)"),
                       "l2_policy": "inputs (2 GiB per GPU) larger than L2; no flush needed"},
            "roofline": {"bound": "hbm", "kernel": "onesweep_kernel (one 8-bit digit pass, per GPU)",
                         "achieved": 4 * one / (sort_ms * 1e-3) / 1e9 if sort_ms > 0 else None, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s",
                         "frac": (4 * one / (sort_ms * 1e-3) / 1e9 / peak) if sort_ms > 0 else None, "traffic": None,
                         "note": "achieved = 4 passes x 2 x N x 8 B / device time of the final local sort "
                                 "(histogram included in the time, not in the bytes)"},
            "phase_ms_max_over_ranks": {k: float(ph[i]) for i, k in enumerate(PH) if float(ph[i]) > 0},
            "final_sort_ops_ms_rank0": final_ops,
            "exchange": {"fused_with_partition_pass": fused, "bytes_out_per_gpu": float(xbytes), "ms": ex_ms,
                         "GBps_per_direction": float(xbytes) / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else None,
                         "nvlink_peak_GBps": 770.0, "peak_source": "measured peer copy (B200_PROFILING.md)"},
            "verified": verified,
            "verification": "on device, every rank: order, stability inside equal-key runs and across rank boundaries, "
                            "pairing keys_in[vals_out] == keys_out, values are a permutation of the input positions",
            "scale_anchor": {"workload": "local SortPairs of one shard (u32, u32), same run, rank 0",
                             "ms_per_step": anchor_ms, "value": n / (anchor_ms * 1e-3) / 1e9, "unit": "Gkeys/s"},
            "efficiency_like_for_like": (total / (ms * 1e-3) / 1e9) / (world * n / (anchor_ms * 1e-3) / 1e9),
            "cpu_baseline": None,
            "e2e": {"value": total / (e2e_ms * 1e-3) / 1e9, "unit": "Gkeys/s", "h2d_bytes_per_step": n * 8 * world,
                    "d2h_bytes_per_step": n * 8 * world, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "overlap": "H2D of step i+1 and D2H of step i on copy streams while the sort runs",
                    "copies_only_ms_per_step": copy_ms,
                    "host_copy_ceiling_GBps_per_gpu_per_direction": n * 8 / (copy_ms * 1e-3) / 1e9,
                    "note": "copies_only = the same pinned-host H2D + D2H of every step on all ranks at once with no "
                            "sort in between: the part of e2e the box's PCIe / host memory sets"},
            "gpu_launches": launches * args.steps * world,
            "gpu_launches_per_step_per_gpu": launches,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()

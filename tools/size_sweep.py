#!/usr/bin/env python
"""Size sweep of the single-GPU sort against the reference's cub::DeviceRadixSort on the same GPU.

Axes follow the reference's own benchmark (cub/benchmarks/bench/radix_sort/keys.cu:60-64, pairs.cu:90-94):
Elements 2^16 .. 2^28 in steps of 2^4, bit entropy {1.000, 0.544, 0.201}; u32 keys and (u32, u32) pairs.

    python tools/size_sweep.py [--out gpurun_out/size_sweep.jsonl]

One JSON line per cell: ours (device-resident, CUDA events, temp pre-allocated) and cub (oracle/_ref tool, context only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cccl_b200 import _native  # noqa: E402


def ours(log2n, rounds, pairs, iters):
    n = 1 << log2n
    g = torch.Generator(device="cuda").manual_seed(42)
    words = max(1, n * 4 // 8)
    raw = torch.randint(-(2**63), 2**63 - 1, (words,), dtype=torch.int64, device="cuda", generator=g)
    for _ in range(rounds - 1):
        raw &= torch.randint(-(2**63), 2**63 - 1, (words,), dtype=torch.int64, device="cuda", generator=g)
    keys = raw.view(torch.uint8)[: n * 4]
    vals = torch.arange(n, dtype=torch.int32, device="cuda").view(torch.uint8) if pairs else None
    keys_out = torch.empty_like(keys)
    vals_out = torch.empty_like(vals) if pairs else None
    p = lambda t: t.data_ptr() if t is not None else 0
    stream = torch.cuda.current_stream().cuda_stream
    vb = 4 if pairs else 0
    need, _ = _native.sort_raw(0, 0, p(keys), p(keys_out), p(vals), p(vals_out), n, 0, 4, vb, 0, 32, False, False, stream)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")

    def step():
        _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, 0, 4, vb, 0, 32, False,
                         False, stream)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # the same loop replayed from one captured CUDA graph (what a latency-sensitive caller would do)
    gstream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(gstream):
        _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, 0, 4, vb, 0, 32, False,
                         False, gstream.cuda_stream)
        gstream.synchronize()
        with torch.cuda.graph(graph, stream=gstream):
            _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, 0, 4, vb, 0, 32,
                             False, False, gstream.cuda_stream)
        for _ in range(3):
            graph.replay()
        gstream.synchronize()
        e0.record(gstream)
        for _ in range(iters):
            graph.replay()
        e1.record(gstream)
        gstream.synchronize()
    gms = e0.elapsed_time(e1) / iters
    return ms, gms, _native.lib().b200rs_last_launch_count()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "size_sweep.jsonl"))
    args = ap.parse_args()
    _native.lib()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for pairs in (False, True):
            for log2n in (16, 20, 24, 28):
                iters = {16: 300, 20: 200, 24: 50, 28: 10}[log2n]
                for rounds, ent in ((1, "1.000"), (3, "0.544"), (5, "0.201")):
                    ms, gms, launches = ours(log2n, rounds, pairs, iters)
                    c = bench.cub_same_gpu("uint32", 4 if pairs else 0, log2n, f"entropy{rounds}", False, 0, 32,
                                           iters=iters)
                    n = 1 << log2n
                    rec = {"workload": ("sortpairs_u32_u32" if pairs else "sortkeys_u32"), "log2n": log2n,
                           "bit_entropy": ent, "ms": ms, "gkeys_s": n / ms / 1e6, "ms_cuda_graph": gms,
                           "launches": launches, "cub_ms": c.get("ms_per_step") if c else None,
                           "cub_gkeys_s": c.get("value") if c else None,
                           "speedup_vs_cub": (c["ms_per_step"] / ms) if c and "ms_per_step" in c else None}
                    print(json.dumps(rec), flush=True)
                    f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Times every compiled onesweep tile configuration on the BASELINE workloads (device-resident inputs).

    python tools/sweep.py [--workloads a,b] [--steps 5] [--out gpurun_out/sweep.json]

Prints one line per (workload, config): whole-sort ms, Gkeys/s, per-kernel ms (from the library's stream events).
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
from cccl_b200.radix_sort import key_kind_of  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="sortkeys_u32_2^28_uniform,sortpairs_u32_u32_2^28_uniform,"
                                           "sortkeys_i64_desc_2^28,sortpairs_u64_u32_2^28_uniform")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    lib = _native.lib()
    results = []
    for name in args.workloads.split(","):
        kdt, vdt, log2n, dist, desc, b, e = bench.WORKLOADS[name]
        n = 1 << log2n
        kb = np.dtype(kdt).itemsize
        vb = np.dtype(vdt).itemsize if vdt else 0
        kind = key_kind_of(np.dtype(kdt))
        keys, vals = bench.make_device_input(torch, np, name)
        keys_out = torch.empty_like(keys)
        vals_out = torch.empty_like(vals) if vals is not None else None
        p = lambda t: t.data_ptr() if t is not None else 0
        stream = torch.cuda.current_stream().cuda_stream
        for ci, desc_str in enumerate(_native.describe_configs(kb, vb)):
            lib.b200rs_set_config(ci)
            need, _ = _native.sort_raw(0, 0, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e, desc,
                                       False, stream)
            temp = torch.empty(need, dtype=torch.uint8, device="cuda")

            def step():
                _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb,
                                 b, e, desc, False, stream)

            try:
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
            except Exception as ex:  # e.g. launch failure for an over-sized configuration
                print(f"{name} cfg{ci} {desc_str}: FAILED {ex}")
                continue
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(args.steps):
                step()
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / args.steps
            lib.b200rs_timing_enable(1)
            step()
            ops = _native.timing_read()
            lib.b200rs_timing_enable(0)
            one = [t for o, t in ops if o == "onesweep"]
            hist = [t for o, t in ops if o == "histogram"]
            rec = {"workload": name, "config": ci, "desc": desc_str, "ms": ms, "gkeys_s": n / ms / 1e6,
                   "onesweep_ms": one, "hist_ms": hist,
                   "onesweep_gbs": 2.0 * n * (kb + vb) / (sum(one) / len(one)) / 1e6}
            results.append(rec)
            print(f"{name} cfg{ci:2d} {desc_str}: {ms:8.3f} ms {rec['gkeys_s']:7.2f} Gkeys/s | hist {hist[0]:.3f} ms | "
                  f"onesweep avg {sum(one) / len(one):.3f} ms = {rec['onesweep_gbs']:.0f} GB/s", flush=True)
            del temp
        del keys, vals, keys_out, vals_out
        torch.cuda.empty_cache()
    lib.b200rs_set_config(-1)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

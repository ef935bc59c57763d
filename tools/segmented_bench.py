#!/usr/bin/env python
"""Segmented sort against the reference's cub::DeviceSegmentedRadixSort on the same GPU (context numbers).

    python tools/segmented_bench.py [--out gpurun_out/segmented_bench.jsonl]

Shapes: 2^24 (u32, u32) pairs cut into contiguous segments of random length with mean 100 / 1 000 / 20 000 / 2^20
(the reference's own benchmark sweeps the segment size the same way, cub/benchmarks/bench/segmented_sort/keys.cu)."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from cccl_b200 import _native  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_cub_radix_sort")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "segmented_bench.jsonl"))
    ap.add_argument("--log2n", type=int, default=24)
    ap.add_argument("--long-min", type=int, nargs="*", default=None,
                    help="sweep b200rs_set_segmented_long_min over these values (0 = one CTA per segment always)")
    ap.add_argument("--means", type=int, nargs="*", default=[100, 1000, 20000, 1 << 20])
    args = ap.parse_args()
    lib = _native.lib()
    n = 1 << args.log2n
    rng = np.random.default_rng(7)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f, tempfile.TemporaryDirectory() as d:
        for mean in args.means:
            lens = rng.integers(0, 2 * mean + 1, size=max(2, 2 * n // mean)).astype(np.int64)
            ends = np.cumsum(lens)
            ends = ends[ends <= n]
            begins = np.concatenate([[0], ends[:-1]]).astype(np.int64)
            segs = len(begins)
            dk, dv = torch.from_numpy(keys.view(np.int32)).cuda(), torch.from_numpy(vals.view(np.int32)).cuda()
            ko, vo = torch.empty_like(dk), torch.empty_like(dv)
            db, de = torch.from_numpy(begins).cuda(), torch.from_numpy(ends.astype(np.int64)).cuda()
            st = torch.cuda.current_stream().cuda_stream
            a = (n, segs, db.data_ptr(), de.data_ptr(), 8, 0, 4, 4, 0, 32, 0, st)
            if args.long_min:  # threshold sweep: our side only
                for lm in args.long_min:
                    lib.b200rs_set_segmented_long_min(lm)
                    nb = ctypes.c_size_t(0)
                    _native.check(lib.b200rs_segmented_sort(None, ctypes.byref(nb), None, None, None, None, *a), "query")
                    temp = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
                    run = lambda: _native.check(lib.b200rs_segmented_sort(
                        temp.data_ptr(), ctypes.byref(nb), dk.data_ptr(), ko.data_ptr(), dv.data_ptr(), vo.data_ptr(), *a), "sort")
                    for _ in range(3):
                        run()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(20):
                        run()
                    e1.record()
                    torch.cuda.synchronize()
                    rec = {"mean_segment": mean, "long_min": lm, "ms": e0.elapsed_time(e1) / 20, "temp_bytes": nb.value}
                    print(json.dumps(rec), flush=True)
                    f.write(json.dumps(rec) + "\n")
                lib.b200rs_set_segmented_long_min(1 << 16)
                continue
            nb = ctypes.c_size_t(0)
            _native.check(lib.b200rs_segmented_sort(None, ctypes.byref(nb), None, None, None, None, *a), "query")
            temp = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
            run = lambda: _native.check(lib.b200rs_segmented_sort(temp.data_ptr(), ctypes.byref(nb), dk.data_ptr(),
                                                                  ko.data_ptr(), dv.data_ptr(), vo.data_ptr(), *a), "sort")
            iters = 20
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            rec = {"workload": "segmented_sortpairs_u32_u32", "n": n, "segments": segs, "mean_segment": mean, "ms": ms,
                   "gkeys_s": n / ms / 1e6, "temp_bytes": nb.value}
            if os.path.exists(REF):
                kf, vf, bf, ef = (os.path.join(d, x) for x in ("k", "v", "b", "e"))
                keys.tofile(kf), vals.tofile(vf), begins.tofile(bf), ends.astype(np.int64).tofile(ef)
                env = dict(os.environ, REF_SEG_ITERS=str(iters))
                out = subprocess.run([REF, "segsort", "u32", "4", str(n), "0", "0", "32", kf, vf, os.path.join(d, "ko"),
                                      os.path.join(d, "vo"), str(segs), bf, ef], capture_output=True, text=True, env=env,
                                     timeout=600)
                for line in out.stdout.splitlines():
                    if line.startswith("{"):
                        c = json.loads(line)
                        rec["cub_ms"], rec["cub_temp_bytes"] = c["ms"], c["temp_bytes"]
                        rec["speedup_vs_cub"] = c["ms"] / ms
                # and the results agree bit for bit on the way
                rk = np.fromfile(os.path.join(d, "ko"), dtype=np.uint32)
                rec["matches_cub"] = bool(np.array_equal(rk[: int(ends[-1])], ko.cpu().numpy().view(np.uint32)[: int(ends[-1])]))
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()

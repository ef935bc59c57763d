#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the LSD radix-sort hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one complete device-wide sort of one batch of synthetic keys (histogram upsweep + every digit pass).

N = 1   workload = BASELINE.json configs[1]: cub::DeviceRadixSort::SortKeys of 2^28 uniform u32 keys (pointer API,
        temp storage pre-allocated).  `value` = keys sorted per second with the input already in HBM; `e2e` = the
        same sort through the public API with HOST (pinned) buffers: H2D of the keys, sort, D2H of the sorted
        keys, all inside the timed region.
N > 1   workload = configs[4]: distributed SortPairs of 2^28 (u32 key, u32 value) pairs PER GPU (weak scaling),
        one process per GPU under torchrun, exchange over NCCL/NVLink; `value` = all ranks' pairs / max-over-ranks
        device time.
--impl reference   the reference's own CPU implementation of the path (thrust::sort, OMP backend, compiled from the
        unmodified reference into oracle/_ref) on the box's host cores; each step sorts a bounded sample
        (2^24 keys = configs[0]).  Rank 0 only.

One JSON line on stdout (rank 0).  Inputs are larger than L2 (1 GiB vs 126 MB), so no flush is needed between
timed iterations; every iteration re-sorts the same unsorted input buffer.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (key dtype, value dtype or None, log2 n, distribution, descending, begin_bit, end_bit)
    "sortkeys_u32_2^28_uniform": ("uint32", None, 28, "uniform", False, 0, 32),
    "sortpairs_u64_u32_2^28_uniform": ("uint64", "uint32", 28, "uniform", False, 0, 64),
    "sortpairs_u64_u32_2^28_entropy0.201": ("uint64", "uint32", 28, "entropy5", False, 0, 64),
    "sortkeys_f32_desc_2^28_bits8_24": ("float32", None, 28, "uniform", True, 8, 24),
    "sortkeys_f32_desc_2^28": ("float32", None, 28, "uniform", True, 0, 32),
    "sortkeys_i64_desc_2^28_bits16_48": ("int64", None, 28, "uniform", True, 16, 48),
    "sortkeys_i64_desc_2^28": ("int64", None, 28, "uniform", True, 0, 64),
    "sortpairs_u32_u32_2^28_uniform": ("uint32", "uint32", 28, "uniform", False, 0, 32),
    # skew rows of SURVEY.md 10.16 (reported for "no pathological slowdown", not headline numbers)
    "sortkeys_u32_2^28_entropy0.201": ("uint32", None, 28, "entropy5", False, 0, 32),
    "sortkeys_u32_2^28_equal": ("uint32", None, 28, "equal", False, 0, 32),
    "sortkeys_u32_2^28_few16": ("uint32", None, 28, "few16", False, 0, 32),
    "sortkeys_u32_2^28_sorted": ("uint32", None, 28, "sorted", False, 0, 32),
}
DEFAULT_1GPU = "sortkeys_u32_2^28_uniform"
DEFAULT_NGPU = "dist_sortpairs_u32_u32_2^28_per_gpu"
METRIC = "radix sort Gkeys/s (u32 keys, u64/u32 pairs) at 1/2/4/8 B200; % HBM roofline"


def algorithmic_bytes_per_item(kb, vb, bits):
    """B = N*k + ceil(bits/8) * 2 * N * (k+v)."""
    passes = (bits + 7) // 8
    return kb + passes * 2 * (kb + vb)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.05):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.period = period
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_run(steps, warmup, log2n=24, threads=None, pairs=False):
    """thrust::sort / sort_by_key (OMP) of the unmodified reference on a bounded sample: 2^24 uniform u32 keys
    (+ u32 values when `pairs`, the multi-GPU arm's workload) per step."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gen import make_keys
    from oracle_lib import ref_thrust, ref_thrust_sort, oracle_sort

    n = 1 << log2n
    keys = make_keys("uniform", n, np.uint32, seed=42)
    vals = np.arange(n, dtype=np.uint32) if pairs else None
    lib = ref_thrust("omp")
    kind = "reference"
    if lib is None:
        kind = "port"
    if lib is not None:
        # all the host cores this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently
        # turn the N > 1 reference arm into a one-core run
        if not threads:
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        lib.ref_thrust_set_threads(threads)
        cores = threads
        times = []
        for i in range(warmup + steps):
            secs = ref_thrust_sort(keys, vals, backend="omp")[-1]
            if i >= warmup:
                times.append(secs)
    else:
        cores = 1
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle_sort(keys, vals) if pairs else oracle_sort(keys)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return {
        "value": n * len(times) / total / 1e9,
        "unit": "Gkeys/s",
        "cores": cores,
        "kind": kind,
        "sample": f"thrust::{'sort_by_key' if pairs else 'sort'} (OMP backend) of 2^{log2n} uniform u32 keys"
                  f"{' + u32 values' if pairs else ''} per step, {len(times)} steps, sort call only (host copies excluded)",
        "ms_per_step": total / len(times) * 1e3,
    }


def cub_same_gpu(kdt, vb, log2n, dist, desc, b, e, iters=10):
    """Context number: the UNMODIFIED reference cub::DeviceRadixSort (oracle/_ref/ref_cub_radix_sort, built from the
    reference headers for sm_100a) on the same GPU and input shape.  Separate process; reported, never on our path."""
    import subprocess

    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cub_radix_sort")
    kt = {"uint32": "u32", "uint64": "u64", "float32": "f32", "int64": "i64"}[kdt]
    if not os.path.exists(exe) or vb not in (0, 4, 8):
        return None
    # the tool's own generators: AND rounds for the entropy rows, equal / fewK / sorted by name (uniform otherwise)
    rounds = dist[7:] if dist.startswith("entropy") else (dist if dist == "equal" or dist == "sorted"
                                                          or dist.startswith("few") else "1")
    try:
        out = subprocess.run([exe, "bench", kt, str(vb), str(log2n), str(int(desc)), str(b), str(e), str(rounds),
                              str(iters)], capture_output=True, text=True, timeout=300).stdout
        r = json.loads(out.strip().splitlines()[-1])
        return {"value": r["gkeys_s"], "unit": "Gkeys/s", "ms_per_step": r["ms"], "impl": r["impl"],
                "temp_bytes": r["temp_bytes"]}
    except Exception as ex:  # context only
        return {"unavailable": str(ex)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    r = cpu_reference_run(steps, min(args.warmup, 1), pairs=args.gpus > 1)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": r["value"],
        "unit": "Gkeys/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": min(args.warmup, 1),
        "ms_per_step": r["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": DEFAULT_1GPU if args.gpus == 1 else DEFAULT_NGPU,
                   "reference_arm": ("thrust::sort_by_key" if args.gpus > 1 else "thrust::sort")
                   + ", THRUST_DEVICE_SYSTEM=OMP, host cores, bounded sample 2^24 items/step"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "Gkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def make_device_input(torch, np, name):
    kdt, vdt, log2n, dist, desc, b, e = WORKLOADS[name]
    n = 1 << log2n
    g = torch.Generator(device="cuda").manual_seed(42)
    kb = np.dtype(kdt).itemsize
    words = n * kb // 8
    raw = torch.randint(-(2**63), 2**63 - 1, (words,), dtype=torch.int64, device="cuda", generator=g)
    if dist.startswith("entropy"):
        for _ in range(int(dist[7:]) - 1):
            raw &= torch.randint(-(2**63), 2**63 - 1, (words,), dtype=torch.int64, device="cuda", generator=g)
    if dist == "equal":
        raw.fill_(0x0123456701234567)
    elif dist.startswith("few"):
        k = int(dist[3:])
        pool = torch.randint(-(2**63), 2**63 - 1, (k,), dtype=torch.int64, device="cuda", generator=g)
        idx = torch.randint(0, k, (n,), dtype=torch.int64, device="cuda", generator=g)
        raw = pool.view(torch.int32 if kb == 4 else torch.int64)[:k][idx].contiguous().view(torch.int64)
    elif dist == "sorted":
        raw = torch.sort(raw.view(torch.int32 if kb == 4 else torch.int64)).values.view(torch.int64)
    keys = raw.view(torch.uint8)
    vals = None
    if vdt is not None:
        vals = torch.arange(n, dtype=torch.int32, device="cuda").view(torch.uint8)
    return keys, vals


def device_leg(name, steps, warmup, with_cub=True):
    """Device-resident timing of one named workload (inputs in HBM, temp pre-allocated, CUDA events): the compact
    record that the default bench line carries for every BASELINE config besides the headline one."""
    import numpy as np
    import torch

    from cccl_b200 import _native
    from cccl_b200.radix_sort import key_kind_of

    kdt, vdt, log2n, dist, desc, b, e = WORKLOADS[name]
    n = 1 << log2n
    kb = np.dtype(kdt).itemsize
    vb = np.dtype(vdt).itemsize if vdt else 0
    kind = key_kind_of(np.dtype(kdt))
    lib = _native.lib()
    keys, vals = make_device_input(torch, np, name)
    keys_out = torch.empty_like(keys)
    vals_out = torch.empty_like(vals) if vals is not None else None
    p = lambda t: t.data_ptr() if t is not None else 0
    stream = torch.cuda.current_stream().cuda_stream
    need, _ = _native.sort_raw(0, 0, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e, desc, False,
                               stream)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")

    def step():
        _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e,
                         desc, False, stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    lib.b200rs_timing_enable(1)
    step()
    ops = _native.timing_read()
    lib.b200rs_timing_enable(0)
    one = [t for o, t in ops if o == "onesweep"]
    peak, _ = measured_peaks()
    whole = algorithmic_bytes_per_item(kb, vb, e - b) * n
    rec = {"workload": name, "value": n / (ms * 1e-3) / 1e9, "unit": "Gkeys/s", "ms_per_step": ms,
           "onesweep_ms": sum(one) / len(one), "onesweep_frac": 2.0 * n * (kb + vb) / (sum(one) / len(one) * 1e-3) / 1e9 / peak,
           "whole_sort_frac": whole / (ms * 1e-3) / 1e9 / peak, "steps": steps,
           "tile_config": _native.describe_configs(kb, vb)[0]}
    del keys, vals, keys_out, vals_out, temp
    torch.cuda.empty_cache()
    if with_cub:
        c = cub_same_gpu(kdt, vb, log2n, dist, desc, b, e, iters=5)
        rec["cub_same_gpu"] = c.get("value") if c and "value" in c else None
    return rec


# every other BASELINE.json config (C3 uniform / entropy 0.201, C4 f32 and i64 descending on a bit window and on all
# bits), the like-for-like anchor of the multi-GPU arm, and the skew rows of SURVEY.md 10.16
EXTRA_WORKLOADS = [
    "sortpairs_u64_u32_2^28_uniform", "sortpairs_u64_u32_2^28_entropy0.201",
    "sortkeys_f32_desc_2^28_bits8_24", "sortkeys_f32_desc_2^28",
    "sortkeys_i64_desc_2^28_bits16_48", "sortkeys_i64_desc_2^28",
    "sortpairs_u32_u32_2^28_uniform",
    "sortkeys_u32_2^28_entropy0.201", "sortkeys_u32_2^28_equal", "sortkeys_u32_2^28_few16", "sortkeys_u32_2^28_sorted",
]


def run_single_gpu(args):
    import numpy as np
    import torch

    from cccl_b200 import _native
    from cccl_b200.radix_sort import key_kind_of

    name = args.workload or DEFAULT_1GPU
    kdt, vdt, log2n, dist, desc, b, e = WORKLOADS[name]
    n = 1 << log2n
    kb = np.dtype(kdt).itemsize
    vb = np.dtype(vdt).itemsize if vdt else 0
    kind = key_kind_of(np.dtype(kdt))
    torch.cuda.set_device(0)
    lib = _native.lib()  # raises if the CUDA library is missing: there is no fallback
    if args.config is not None:
        lib.b200rs_set_config(args.config)

    keys, vals = make_device_input(torch, np, name)
    keys_out = torch.empty_like(keys)
    vals_out = torch.empty_like(vals) if vals is not None else None
    p = lambda t: t.data_ptr() if t is not None else 0
    stream = torch.cuda.current_stream().cuda_stream
    need, _ = _native.sort_raw(0, 0, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e, desc, False,
                               stream)
    temp = torch.empty(need, dtype=torch.uint8, device="cuda")

    def step():
        _native.sort_raw(temp.data_ptr(), need, p(keys), p(keys_out), p(vals), p(vals_out), n, kind, kb, vb, b, e,
                         desc, False, stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = lib.b200rs_last_launch_count()

    # ---- timed region: device-resident inputs
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clocks:
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    value = n / (ms * 1e-3) / 1e9

    # ---- per-kernel timing of the same steps (events recorded inside the library on the launching stream)
    lib.b200rs_timing_enable(1)
    per_op = {}
    for _ in range(args.steps):
        step()
        for opname, t in _native.timing_read():
            per_op.setdefault(opname, []).append(t)
    lib.b200rs_timing_enable(0)
    passes = (e - b + 7) // 8
    one_ms = sum(per_op["onesweep"]) / len(per_op["onesweep"])
    hist_ms = sum(per_op["histogram"]) / len(per_op["histogram"])
    step_ms_timed = sum(sum(v) for v in per_op.values()) / args.steps
    peak, peak_src = measured_peaks()
    bytes_per_launch = 2.0 * n * (kb + vb)
    achieved = bytes_per_launch / (one_ms * 1e-3) / 1e9
    whole_bytes = algorithmic_bytes_per_item(kb, vb, e - b) * n
    roofline = {
        "bound": "hbm",
        "kernel": "onesweep_kernel (one 8-bit digit pass)",
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "bytes_per_launch": bytes_per_launch,
        "launch_ms": one_ms,
        "launches_per_step": passes,
        "kernel_share_of_step": one_ms * passes / step_ms_timed,
        "histogram_ms": hist_ms,
        "histogram_gbs": n * kb / (hist_ms * 1e-3) / 1e9,
        "whole_sort": {"algorithmic_bytes": whole_bytes, "achieved_gbs": whole_bytes / (ms * 1e-3) / 1e9,
                       "frac": whole_bytes / (ms * 1e-3) / 1e9 / peak,
                       "frac_of_nominal_8TBs": whole_bytes / (ms * 1e-3) / 1e9 / 8000.0},
    }
    prof = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(prof)) if os.path.isdir(prof) else []:
        if f.endswith("traffic.json"):
            try:
                t = json.load(open(os.path.join(prof, f)))
                if t.get("workload") == name:
                    roofline["traffic"] = t.get("dram_bytes_per_launch")
                    roofline["traffic_source"] = "profiles/" + f
            except Exception:
                pass

    # ---- end to end through the public API with HOST buffers (pinned): H2D + sort + D2H every step
    from cccl_b200 import SortOrder, make_radix_sort

    tdt = {"uint32": torch.uint32, "uint64": torch.uint64, "float32": torch.float32, "int64": torch.int64}[kdt]
    h_in = torch.empty(n * kb, dtype=torch.uint8).pin_memory()
    h_in.copy_(keys.cpu())
    h_out = torch.empty(n * kb, dtype=torch.uint8).pin_memory()
    h_vin = h_vout = None
    if vals is not None:
        h_vin = torch.empty(n * vb, dtype=torch.uint8).pin_memory()
        h_vin.copy_(vals.cpu())
        h_vout = torch.empty(n * vb, dtype=torch.uint8).pin_memory()
    d_in, d_out = keys.view(tdt), keys_out.view(tdt)
    d_vin = vals.view(torch.int32) if vals is not None else None
    d_vout = vals_out.view(torch.int32) if vals is not None else None
    order = SortOrder.DESCENDING if desc else SortOrder.ASCENDING
    sorter = make_radix_sort(d_in_keys=d_in, d_out_keys=d_out, d_in_values=d_vin, d_out_values=d_vout, order=order)
    kw = dict(d_in_keys=d_in, d_out_keys=d_out, d_in_values=d_vin, d_out_values=d_vout, num_items=n, begin_bit=b,
              end_bit=e)

    # Several steps in flight: each step is a chain H2D -> sort -> D2H on its own stream with its own device buffers and
    # temp storage, so step i's D2H overlaps step i+1's H2D (PCIe is full duplex).  Every step still copies its inputs
    # from pinned host memory and reads its result back inside the timed region.
    DEPTH = 3
    slots = [dict(keys=keys, keys_out=keys_out, vals=vals, vals_out=vals_out, temp=temp)]
    for _ in range(DEPTH - 1):
        slots.append(dict(keys=torch.empty_like(keys), keys_out=torch.empty_like(keys_out),
                          vals=torch.empty_like(vals) if vals is not None else None,
                          vals_out=torch.empty_like(vals_out) if vals is not None else None,
                          temp=torch.empty_like(temp)))
    for i, sl in enumerate(slots):
        sl["h_out"] = h_out if i == 0 else torch.empty(n * kb, dtype=torch.uint8).pin_memory()
        sl["h_vout"] = h_vout if (i == 0 or vals is None) else torch.empty(n * vb, dtype=torch.uint8).pin_memory()
        sl["stream"] = torch.cuda.Stream()
        sl["kw"] = dict(d_in_keys=sl["keys"].view(tdt), d_out_keys=sl["keys_out"].view(tdt),
                        d_in_values=sl["vals"].view(torch.int32) if vals is not None else None,
                        d_out_values=sl["vals_out"].view(torch.int32) if vals is not None else None,
                        num_items=n, begin_bit=b, end_bit=e)

    def e2e_step(i, depth, do_sort=True):
        sl = slots[i % depth]
        with torch.cuda.stream(sl["stream"]):
            sl["keys"].copy_(h_in, non_blocking=True)
            if vals is not None:
                sl["vals"].copy_(h_vin, non_blocking=True)
            if do_sort:
                sorter(temp_storage=sl["temp"], stream=sl["stream"], **sl["kw"])
            sl["h_out"].copy_(sl["keys_out"], non_blocking=True)
            if vals is not None:
                sl["h_vout"].copy_(sl["vals_out"], non_blocking=True)

    def e2e_run(depth, do_sort=True):
        e2e_step(0, depth, do_sort)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(slots[0]["stream"])  # every stream is idle here: this is the start of the first H2D
        for i in range(e2e_steps):
            e2e_step(i, depth, do_sort)
        cur = torch.cuda.current_stream()
        for sl in slots[:depth]:
            cur.wait_stream(sl["stream"])
        t1.record(cur)
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / e2e_steps

    e2e_steps = max(3, min(args.steps, 12))
    e2e_serial_ms = e2e_run(1)
    # the host-link ceiling: the same pinned-host H2D + D2H of every step with NO sort in between (not part of `value`)
    e2e_copies_ms = e2e_run(DEPTH, do_sort=False)
    e2e_ms = e2e_run(DEPTH)
    e2e = {"value": n / (e2e_ms * 1e-3) / 1e9, "unit": "Gkeys/s", "h2d_bytes_per_step": n * (kb + vb),
           "d2h_bytes_per_step": n * (kb + vb), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "steps_in_flight": DEPTH, "timer": "CUDA events: first H2D start -> join of all streams",
           "serial_value": n / (e2e_serial_ms * 1e-3) / 1e9, "serial_ms_per_step": e2e_serial_ms,
           "copies_only_ms_per_step": e2e_copies_ms,
           "note": "copies_only = the same H2D + D2H per step, same streams, without the sort: the host-link ceiling"}

    # spot check the last e2e result on the host (sortedness of a sample); full parity lives in tests/
    if kdt == "uint32" and b == 0 and e == 32 and not desc:
        s = slots[(e2e_steps - 1) % DEPTH]["h_out"].view(torch.int32)[: 1 << 20].numpy().view(np.uint32)
        assert (s[1:] >= s[:-1]).all(), "bench: output not sorted"

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    cub = None if args.no_cub else cub_same_gpu(kdt, vb, log2n, dist, desc, b, e)

    # the other BASELINE configs and the skew rows, device-resident, so that the driver's own run times them too
    extra = None
    if args.workload is None and not args.no_extra:
        del slots, sorter, keys, keys_out, vals, vals_out, temp, d_in, d_out, d_vin, d_vout, kw
        torch.cuda.empty_cache()
        extra = []
        for w in EXTRA_WORKLOADS:
            try:
                extra.append(device_leg(w, max(3, min(args.steps, 10)), 3, with_cub=not args.no_cub))
            except Exception as ex:  # reported, never fatal for the headline line
                extra.append({"workload": w, "error": str(ex)[:200]})
        # the adjacent caller of SURVEY 8f-4: top-k selection (b200rs_topk vs the unmodified cub::DeviceTopK on this GPU)
        try:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
            import topk_bench

            for log2k, rounds in ((11, 1), (11, 5), (23, 1)):
                o = topk_bench.ours(28, log2k, rounds, iters=10)
                c = None if args.no_cub else topk_bench.cub(28, log2k, rounds, iters=5)
                extra.append({"workload": f"topk_u32_2^28_k2^{log2k}_" + ("uniform" if rounds == 1 else "entropy0.201"),
                              "value": o["gkeys_per_s"], "unit": "Gkeys/s", "ms_per_step": o["ms"],
                              "algorithmic_bytes": (1 << 28) * 4 + (1 << log2k) * 4,
                              "frac_of_one_read": ((1 << 28) * 4 / (o["ms"] * 1e-3) / 1e9) / measured_peaks()[0],
                              "cub_same_gpu": c.get("gkeys_per_s") if c else None})
        except Exception as ex:
            extra.append({"workload": "topk_u32_2^28", "error": str(ex)[:200]})

    line = {
        "metric": METRIC,
        "value": value,
        "unit": "Gkeys/s",
        "n_gpus": 1,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"uint32": "u32", "uint64": "u64", "float32": "f32(bits)", "int64": "i64"}[kdt],
        "data": "synthetic",
        "config": {"workload": name, "keys": kdt, "values": vdt, "num_items": n, "distribution": dist,
                   "descending": desc, "begin_bit": b, "end_bit": e, "api": "pointer (is_overwrite_okay=0)",
                   "l2_policy": "inputs (>=1 GiB) larger than L2; no flush needed",
                   "tile_config": _native.describe_configs(kb, vb)[args.config or 0]},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "cub_same_gpu": cub,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "per_op_ms": {k: sum(v) / len(v) for k, v in per_op.items()},
        "clocks": clocks.summary(),
        "extra_workloads": extra,
    }
    print(json.dumps(line))


def run_multi_gpu(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        # launched as plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
        raise SystemExit(subprocess.call(cmd + sys.argv[1:]))
    from cccl_b200 import multi_gpu_bench

    multi_gpu_bench.run(args, METRIC, DEFAULT_NGPU, ClockSampler, measured_peaks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="one of: " + ", ".join(WORKLOADS))
    ap.add_argument("--config", type=int, default=None, help="force an onesweep tile configuration index")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cub", action="store_true", help="skip the cub-on-the-same-GPU context run")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads legs of the default line")
    ap.add_argument("--log2-per-gpu", type=int, default=28, help="multi-GPU: log2 of pairs per GPU")
    ap.add_argument("--multi-protocol", default="native", choices=["native", "partition", "sort"],
                    help="multi-GPU: native = b200rs_sort_multi (C++ host, kernels only); partition / sort = the Python "
                         "host over torch.distributed")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        return run_multi_gpu(args)
    return run_single_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns one GPU-box visit (gpurun_out/<tag>/, written by tools/gpu_round.sh) into the tracked files under profiles/:

  profiles/<tag>_launches.csv        every launch of `bench.py` under ncu (kernel, grid, block, ns) + per-kernel totals/shares
  profiles/<tag>_<kernel>_ncu.txt    the key raw metrics of the `ncu --set full` capture + per-phase (BAR.SYNC-delimited) table
  profiles/<tag>_traffic.json        dram bytes per launch of the dominant kernel (bench.py reads it for roofline.traffic)
  profiles/<tag>_bench.json          the bench line of the same visit (NOT taken under the profiler)

usage: python tools/profile_summary.py <tag> [workload]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "sortkeys_u32_2^28_uniform"
log2n = int(re.search(r"2\^(\d+)", workload).group(1))
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "memory_l1_wavefronts_shared_ideal",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return list(csv.reader(out.splitlines()))


def launches():
    path = os.path.join(src, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    per = collections.OrderedDict()
    with open(os.path.join(dst, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes.  kernel,grid,block,ns\n")
        for r in rows[1:]:
            name = re.sub(r"\(.*", "", r[ik])[:80]
            if "at::" in name:
                name = "torch:" + name.split("::")[-1][:40]
            ns = float(r[iv].replace(",", ""))
            f.write(f"{name},{r[ig]},{r[ib]},{ns:.0f}\n")
            per.setdefault(name, []).append(ns)
        ours = {k: v for k, v in per.items() if not k.startswith("torch:")}
        tot = sum(sum(v) for v in ours.values()) or 1.0
        f.write("# per-kernel totals over our kernels (share of the sort's device time):\n")
        for k, v in ours.items():
            f.write(f"# {k}: launches={len(v)} avg_ns={sum(v)/len(v):.0f} share={100*sum(v)/tot:.1f}%\n")


def full(kernel):
    rep = os.path.join(src, f"{kernel}_full.ncu-rep")
    if not os.path.exists(rep):
        return None
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    lines = [f"ncu --set full --clock-control none --import-source on, one launch of {m.get('Kernel Name', ('?',))[0][:100]}",
             f"workload {workload}; source: gpurun_out/{tag}/{kernel}_full.ncu-rep (not tracked)", ""]
    for k in KEEP:
        if k in m:
            lines.append(f"{k:90s} {m[k][0]:>18s} {m[k][1]}")
    wk = (1 << log2n) / 32.0
    try:
        inst = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
        shw = float(m["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0].replace(",", ""))
        lines += ["", f"per 32 keys: warp instructions {inst/wk:.1f}, shared-memory wavefronts {shw/wk:.2f}"]
    except Exception:
        pass
    srcp = os.path.join("/tmp", f"{tag}_{kernel}_src.csv")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    open(srcp, "w").write(out)
    ph = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), srcp, str(log2n)],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    lines += ["", "per phase (delimited by BAR.SYNC), inst/key = warp instructions per 32 keys:", ph]
    open(os.path.join(dst, f"{tag}_{kernel}_ncu.txt"), "w").write("\n".join(lines))
    rd = float(m["dram__bytes_read.sum"][0].replace(",", ""))
    wr = float(m["dram__bytes_write.sum"][0].replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    return rd * scale[m["dram__bytes_read.sum"][1]] + wr * scale[m["dram__bytes_write.sum"][1]]


launches()
traffic = full("onesweep")
full("histogram")
if traffic is not None:
    json.dump({"workload": workload, "kernel": "onesweep_kernel", "dram_bytes_per_launch": traffic,
               "source": f"ncu --set full, gpurun_out/{tag}/onesweep_full.ncu-rep"},
              open(os.path.join(dst, f"{tag}_traffic.json"), "w"))
b = os.path.join(src, "bench.json")
if os.path.exists(b) and os.path.getsize(b) > 0:
    open(os.path.join(dst, f"{tag}_bench.json"), "w").write(open(b).read())
print(sorted(f for f in os.listdir(dst) if f.startswith(tag)))

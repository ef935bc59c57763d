"""Per-phase table (phases delimited by BAR.SYNC) of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv`:
    python tools/ncu_phase_table.py src.csv
Prints warp instructions and shared-memory wavefronts per 32 keys (2^28 keys per launch assumed), the opcode mix, the
wavefronts per memory opcode and the stall-sample shares of every phase."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h:i for i,h in enumerate(hdr)}
NW = 2**28/32
# phases delimited by BAR.SYNC
phase=0; ph = collections.defaultdict(lambda: collections.Counter())
tot_inst=0
for r in data:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op0 = op.split('.')[0]
    ie = float(r[ix["Instructions Executed"]]); 
    wf = float(r[ix["L1 Wavefronts Shared"]]); wfi = float(r[ix["L1 Wavefronts Shared Ideal"]])
    smp = float(r[ix["# Samples"]])
    ph[phase]["inst"] += ie; ph[phase]["wf"] += wf; ph[phase]["wfi"] += wfi; ph[phase]["samples"] += smp
    ph[phase]["op_"+op0] += ie
    if wf>0: ph[phase]["wfop_"+op] += wf
    for k in ("stall_long_sb","stall_short_sb","stall_wait","stall_math","stall_mio","stall_barrier","stall_not_selected","stall_selected","stall_lg","stall_no_inst","stall_dispatch","stall_branch_resolving"):
        ph[phase]["S_"+k] += float(r[ix[k]])
    if op0 == "BAR": phase += 1
ts = sum(p["samples"] for p in ph.values())
for p in sorted(ph):
    c = ph[p]
    print(f"phase {p}: inst/32keys {c['inst']/NW:6.2f}  smem wf {c['wf']/NW:5.2f} (ideal {c['wfi']/NW:5.2f})  samples {100*c['samples']/ts:5.1f}%")
    ops = sorted(((v/NW,k[3:]) for k,v in c.items() if k.startswith('op_')), reverse=True)[:12]
    print("    ", " ".join(f"{k}={v:.2f}" for v,k in ops))
    w = sorted(((v/NW,k[5:]) for k,v in c.items() if k.startswith('wfop_')), reverse=True)[:8]
    print("     wf:", " ".join(f"{k}={v:.2f}" for v,k in w))
    s = sorted(((v,k[8:]) for k,v in c.items() if k.startswith('S_')), reverse=True)[:6]
    print("     stalls:", " ".join(f"{k}={100*v/ts:.1f}%" for v,k in s))

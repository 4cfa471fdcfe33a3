"""Compact view of an `ncu --page details --csv` export: the key numbers per profiled launch.
usage: python tools/ncu_details.py details.csv [launch id]"""
import csv
import sys

KEYS = ["Duration", "Elapsed Cycles", "SM Active Cycles", "Executed Instructions", "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block", "Block Size", "Grid Size", "Cluster Size", "Achieved Occupancy", "Theoretical Occupancy",
        "Warp Cycles Per Issued Instruction", "Memory Throughput", "DRAM Throughput", "L2 Cache Throughput", "L1/TEX Cache Throughput",
        "Compute (SM) Throughput", "L2 Hit Rate", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler", "No Eligible", "Avg. Active Threads Per Warp"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
want = sys.argv[2] if len(sys.argv) > 2 else None
by = {}
for r in rows[1:]:
    if len(r) < 15:
        continue
    r = r + [""] * (len(hdr) - len(r))
    by.setdefault((r[ix["ID"]], r[ix["Kernel Name"]][:60]), []).append(r)
for (lid, name), rs in by.items():
    if want is not None and lid != want:
        continue
    print(f"--- launch {lid}: {name}")
    for r in rs:
        if r[ix["Metric Name"]] in KEYS:
            print(f"   {r[ix['Metric Name']]:40s} {r[ix['Metric Value']]:>14s} {r[ix['Metric Unit']]}")
    for r in rs:
        if r[ix["Rule Name"]] and r[ix["Estimated Speedup"]]:
            try:
                if float(r[ix["Estimated Speedup"]]) >= 20:
                    print(f"   [rule {r[ix['Rule Name']]}: est. speedup {r[ix['Estimated Speedup']]} %] {r[ix['Rule Description']][:300]}")
            except ValueError:
                pass

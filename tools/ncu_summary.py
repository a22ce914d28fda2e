#!/usr/bin/env python
"""Key launch metrics + stall-reason shares of the first kernel in an ncu report.  usage: ncu_summary.py <rep>"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = raw[0], raw[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit', 'smsp__issue_active.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct', 'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'smsp__sass_inst_executed_op_local',
        'launch__waves_per_multiprocessor', 'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_op_shared']
for i, h in enumerate(hdr):
    if any(w in h for w in want) and 'per_second' not in h and 'pct_of_peak_sustained_elapsed' not in h:
        print(f"{h} [{units[i]}] = {', '.join(r[i] for r in raw[2:])}")
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]
agg = collections.Counter()
for r in src[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) < len(h2) - 2:
        continue
    for i, h in enumerate(h2):
        if h.startswith('stall_') and 'Not Issued' not in h:
            agg[h] += int(r[i] or 0)
tot = sum(agg.values()) or 1
print("stall reasons (all samples): " + ", ".join(f"{k[6:]} {v/tot*100:.1f}%" for k, v in agg.most_common(8)))

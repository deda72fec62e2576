#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a short text report.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:100])
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
            "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:80s} {vals[i]:>16s} {units[i]}")
    print("  -- warp stall cycles per issued instruction:")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, nme in sorted(st, reverse=True)[:8]:
        print(f"     {nme:28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for hi, r in enumerate(rows):
    if "Source" in r and "Address" in r:
        hdr = r
        data = rows[hi + 1:]
        break
i_s = hdr.index("Warp Stall Sampling (All Samples)")
i_src = hdr.index("Source")
i_ex = hdr.index("Instructions Executed")
data = [r for r in data if len(r) > i_s and r[i_s].isdigit()]
tot = sum(int(r[i_s]) for r in data)
print(f"== source page: {len(data)} SASS instructions, {tot} stall samples; top {topn}:")
for idx, r in sorted(sorted(enumerate(data), key=lambda t: -int(t[1][i_s]))[:topn]):
    print(f"  #{idx:5d} samples {int(r[i_s]):7d} ({100.0*int(r[i_s])/tot:5.1f}%) exec {r[i_ex]:>9s}  {r[i_src].strip()[:80]}")

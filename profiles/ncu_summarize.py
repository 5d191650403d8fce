#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box):  python profiles/ncu_summarize.py <rep> [n_lines]
Prints the headline metrics of the first captured launch, the stall mix, and the source lines
(needs -lineinfo) with the most stall samples."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]
for w in want:
    if w in hdr:
        print("%-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
for i, h in enumerate(hdr):
    if "pipe_tensor" in h and h not in want:
        print("%-70s %s %s" % (h, r[i], units[i]))
print("--- stall mix (warps stalled per issue-active cycle)")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        try:
            v = float(r[i])
        except ValueError:
            continue
        if v > 0.02:
            print("  %-40s %.3f" % (h.split("stalled_")[1].split("_per_issue")[0], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
if not his:
    sys.exit(0)
hdr = rows[his[0]]
ix = {h: i for i, h in enumerate(hdr)}
end = his[1] - 1 if len(his) > 1 else len(rows)
data = [x for x in rows[his[0] + 1:end] if len(x) == len(hdr)]
if "# Samples" in ix:
    tot = sum(int(x[ix["# Samples"]] or 0) for x in data)
    print("--- top SASS by stall samples (total %d)" % tot)
    for x in sorted(data, key=lambda x: -int(x[ix["# Samples"]] or 0))[:top]:
        print("  %6.2f%%  %s" % (100.0 * int(x[ix["# Samples"]]) / max(tot, 1), x[ix["Source"]][:90]))

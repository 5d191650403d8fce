#!/usr/bin/env python3
"""Reads an ncu report (`ncu -i X.ncu-rep --page raw --csv`) here on the CPU box and writes
  * profiles/<tag>_ncu_summary.txt : the metrics the roofline discussion uses, one block per
    kernel (the LAST captured launch of each kernel name: the first is the warm-up), and
  * profiles/traffic.json          : dram__bytes_read.sum + dram__bytes_write.sum per launch,
    which bench.py reports as `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/r2a_prof.ncu-rep r2a
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    name_i = col["Kernel Name"]
    last = {}
    for r in body:
        last[r[name_i]] = r         # the last launch of each kernel
    out = ["ncu --set full capture %s (last launch of each kernel; clock control none)" % os.path.basename(rep)]
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}

    def num(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return None

    def to_bytes(v, unit):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        return v * mult.get(unit, 1)

    for name, r in last.items():
        short = re.sub(r"^void\s+", "", name).replace("srcnn::", "")
        short = short[:short.index("(")] if "(" in short.replace("(bool)", "") else short
        short = re.sub(r"\(bool\)([01])", r"\1", short)
        depth, cut = 0, len(short)      # up to the end of the (nested) template argument list
        for i, ch in enumerate(short):
            depth += ch == "<"
            depth -= ch == ">"
            if ch == "(" and depth == 0:
                cut = i
                break
        short = short[:cut]
        out.append("")
        out.append("== %s" % short)
        for k in KEEP:
            if k in col:
                out.append("  %-86s %s %s" % (k, r[col[k]], units[col[k]]))
        rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
        if rd is not None and wr is not None:
            rb = to_bytes(rd, units[col["dram__bytes_read.sum"]])
            wb = to_bytes(wr, units[col["dram__bytes_write.sum"]])
            key = re.sub(r"\b\w+::", "", short)      # drop every namespace qualifier
            key = re.sub(r"<\(bool\)([01])>", r"<\1>", key)
            traffic[key] = {"dram_bytes": rb + wb, "dram_read": rb, "dram_write": wb,
                            "source": "profiles/%s_ncu_summary.txt" % tag}
            out.append("  => dram traffic per launch: %.1f MB read + %.1f MB written" % (rb / 1e6, wb / 1e6))
    open(os.path.join(ROOT, "profiles", "%s_ncu_summary.txt" % tag), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns an .ncu-rep (brought back under gpurun_out/) into the small text/JSON summaries that are
committed under profiles/.  Usage: python profiles/summarize.py <rep> <out_prefix> <input_bytes>"""
import csv
import json
import subprocess
import sys

rep, out, nbytes = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        # which execution pipe binds an issue-bound kernel (integer LOP3/SHF/ISETP/IADD3 share the ALU pipe)
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active"]
lines = ["# ncu summary of %s" % rep, ""]
for k in keys:
    if k in m:
        lines.append("%-85s %s %s" % (k, m[k][0], m[k][1]))


def num(k):
    v, u = m[k]
    f = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1,
             "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1}.get(u, 1)
    return f * scale


rd, wr, t = num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), num("gpu__time_duration.sum")
inst = float(m["smsp__inst_executed.sum"][0])
lines += ["", "input bytes            %d" % nbytes, "dram read / input      %.4f" % (rd / nbytes),
          "dram bytes per launch  %d" % (rd + wr), "warp-instr per byte    %.4f" % (inst / nbytes),
          "GB/s under ncu (cold, serialised; NOT a bench value) %.1f" % (nbytes / t / 1e9)]
open(out + ".txt", "w").write("\n".join(lines) + "\n")
json.dump({"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
           "input_bytes": int(nbytes), "kernel": m["Kernel Name"][0],
           "note": "ncu --set full, one launch; traffic ~= input once + 16 B per match"},
          open(out + ".json", "w"))
print("\n".join(lines))

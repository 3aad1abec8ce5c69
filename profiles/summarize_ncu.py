#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the handful of numbers the roofline discussion uses.
Usage: python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sass__inst_executed_local_loads",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name}")
        for i, h in enumerate(hdr):
            if h in KEEP:
                print(f"{h:82s} {r[i]:>18s} {units[i]}")
        print("-- warp stall reasons (average warps stalled per issue-active cycle, > 0.05)")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.05:
                    print(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:8.3f}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    try:
        h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    except StopIteration:
        return
    ix = {c: i for i, c in enumerate(rows[h])}
    ops, tot = {}, 0
    for r in rows[h + 1:]:
        if len(r) <= ix["Instructions Executed"]:
            continue
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        key = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")
        n = int(r[ix["Instructions Executed"]] or 0)
        ops[key] = ops.get(key, 0) + n
        tot += n
    print(f"-- executed warp instructions by opcode (total {tot})")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]:
        print(f"{k:14s} {100.0 * v / tot:6.2f} %")


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [--stalls]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    h, u, data = raw(rep)
    ki = h.index("Kernel Name")
    for row in data:
        print("==", row[ki].split("(")[0], "id", row[0])
        for name in WANT:
            if name in h:
                i = h.index(name)
                print("   %-62s %14s %s" % (name, row[i], u[i]))
    if "--stalls" in sys.argv:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        blocks = out.split('"Kernel Name",')
        for blk in blocks[1:]:
            rows = list(csv.reader(io.StringIO('"Kernel Name",' + blk)))
            name = rows[0][1].split("(")[0]
            hh = rows[1]
            dd = [r for r in rows[2:] if len(r) == len(hh)]
            idx = {n: i for i, n in enumerate(hh)}
            tot = sum(int(r[idx["# Samples"]]) for r in dd) or 1
            st = [n for n in hh if n.startswith("stall_") and "Not Issued" not in n]
            agg = sorted(((sum(int(r[idx[n]]) for r in dd), n) for n in st), reverse=True)[:6]
            print("== stalls", name, "samples", tot, " ".join("%s=%.1f%%" % (n, 100.0 * v / tot) for v, n in agg))
            top = sorted(dd, key=lambda r: -int(r[idx["# Samples"]]))[:12]
            for r in top:
                print("   %6s %9s  %s" % (r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:90]))


if __name__ == "__main__":
    main()

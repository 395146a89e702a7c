#!/usr/bin/env python3
"""Per-kernel totals of the LAST step in an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python profiles/launch_summary.py launches.csv [first-kernel-of-a-step, default k_clear]"""
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    first = sys.argv[2] if len(sys.argv) > 2 else "k_clear"
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("trbk::", "").replace("trbr::", "")
        name = name.split("<")[0]
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1000.0
        elif r[ui] in ("ms", "msecond"):
            v *= 1000.0
        launches.append((name, v))
    starts = [i for i, (n, _) in enumerate(launches) if n == first]
    # the capture may stop in the middle of a step (ncu -c N): take the last COMPLETE one
    last = launches[starts[-2]:starts[-1]] if len(starts) > 1 else launches[starts[-1]:]
    tot = {}
    for n, v in last:
        a = tot.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(v for _, v in tot.values())
    print("# last step: %d launches, %.1f us of kernel time (cold-cache, serialised under ncu)" % (len(last), total))
    for n, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-24s %3d %10.1f us %6.1f %%" % (n, c, v, 100.0 * v / total))


if __name__ == "__main__":
    main()

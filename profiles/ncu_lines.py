#!/usr/bin/env python3
"""Per source line totals (instructions executed, stall samples) of every kernel in an .ncu-rep
captured with --import-source on and -lineinfo.
usage: python profiles/ncu_lines.py x.ncu-rep [kernel-substring] [top-n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    i = 0
    seen = set()
    per = {}
    fn = path = None
    hdr = None
    cur_line = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1].split("(")[0]
            continue
        if r[0] == "Line No":
            hdr = {n: k for k, n in enumerate(r)}
            # the two "Source" columns: first = CUDA text, second (after Address) = SASS
            continue
        if hdr is None or len(r) < 8:
            continue
        if r[0] != "":
            cur_line = (path, r[0], r[1].strip())
            continue
        try:
            inst = int(r[hdr["Instructions Executed"]])
            smp = int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        d = per.setdefault(fn, {})
        a = d.setdefault(cur_line, [0, 0, 0])
        a[0] += inst
        a[1] += smp
        a[2] += 1
    for fn, d in per.items():
        if want not in fn:
            continue
        tot_i = sum(v[0] for v in d.values()) or 1
        tot_s = sum(v[1] for v in d.values()) or 1
        print("== %s: %d warp instructions, %d samples" % (fn, tot_i, tot_s))
        for (p, ln, src), v in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print("  %5.1f%% inst %5.1f%% smp %3d sass  %s:%s  %s" % (100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, v[2], p, ln, src[:100]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Executed warp instructions per source line of ONE kernel of an .ncu-rep (captured with --import-source on), joined
with the line table of the same build (nvdisasm -g of the cubin inside libtrb.so), divided by a unit count (e.g. warp-pixels)
so that the numbers read as "instructions per pixel".
usage: python profiles/ncu_perpixel.py x.ncu-rep kernel-regex mangled-name-substring units [top] [launch-index]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kre, mangled, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
    which = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kre], capture_output=True, text=True).stdout
    hdr, data, k = None, [], 0
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            k += 1
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and k == which:
            data.append(r)
    H = {n: i for i, n in enumerate(hdr)}
    base = int(data[0][0], 16)
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "tinyrenderder_b200", "libtrb.so")], cwd=tmp,
                   capture_output=True)
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "trb.sm_100a.cubin")], capture_output=True, text=True).stdout
    cur, inside, off2line = None, False, {}
    for l in sass.splitlines():
        if l.startswith(".text."):
            inside = mangled in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    ops = collections.Counter()
    stalls = collections.Counter()
    for r in data:
        off = int(r[0], 16) - base
        n, s, lsb = int(r[H["Instructions Executed"]]), int(r[H["# Samples"]]), int(r[H["stall_long_sb"]])
        op = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[H["Source"]]).group(2)
        a = per[off2line.get(off)]
        a[0] += n; a[1] += s; a[2] += lsb; a[3][op] += n
        ops[op] += n
        for name in H:
            if name.startswith("stall_") and "Not Issued" not in name:
                stalls[name] += int(r[H[name]])
    tot = sum(a[0] for a in per.values())
    print("total %.1f per unit (%d warp instructions), %d static" % (tot / units, tot, len(data)))
    print("ops:", " ".join("%s %.1f" % (o, n / units) for o, n in ops.most_common(28)))
    print("stalls:", " ".join("%s %d" % (o[6:], n) for o, n in stalls.most_common(8)))
    for ln, a in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(ln, "%.1f" % (a[0] / units), "smp", a[1], "lsb", a[2], {o: round(v / units, 1) for o, v in a[3].most_common(6)})


if __name__ == "__main__":
    main()

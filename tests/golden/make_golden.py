#!/usr/bin/env python3
"""Regenerate tests/golden/golden.json from the reference's own rasterizer.

Runs every case of tests/cases.py through oracle/_ref/libtrb_ref.so (the reference's our_gl.cpp +
tgaimage.cpp compiled in place, needs /root/reference) and records a SHA-256 of every output
array plus the counters.  The order-independent counters the reference cannot report
(`stats_port*`) come from oracle/libtrb_port.so, whose arrays must equal the reference's here.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import __graft_entry__ as g  # noqa: E402

g.build()
import tinyrenderder_b200 as trb  # noqa: E402
import cases  # noqa: E402


def digest(v):
    if isinstance(v, np.ndarray):
        return {"sha256": hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest(), "shape": list(v.shape),
                "dtype": str(v.dtype)}
    if isinstance(v, dict):
        return {k: (float(x) if isinstance(x, float) else int(x)) for k, x in v.items()}
    return v


def main():
    ref = trb.Api(os.path.join(ROOT, "oracle", "_ref", "libtrb_ref.so"), "orc")
    port = trb.Api(os.path.join(ROOT, "oracle", "libtrb_port.so"), "orc")
    out_path = os.path.join(HERE, "golden.json")
    # `--missing`: keep the committed entries and add only the cases that have none yet
    gold = json.load(open(out_path)) if "--missing" in sys.argv and os.path.exists(out_path) else {}
    allc = dict(cases.CASES)
    allc.update(cases.FULL_SIZE_CASES)
    for name, fn in allc.items():
        if name in gold:
            continue
        with trb.Renderer(ref) as r:
            a = fn(ref, r)
        with trb.Renderer(port) as r:
            b = fn(port, r)
        entry = {}
        for k, v in a.items():
            if k.startswith("stats_port"):
                entry[k] = digest(b[k])
                continue
            if isinstance(v, np.ndarray):
                assert np.array_equal(v.view(np.uint8), b[k].view(np.uint8)), (name, k)
            entry[k] = digest(v)
        gold[name] = entry
        print(name, "ok", file=sys.stderr)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Developer aid: kernel timeline of the end-to-end loop of bench.py (config 3) from the library's own
CUDA events (TRB_TRACE), to find where the render stream idles.  usage: python profiles/trace_e2e.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE = os.path.join(ROOT, "gpurun_out", "trace_e2e.txt")
os.makedirs(os.path.dirname(TRACE), exist_ok=True)
for _p in (TRACE, TRACE + ".host"):
    if os.path.exists(_p):
        os.remove(_p)
os.environ["TRB_TRACE"] = TRACE

import numpy as np  # noqa: E402
import torch  # noqa: E402
import tinyrenderder_b200 as trb  # noqa: E402
from tinyrenderder_b200 import scenes  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    mode = sys.argv[2] if len(sys.argv) > 2 else "e2e"
    api = trb.load_cuda()
    r = trb.Renderer(api, 0)
    sc = scenes.orbit_scene()
    pr = api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar)
    pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    seen = {}
    for it in sc.items:
        m = it.mesh
        if id(m) not in seen:
            m.pos, m.nrm, m.uv, m.idx = pin(m.pos), pin(m.nrm), pin(m.uv), pin(m.idx)
            seen[id(m)] = True
        for k, t in list(it.textures.items()):
            if id(t) not in seen:
                seen[id(t)] = pin(t)
                seen[id(seen[id(t)])] = seen[id(t)]
            it.textures[k] = seen[id(t)]
    nv = 32
    buf = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    cols = [[buf((sc.height, sc.width, 3), torch.uint8) for _ in range(nv)] for _ in range(2)]
    resident = scenes.UploadedScene(r, sc)

    def step(s):
        views = scenes.orbit_views(api, [(s * nv + j) % 1024 for j in range(nv)])
        if mode == "plain":
            resident.render(views, pr)
            return
        up = scenes.UploadedScene(r, sc)
        up.render(views, pr)
        r.readback_async(cols[s & 1], None)
        up.free()

    for s in range(3):
        step(s)
    r.readback_wait()
    r.synchronize()
    r.profile_enable(True)
    r.profile_read(reset=True)
    for s in range(steps):
        step(3 + s)
    r.readback_wait()
    r.profile_read(reset=True)      # collects and writes the trace
    rows = []
    for line in open(TRACE):
        a, b, name = line.split()
        rows.append((float(a), float(b), name))
    rows.sort()
    other = ("k_interleave_mesh", "copy_d2h_frames")
    render = [x for x in rows if x[2] not in other]
    for name in other:
        xs = [x for x in rows if x[2] == name]
        if xs:
            print("%s: %d spans, mean %.3f ms, starts %s" % (name, len(xs), sum(x[1] for x in xs) / len(xs),
                                                          " ".join("%.2f" % x[0] for x in xs[:8])))
    stage = [x for x in rows if x[2] == "copy_stage_color"]
    if stage:
        print("copy_stage_color ends:", " ".join("%.2f" % (x[0] + x[1]) for x in stage[:8]))
    span = render[-1][0] + render[-1][1] - render[0][0]
    busy = sum(x[1] for x in render)
    print("mode %s: %d launches on the render stream over %.2f ms (%.2f ms per step), busy %.2f ms (%.1f %%)" % (
        mode, len(render), span, span / steps, busy, 100 * busy / span))
    gaps = []
    for p, q in zip(render[:-1], render[1:]):
        g = q[0] - (p[0] + p[1])
        gaps.append((g, p[2], q[2]))
    agg = {}
    for g, a, b in gaps:
        k = "%s -> %s" % (a, b)
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += g
    print("largest idle gaps by kernel pair (count, total ms, ms per step):")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print("  %-46s %4d %8.3f %8.3f" % (k, n, t, t / steps))
    hp = TRACE + ".host"
    if os.path.exists(hp):
        print("host calls longer than 0.2 ms:")
        for line in open(hp).read().splitlines()[-40:]:
            print("  ", line)


if __name__ == "__main__":
    main()

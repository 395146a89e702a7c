#!/usr/bin/env python3
"""Vertex stage of one rank's share (trb_draw_shard inside a composite group): k_vertex_mesh / k_composite_shade_p2p of rank 0
for n = 1, 2, 4, 8 ranks (contexts of one process on one GPU), icosphere of a given level at 1920x1080.
usage: python profiles/experiments/share_vertex_stage.py [level]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("TRB_MESH_ORDER_MIN_TRIS", "1")
import tinyrenderder_b200 as trb  # noqa: E402
from tinyrenderder_b200 import scenes  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    api = trb.load_cuda()
    m = scenes.icosphere(level)
    w, h = 1920, 1080
    mv, pr = api.lookat([0, 0, 2.2], [0, 0, 0], [0, 1, 0]), api.perspective(60, w / h, 0.1, 10)
    out = {"level": level, "triangles": m.ntris, "vertices": m.nverts}
    for n in (1, 2, 4, 8):
        rs = [trb.Renderer(api) for _ in range(n)]
        meshes = [r.upload_mesh(m.pos, m.nrm, m.uv, m.idx) for r in rs]
        for r in rs:
            r.begin_frame(w, h)
        if n > 1:
            trb.comm_init(rs)
        for it in range(4):
            if it == 2:
                rs[0].synchronize()
                rs[0].profile_enable(True)
                rs[0].profile_read(reset=True)
            for rank, r in enumerate(rs):
                r.begin_frame(w, h)
                if n > 1:
                    r.draw_shard(meshes[rank], mv, pr, rank, n)
                else:
                    r.draw(meshes[rank], mv, pr, ntris=m.ntris)
            if n > 1:
                trb.composite_group(rs)
            else:
                rs[0].end_frame()
        for r in rs:
            r.synchronize()
        prof = rs[0].profile_read(reset=True)
        rs[0].profile_enable(False)
        out["n%d" % n] = {k: round(v[1] / max(1, v[0]), 4) for k, v in prof.items() if v[1] / max(1, v[0]) > 0.004}
        for r in rs:
            r.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

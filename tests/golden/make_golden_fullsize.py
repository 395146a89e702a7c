#!/usr/bin/env python3
"""Regenerate tests/golden/golden_fullsize.json: SHA-256 digests of what the REFERENCE's own rasterizer
(oracle/_ref/libtrb_ref.so = /root/reference/our_gl.cpp + tgaimage.cpp compiled in place) produces for the
BASELINE configs at their stated sizes, driven like main.cpp:647-730.

  c3_orbit   the z-buffer of every one of the 1024 frames of the 1920x1080 orbit of the bench scene
             (bench.py checks the frames of the step it timed against these; tests check a wrap-around set)
  c4_sphere  icosphere level 10 (20 971 520 triangles) at 3840x2160: z-buffer and (flat-shaded, hence exact)
             colour
  c5_soup    100 000 000 sub-pixel triangles at 8192x8192: z-buffer and colour.  The reference submits
             triangles one by one, so the soup is fed to it in ranges of 10 M triangles (same submission
             order, a fraction of the memory)

Needs /root/reference (this container only); takes ~10 minutes on 8 cores.  Usage:
  python tests/golden/make_golden_fullsize.py [c3] [c4] [c5]      (default: all; existing entries are kept)
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "golden_fullsize.json")
REF = os.path.join(ROOT, "oracle", "_ref", "libtrb_ref.so")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


_S = {}


def _c3_init():
    import tinyrenderder_b200 as trb
    from tinyrenderder_b200 import scenes
    api = trb.Api(REF, "orc")
    sc = scenes.orbit_scene()
    r = trb.Renderer(api)
    _S.update(api=api, sc=sc, r=r, up=scenes.UploadedScene(r, sc), scenes=scenes,
              pr=api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar))


def _c3_frame(k):
    api, r, up, scenes = _S["api"], _S["r"], _S["up"], _S["scenes"]
    up.render(scenes.orbit_views(api, [k]), _S["pr"])
    return k, sha(r.read_depth(0))


def make_c3(procs):
    with mp.get_context("fork").Pool(procs, initializer=_c3_init) as pool:
        z = {}
        for k, d in pool.imap_unordered(_c3_frame, range(1024), chunksize=4):
            z[k] = d
            if len(z) % 64 == 0:
                print("c3", len(z), "frames", file=sys.stderr)
    from tinyrenderder_b200 import scenes
    sc = scenes.orbit_scene()
    return {"workload": "c3_orbit_%dx%d_%dtri" % (sc.width, sc.height, sc.ntris), "frames": 1024,
            "z_sha256": [z[k] for k in range(1024)]}


def make_c4():
    import tinyrenderder_b200 as trb
    from tinyrenderder_b200 import scenes
    api = trb.Api(REF, "orc")
    sc = scenes.sphere_scene(10)
    with trb.Renderer(api) as r:
        up = scenes.UploadedScene(r, sc)
        up.render(scenes.sphere_view(api)[None], api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar))
        z, c = r.read_depth(0), r.read_color(0)
    return {"workload": "c4_icosphere_l10_%dx%d_%dtri" % (sc.width, sc.height, sc.ntris), "z_sha256": sha(z),
            "bgr_sha256": sha(c), "pixels_shaded": int(np.isfinite(z).sum())}


def make_c5(n=100_000_000, chunk=10_000_000):
    import tinyrenderder_b200 as trb
    from tinyrenderder_b200 import scenes
    api = trb.Api(REF, "orc")
    w = h = 8192
    _, pos = scenes.triangle_soup(n, w, h, 0.4, 5, True, want_clip=False)
    eye = np.eye(4)
    with trb.Renderer(api) as r:
        r.begin_frame(w, h)
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            m = r.upload_mesh(pos[3 * a:3 * b])
            r.draw(m, eye, eye, kind=0, ntris=b - a)
            r.free_mesh(m)
            print("c5", b, "triangles", file=sys.stderr)
        r.end_frame()
        z, c = r.read_depth(0), r.read_color(0)
    return {"workload": "c5_soup_8192x8192_%dtri_r0.4" % n, "z_sha256": sha(z), "bgr_sha256": sha(c),
            "pixels_shaded": int(np.isfinite(z).sum())}


def main():
    import __graft_entry__ as g
    g.build()
    which = [a for a in sys.argv[1:] if a in ("c3", "c4", "c5")] or ["c3", "c4", "c5"]
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for w in which:
        t = time.time()
        if w == "c3":
            gold["c3_orbit"] = make_c3(min(os.cpu_count() or 1, 8))
        elif w == "c4":
            gold["c4_sphere"] = make_c4()
        else:
            gold["c5_soup"] = make_c5()
        print(w, "done in %.0f s" % (time.time() - t), file=sys.stderr)
        with open(OUT, "w") as f:
            json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

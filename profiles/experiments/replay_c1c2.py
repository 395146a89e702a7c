#!/usr/bin/env python3
"""Launch-bound single-frame configs: plain calls vs one CUDA-graph launch per frame (trb_replay).
Config 1 (head, 800x800, 2520 triangles) and config 2 (two-pass shadow-mapped Phong, 2048x2048): ms per frame over N
frames with a moving camera, device-timed (trb_timer) and host wall clock; the replayed frames are compared with the
plain ones bit for bit first."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tinyrenderder_b200 as trb  # noqa: E402
from tinyrenderder_b200 import scenes  # noqa: E402


def cams(api, n, radius=3.3, height=1.0):
    return [api.lookat([radius * math.cos(2 * math.pi * k / n), height, radius * math.sin(2 * math.pi * k / n)],
                       [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]) for k in range(n)]


def timed(r, fn, n):
    r.synchronize()
    t0 = time.perf_counter()
    r.timer_start()
    for k in range(n):
        fn(k)
    ms = r.timer_stop_ms()
    return ms / n, (time.perf_counter() - t0) * 1e3 / n


def main():
    api = trb.load_cuda()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    out = {}
    with trb.Renderer(api) as r:
        # config 1
        sc = scenes.head_scene()
        pr = api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar)
        up = scenes.UploadedScene(r, sc)
        cs = cams(api, n)
        up.render(cs[0][None], pr)
        want = [None] * 4
        for k in range(4):
            up.render(cs[k * 7][None], pr)
            want[k] = (r.read_depth(0).copy(), r.read_color(0).copy())
        rec = up.record(cs[0][None], pr)
        same = True
        for k in range(4):
            up.replay(rec, cs[k * 7][None], pr)
            same &= bool(np.array_equal(r.read_depth(0).view(np.uint64), want[k][0].view(np.uint64)) and
                         np.array_equal(r.read_color(0), want[k][1]))
        before = r.launch_count()
        up.render(cs[1][None], pr)
        launches = r.launch_count() - before
        plain = timed(r, lambda k: up.render(cs[k][None], pr), n)
        replay = timed(r, lambda k: up.replay(rec, cs[k][None], pr), n)
        same_params = timed(r, lambda k: r.replay(rec), n)
        out["c1_head_800x800"] = {"kernel_launches_per_frame": launches, "bit_identical": same,
                                  "plain_ms_per_frame": {"device": plain[0], "host_wall": plain[1]},
                                  "replay_ms_per_frame": {"device": replay[0], "host_wall": replay[1]},
                                  "replay_unchanged_ms_per_frame": {"device": same_params[0], "host_wall": same_params[1]}}
        r.recording_free(rec)
        up.free()
        # config 2
        sc = scenes.shadow_scene()
        pr = api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar)
        up = scenes.UploadedScene(r, sc)
        cs = cams(api, n, 4.5, 2.0)
        scenes.render_shadowed(up, cs[0], pr)
        want = []
        for k in range(3):
            scenes.render_shadowed(up, cs[k * 5], pr)
            want.append((r.read_depth(0).copy(), r.read_color(0).copy()))
        r.release_shadow_maps()
        r.record_begin()
        scenes.render_shadowed(up, cs[0], pr, release=False)
        rec = r.record_end()

        def replay2(k):
            r.release_shadow_maps()
            with r.collect_draws() as draws:
                scenes.render_shadowed(up, cs[k], pr, release=False)
            r.replay(rec, draws)

        same = True
        for k in range(3):
            replay2(k * 5)
            same &= bool(np.array_equal(r.read_depth(0).view(np.uint64), want[k][0].view(np.uint64)) and
                         np.array_equal(r.read_color(0), want[k][1]))
        before = r.launch_count()
        scenes.render_shadowed(up, cs[1], pr)
        launches = r.launch_count() - before
        plain = timed(r, lambda k: scenes.render_shadowed(up, cs[k], pr), n)
        replay = timed(r, replay2, n)
        same_params = timed(r, lambda k: r.replay(rec), n)
        out["c2_shadow_2048x2048"] = {"kernel_launches_per_frame": launches, "bit_identical": same,
                                      "plain_ms_per_frame": {"device": plain[0], "host_wall": plain[1]},
                                      "replay_ms_per_frame": {"device": replay[0], "host_wall": replay[1]},
                                      "replay_unchanged_ms_per_frame": {"device": same_params[0], "host_wall": same_params[1]}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Frame recordings (trb_record_begin / trb_record_end / trb_replay): a launch-bound frame as one CUDA graph launch.
A replayed frame must be bit for bit the frame the same calls produce when issued one by one - for the recorded camera
and for every other camera whose matrices / uniform blocks the replay is given."""
import numpy as np
import pytest

import tinyrenderder_b200 as trb
from tinyrenderder_b200 import scenes

pytestmark = pytest.mark.gpu


def _frame(r, nviews=1):
    return [(r.read_depth(v).copy(), r.read_color(v).copy()) for v in range(nviews)]


def _same(got, want):
    assert len(got) == len(want)
    for (gz, gc), (wz, wc) in zip(got, want):
        assert np.array_equal(gz.view(np.uint64), wz.view(np.uint64))
        assert np.array_equal(gc, wc)


def test_head_frame_replays_for_other_cameras(cuda_api):
    """config 1 shape (one lit model, one frame): record with camera 0, replay unchanged, with camera 1, unchanged again
    (the new parameters stay), and back with camera 0"""
    sc = scenes.head_scene(320, 320, tex_size=64)
    pr = cuda_api.perspective(sc.fov, 1.0, sc.znear, sc.zfar)
    cams = [cuda_api.lookat(e, [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]) for e in ([1.0, 1.0, 3.0], [-2.0, 0.5, 2.0], [0.2, 2.5, 1.0])]
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        want = []
        for v in cams:
            up.render(v[None], pr)
            want.append(_frame(r))
        before = r.launch_count()
        rec = up.record(cams[0][None], pr)
        _same(_frame(r), want[0])                        # record_end ran the frame
        per_frame = (r.launch_count() - before) // 2     # the warm-up frame + the recorded one
        r.replay(rec)
        _same(_frame(r), want[0])
        for k in (1, 2, 1, 0):
            n0 = r.launch_count()
            up.replay(rec, cams[k][None], pr)
            assert r.launch_count() - n0 == per_frame     # gpu_launches keeps counting kernels, not graph launches
            _same(_frame(r), want[k])
            if k == 1:
                r.replay(rec)
                _same(_frame(r), want[1])
        # first use of another subsystem (the TGA packetiser allocates its work areas) does not make the recording stale
        files = r.encode_tga()
        assert len(files) == 1 and len(files[0]) > 18
        # plain frames and replays interleave
        up.render(cams[2][None], pr)
        _same(_frame(r), want[2])
        up.replay(rec, cams[0][None], pr)
        _same(_frame(r), want[0])
        r.recording_free(rec)
        with pytest.raises(trb.TrbError):
            r.replay(rec)


def test_orbit_batch_with_snapshot_and_restore_replays(cuda_api):
    """config 3 shape: three models, two flushes, z snapshot before the eyes and the pointer-swap restore after them,
    several cameras per launch set; replayed for other orbit positions"""
    sc = scenes.orbit_scene(320, 180, room_quads=((16, 8), (16, 4), (8, 8)), tex_size=64)
    pr = cuda_api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
    sets = [scenes.orbit_views(cuda_api, ks) for ks in ([3, 400, 900], [10, 500, 1000], [77, 78, 79])]
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        want = []
        for vs in sets:
            up.render(vs, pr, cull=False)
            want.append(_frame(r, 3))
        rec = up.record(sets[0], pr, cull=False)
        _same(_frame(r, 3), want[0])
        for k in (1, 2, 0, 1):
            up.replay(rec, sets[k], pr, cull=False)
            _same(_frame(r, 3), want[k])
            st = r.stats(0)
            assert st["triangles_submitted"] == sc.ntris


def test_two_pass_shadow_frame_replays(cuda_api):
    """config 2 shape: depth pass from the light, kept as the shadow map, camera pass that samples it - two frames and a
    shadow-map plane inside one recording"""
    sc = scenes.shadow_scene(256, 256, body_res=(16, 12), ground_quads=8, tex_size=64)
    pr = cuda_api.perspective(sc.fov, 1.0, sc.znear, sc.zfar)
    cams = [cuda_api.lookat(e, [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]) for e in ([2.0, 2.0, 4.0], [-3.0, 1.5, 3.0])]
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        want = []
        for v in cams:
            scenes.render_shadowed(up, v, pr)
            want.append(_frame(r))
        r.release_shadow_maps()                      # the plane the recorded frame will keep comes from the pool
        r.record_begin()
        scenes.render_shadowed(up, cams[0], pr, release=False)
        rec = r.record_end()
        _same(_frame(r), want[0])
        for k in (1, 0, 1):
            r.release_shadow_maps()                  # as after every frame of the plain loop
            with r.collect_draws() as draws:
                scenes.render_shadowed(up, cams[k], pr, release=False)
            r.replay(rec, draws)
            _same(_frame(r), want[k])
        r.replay(rec)                                # the plane still held by the frame before is the recording's own: fine
        _same(_frame(r), want[1])


def test_recording_rules(cuda_api):
    """what a recording refuses: growing buffers (no warm-up frame), synchronising calls inside, stale addresses"""
    sc = scenes.head_scene(160, 160, tex_size=32)
    pr = cuda_api.perspective(sc.fov, 1.0, sc.znear, sc.zfar)
    cam = scenes.head_view(cuda_api)
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        r.record_begin()                             # fresh context: nothing has its size yet
        with pytest.raises(trb.TrbError, match="render the frame once"):
            up.render(cam[None], pr)
        with pytest.raises(trb.TrbError):
            r.record_end()
        up.render(cam[None], pr)                     # the context is still usable
        ref = _frame(r)
        r.record_begin()
        up.render(cam[None], pr)
        with pytest.raises(trb.TrbError, match="not inside"):
            r.read_color(0)
        rec = r.record_end()
        _same(_frame(r), ref)
        with pytest.raises(trb.TrbError, match="one entry per draw"):
            r.replay(rec, [{}, {}])
        up.free()                                    # the recording samples these textures
        with pytest.raises(trb.TrbError, match="stale"):
            r.replay(rec)

"""Parity cases shared by the CPU tests (oracle vs golden vectors) and the GPU tests (CUDA vs
oracle).  Every case is a function(api, renderer) -> dict of numpy arrays / ints; the same
function is run on every backend through the identical C ABI, like the reference's draw loops
would be (main.cpp:660-666)."""
import math

import numpy as np

from tinyrenderder_b200 import (PhongUniforms, SHADER_DEPTH, SHADER_EYE, SHADER_FLAT_BARY, SHADER_GOURAUD, SHADER_PHONG,
                                SHADER_SHADOW_PHONG, scenes)

INF = float("inf")


def _grab(r, views=1, post=False, stats=True):
    out = {}
    for v in range(views):
        sfx = "" if views == 1 else "_v%d" % v
        out["z" + sfx] = r.read_depth(v)
        out["bgr" + sfx] = r.read_color(v)
        if post:
            out["ao" + sfx] = r.ssao(v)
            out["zimg" + sfx] = r.depth_image(v)
            out["final" + sfx] = r.composite_ao(v)
        if stats:
            s = r.stats(v)
            out["stats" + sfx] = {k: s[k] for k in ("triangles_submitted", "pixels_shaded")}
            out["stats_port" + sfx] = {k: s[k] for k in ("triangles_binned", "fragments_covered", "bbox_min_x",
                                                         "bbox_min_y", "bbox_max_x", "bbox_max_y", "z_min",
                                                         "z_max_covered")}
    return out


def _screen_tri(pts, z, w=64, h=64):
    """clip = ((sx-w/2)/(w/2), (sy-h/2)/(h/2), z, 1) like SURVEY K2"""
    return np.array([[(x - w / 2) / (w / 2), (y - h / 2) / (h / 2), z, 1.0] for x, y in pts]).reshape(1, 12)


# ---- SURVEY section 4 known-answer tests --------------------------------------------------------
def k1(api, r):
    r.begin_frame(64, 64)
    mesh = r.upload_mesh(np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], dtype=np.float32))
    r.draw(mesh, api.lookat([0, 0, 3], [0, 0, 0], [0, 1, 0]), api.perspective(60, 1, 0.1, 100), ntris=1)
    r.end_frame()
    return _grab(r)


def k2(api, r):
    r.begin_frame(64, 64)
    r.submit_clip_triangles(_screen_tri([(8.5, 8.5), (24.5, 8.5), (24.5, 24.5)], 0.25))
    r.submit_clip_triangles(_screen_tri([(8.5, 8.5), (24.5, 24.5), (8.5, 24.5)], 0.25))
    r.end_frame()
    return _grab(r)


def k2_flush_between(api, r):
    """same as K2 but the first triangle is shaded before the second is drawn: the tie pixels on
    the shared diagonal must stay with the first triangle (our_gl.cpp:165 is a strict <)"""
    r.begin_frame(64, 64)
    r.submit_clip_triangles(_screen_tri([(8.5, 8.5), (24.5, 8.5), (24.5, 24.5)], 0.25))
    r.flush()
    r.submit_clip_triangles(_screen_tri([(8.5, 8.5), (24.5, 24.5), (8.5, 24.5)], 0.25))
    r.end_frame()
    return _grab(r)


def k3(api, r):
    r.begin_frame(64, 64)
    r.submit_clip_triangles(np.array([-0.5, -0.5, 0, 1, 0.5, -0.5, 0, 1, 0, 0.5, 3.0, 1.0]))
    r.end_frame()
    return _grab(r)


def k4(api, r):
    r.begin_frame(64, 64)
    r.submit_clip_triangles(np.array([-0.5, -0.5, 0, 1, 0, 0.5, 0, 1, 0.5, -0.5, 0, 1.0]))
    r.end_frame()
    return _grab(r)


def _k5(order):
    def run(api, r):
        r.begin_frame(64, 64)
        for z in order:
            r.submit_clip_triangles(np.array([-0.5, -0.5, z, 1, 0.5, -0.5, z, 1, 0, 0.5, z, 1.0]))
        r.end_frame()
        return _grab(r)
    return run


k5_far_near = _k5([0.5, -0.5])
k5_near_far = _k5([-0.5, 0.5])


def k6(api, r):
    r.begin_frame(64, 64)
    r.submit_clip_triangles(np.array([-0.5, -0.5, 0, 1, 1.0, -0.5e-11, 0, 1e-11, 0, 0.5, 0, 1.0]))
    r.end_frame()
    return _grab(r)


def _k7(seed, radius, n, w, h):
    def run(api, r):
        clip, _ = scenes.triangle_soup(n, w, h, radius, seed, False)
        r.begin_frame(w, h)
        r.submit_clip_triangles(clip)
        r.end_frame()
        return _grab(r)
    return run


k7a = _k7(1, 0.4, 2_000_000, 2048, 2048)
k7b = _k7(2, 2.0, 1_000_000, 2048, 2048)
k7c = _k7(3, 30.0, 20_000, 1920, 1080)
k7a_small = _k7(1, 0.4, 100_000, 512, 512)
k7b_small = _k7(2, 2.0, 50_000, 512, 512)
k7c_small = _k7(3, 30.0, 2_000, 640, 360)


# ---- edge cases ------------------------------------------------------------------------------
def empty_frame(api, r):
    r.begin_frame(33, 17)
    r.set_clear_color(7, 8, 9)
    r.begin_frame(33, 17)
    r.submit_clip_triangles(np.zeros((0, 12)))
    r.end_frame()
    out = _grab(r, post=True)
    r.set_clear_color(0, 0, 0)
    return out


def one_pixel(api, r):
    r.begin_frame(1, 1)
    r.submit_clip_triangles(np.array([-3, -3, 0.1, 1, 3, -3, 0.2, 1, 0, 3, 0.3, 1.0]))
    r.end_frame()
    return _grab(r)


def rejects(api, r):
    """w <= 1e-12, NaN / inf vertices, all-z-out, z partially out (kept, our_gl.cpp:103-106), back face,
    zero area, off-screen bbox, screen coords beyond 2^31"""
    nan = float("nan")
    tris = [
        [-0.5, -0.5, 0, 1, 0.5, -0.5, 0, 0.0, 0, 0.5, 0, 1],          # w == 0
        [-0.5, -0.5, 0, 1, 0.5, -0.5, 0, -1.0, 0, 0.5, 0, 1],         # w < 0
        [-0.5, -0.5, 0, 1, 0.5, -0.5, 0, 1e-13, 0, 0.5, 0, 1],        # w <= 1e-12
        [nan, -0.5, 0, 1, 0.5, -0.5, 0, 1, 0, 0.5, 0, 1],             # NaN x
        [-0.5, -0.5, 0, 1, 0.5, INF, 0, 1, 0, 0.5, 0, 1],             # inf y
        [-0.5, -0.5, 0, nan, 0.5, -0.5, 0, 1, 0, 0.5, 0, 1],          # NaN w
        [-0.5, -0.5, 2, 1, 0.5, -0.5, 3, 1, 0, 0.5, -4, 1],           # all z out
        [-0.9, -0.9, 2, 1, -0.1, -0.9, 0.5, 1, -0.5, -0.1, -4, 1],    # two z out: drawn
        [0.1, 0.1, 0, 1, 0.5, 0.9, 0, 1, 0.9, 0.1, 0, 1],             # clockwise
        [0.1, 0.1, 0, 1, 0.5, 0.5, 0, 1, 0.9, 0.9, 0, 1],             # zero area
        [1.5, 1.5, 0, 1, 2.5, 1.5, 0, 1, 2.0, 2.5, 0, 1],             # off screen
        [-0.5, -0.5, 0, 1, 1.0, -0.5e-11, 0, 1e-11, 0, 0.5, 0, 1],    # K6
        [-3.0, -0.2, 0.3, 1, 3.0, -0.2, 0.3, 1, 0.0, 0.7, 0.3, 2],    # partially off screen, w != 1
        [1e-7, 1e-7, 0.2, 1, 0.6, 1e-7, 0.2, 1, 0.3, 0.6 + 1e-9, 0.2, 1],
        [0.2, -0.8, 0.1, 1e-11, 0.4, -0.8, 0.1, 1e-11, 0.3, -0.6, 0.1, 1e-11],  # w tiny but > 1e-12
    ]
    r.begin_frame(100, 70)
    r.submit_clip_triangles(np.array(tris, dtype=np.float64))
    r.end_frame()
    return _grab(r)


def signed_zero_ties(api, r):
    """+0.0 and -0.0 depths are equal for the reference's `<`: whoever is submitted first keeps the
    pixel AND its own sign bit in the z-buffer"""
    a = [-0.8, -0.8, 0.0, 1, 0.8, -0.8, 0.0, 1, 0.0, 0.8, 0.0, 1]
    b = [-0.8, -0.8, -0.0, 1, 0.8, -0.8, -0.0, 1, 0.0, 0.8, -0.0, 1]
    c = [-0.6, -0.7, -0.0, 1, 0.9, -0.6, -0.0, 1, 0.1, 0.9, -0.0, 1]
    out = {}
    for name, order in (("pm", [a, b, c]), ("mp", [b, a, c]), ("cab", [c, a, b])):
        r.begin_frame(48, 48)
        r.submit_clip_triangles(np.array(order, dtype=np.float64))
        r.end_frame()
        g = _grab(r, stats=False)
        out["z_" + name], out["bgr_" + name] = g["z"], g["bgr"]
    r.begin_frame(48, 48)   # same with a flush between the two
    r.submit_clip_triangles(np.array([b], dtype=np.float64))
    r.flush()
    r.submit_clip_triangles(np.array([a, c], dtype=np.float64))
    r.end_frame()
    g = _grab(r, stats=False)
    out["z_flush"], out["bgr_flush"] = g["z"], g["bgr"]
    return out


def duplicate_triangles(api, r):
    """coplanar duplicates across draws: first submitted wins everywhere"""
    rng = np.random.default_rng(5)
    t = np.array([[-0.7, -0.6, 0.3, 1, 0.8, -0.7, 0.1, 1.5, 0.1, 0.9, -0.2, 0.7]])
    r.begin_frame(130, 90)
    for _ in range(3):
        r.submit_clip_triangles(np.repeat(t, 5, axis=0))
    r.end_frame()
    return _grab(r)


def big_triangles(api, r):
    """full-screen and multi-tile triangles (pixel-owner path), front to back and back to front"""
    rng = np.random.default_rng(11)
    tris = []
    for i in range(40):
        c = rng.uniform(-1, 1, 2)
        d = rng.uniform(0.4, 2.5)
        a0 = rng.uniform(0, 2 * math.pi)
        pts = [(c[0] + d * math.cos(a0 + k * 2.0944), c[1] + d * math.sin(a0 + k * 2.0944)) for k in range(3)]
        zs = rng.uniform(-0.9, 0.9, 3)
        ws = rng.uniform(0.5, 2.0, 3)
        tris.append([v for (x, y), z, w in zip(pts, zs, ws) for v in (x * w, y * w, z * w, w)])
    r.begin_frame(200, 120)
    r.submit_clip_triangles(np.array(tris))
    r.end_frame()
    return _grab(r)


def queue_overflow(api, r):
    """> 1024 winning candidates inside one tile in one chunk: 256 mid-size triangles stacked back to
    front in a single 16x16 tile"""
    tris = []
    for i in range(256):
        z = 0.9 - i * 0.005
        x0, y0 = 16 + (i % 5) * 0.37, 16 + (i % 7) * 0.29
        tris.append(_screen_tri([(x0, y0), (x0 + 8.3, y0 + 0.4), (x0 + 3.1, y0 + 7.7)], z)[0])
    r.begin_frame(64, 64)
    r.submit_clip_triangles(np.array(tris))
    r.end_frame()
    return _grab(r)


def dense_tile(api, r):
    """several chunks (> 256 triangles) in single tiles, mixed sizes, many ties"""
    rng = np.random.default_rng(3)
    n = 3000
    cx, cy = rng.uniform(20, 44, n), rng.uniform(20, 44, n)
    rad = rng.choice([0.3, 1.0, 4.0, 12.0], n)
    z = np.round(rng.uniform(-1, 1, n), 1)   # coarse depths -> exact ties between triangles
    tris = np.empty((n, 12))
    for i in range(n):
        tris[i] = _screen_tri([(cx[i] - rad[i], cy[i] - rad[i]), (cx[i] + rad[i], cy[i] - rad[i]),
                               (cx[i], cy[i] + rad[i])], z[i])[0]
    r.begin_frame(64, 64)
    r.submit_clip_triangles(tris)
    r.end_frame()
    return _grab(r)


# ---- config-like scenes --------------------------------------------------------------------------
def soup_mesh_fp32(api, r):
    """config-5 path at small scale: fp32 soup through the normal mesh draw with identity matrices"""
    w = h = 640
    _, pos = scenes.triangle_soup(150_000, w, h, 0.4, 5, True)
    mesh = r.upload_mesh(pos)
    r.begin_frame(w, h)
    r.draw(mesh, np.eye(4), np.eye(4), ntris=pos.shape[0] // 3)
    r.end_frame()
    return _grab(r)


def head(size=400, tex=256, res=(36, 35)):
    def run(api, r):
        sc = scenes.head_scene(size, size, res[0], res[1], tex_size=tex)
        up = scenes.UploadedScene(r, sc)
        up.render(scenes.head_view(api)[None], api.perspective(sc.fov, 1.0, sc.znear, sc.zfar))
        return _grab(r, post=True)
    return run


head_small = head(240, 128, (24, 18))
head_c1 = head(800, 1024)


def orbit(width, height, frames, room_quads, tex):
    def run(api, r):
        sc = scenes.orbit_scene(width, height, room_quads=room_quads, tex_size=tex)
        up = scenes.UploadedScene(r, sc)
        views = scenes.orbit_views(api, frames)
        up.render(views, api.perspective(sc.fov, width / height, sc.znear, sc.zfar))
        return _grab(r, views=len(frames), post=True)
    return run


orbit_small = orbit(320, 180, [0, 300, 700], ((32, 16), (32, 8), (16, 16)), 128)
orbit_mid = orbit(960, 540, [5, 517], ((128, 64), (128, 32), (64, 64)), 256)
# BASELINE config 3 at its stated size: the bench scene (262 144-triangle room + head + eyes, 1024^2 maps) at 1920x1080,
# one batch whose frame indices straddle the k mod 1024 wrap of the orbit, plus a frame from the far side
ORBIT_C3_FRAMES = [1022, 1023, 0, 1, 517]
orbit_c3 = orbit(1920, 1080, ORBIT_C3_FRAMES, ((256, 128), (256, 64), (128, 128)), 1024)


def sphere_c4(api, r):
    """BASELINE config 4 at its stated size: icosphere level 10 (20 971 520 triangles) at 3840x2160, one draw"""
    sc = scenes.sphere_scene(10)
    up = scenes.UploadedScene(r, sc)
    up.render(scenes.sphere_view(api)[None], api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar))
    return _grab(r)


def soup_c5(api, r):
    """BASELINE config 5 at its stated size: 100 000 000 sub-pixel triangles at 8192x8192 through the mesh path
    (V = 3T, implicit indices, identity matrices), one draw.  Digest-only: see tests/golden/golden_fullsize.json"""
    w = h = 8192
    n = 100_000_000
    _, pos = scenes.triangle_soup(n, w, h, 0.4, 5, True, want_clip=False)
    mesh = r.upload_mesh(pos)
    r.begin_frame(w, h)
    r.draw(mesh, np.eye(4), np.eye(4), ntris=n)
    r.end_frame()
    return {"z": r.read_depth(0), "bgr": r.read_color(0)}


def orbit_culled(api, r):
    """model-level frustum cull (main.cpp:623-624, 647, 680, 706; bug-for-bug planes): a batch of three cameras of which
    two drop the head - and with it the eyes, which main() tests with the head's box - while the third keeps everything"""
    sc = scenes.orbit_scene(320, 180, room_quads=((32, 16), (32, 8), (16, 16)), tex_size=128)
    up = scenes.UploadedScene(r, sc)
    # camera 0 looks straight at the head, and the reference's frustum (planes of the transposed matrix) still drops it
    cams = [((-1.869, 0.582, -0.816), (0.412, 2.723, -0.129)), ((-3.4019, 2.2001, 1.8026), (1.3555, 1.5116, -0.9686)), ((0, 5, 0), (3, 5, 0))]
    views = np.stack([api.lookat(e, c, [0.0, 1.0, 0.0]) for e, c in cams])
    up.render(views, api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar))
    assert up.culled == 4, up.culled            # head + eyes for two of the three cameras
    out = _grab(r, views=3, post=True, stats=False)
    # the same cameras with the cull switched off draw the head: the test would be vacuous if that changed nothing
    up.render(views, api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar), cull=False)
    assert not np.array_equal(r.read_depth(0).view(np.uint64), out["z_v0"].view(np.uint64))
    return out


def depth_only_then_color(api, r):
    """DEPTH draws write z but no colour; later draws are tested against that z"""
    r.begin_frame(96, 96)
    r.submit_clip_triangles(_screen_tri([(10, 10), (80, 12), (40, 85)], 0.1, 96, 96), kind=SHADER_DEPTH)
    r.submit_clip_triangles(_screen_tri([(5, 40), (90, 30), (50, 90)], 0.4, 96, 96))
    r.end_frame()
    return _grab(r)


def snapshot_restore_twice(api, r):
    """`saved = zbuffer` ... `zbuffer = saved` (main.cpp:700,730) used repeatedly: colours of everything
    drawn persist, depths return to the snapshot each time, and draws after a restore are tested against it"""
    r.begin_frame(96, 96)
    r.submit_clip_triangles(_screen_tri([(10, 10), (80, 12), (40, 85)], 0.5, 96, 96))
    r.depth_snapshot()
    r.submit_clip_triangles(_screen_tri([(5, 40), (90, 30), (50, 90)], 0.2, 96, 96))    # nearer: wins where it overlaps
    r.depth_restore()
    r.depth_restore()                                                                    # idempotent
    r.submit_clip_triangles(_screen_tri([(20, 5), (95, 50), (30, 70)], 0.35, 96, 96))   # behind the 0.2 one, but that depth is gone
    r.depth_restore()
    r.submit_clip_triangles(_screen_tri([(0, 0), (60, 5), (10, 60)], 0.7, 96, 96))      # behind the snapshot where they overlap
    r.end_frame()
    return _grab(r)


def snapshot_signed_zero(api, r):
    """a -0.0 depth that is still unshaded when the z-buffer is saved: the saved copy must hold the sign bit too
    (the device snapshots while it shades; the shade pass is what restores the exact bits)"""
    b = [-0.8, -0.8, -0.0, 1, 0.8, -0.8, -0.0, 1, 0.0, 0.8, -0.0, 1]
    r.begin_frame(64, 48)
    r.submit_clip_triangles(np.array([b], dtype=np.float64))
    r.depth_snapshot()
    r.submit_clip_triangles(_screen_tri([(5, 5), (60, 8), (30, 44)], -0.5, 64, 48))     # nearer: overwrites the zeros
    r.depth_restore()
    r.submit_clip_triangles(_screen_tri([(2, 2), (40, 4), (10, 40)], 0.25, 64, 48))     # behind -0.0 where they overlap
    r.end_frame()
    return _grab(r)


def snapshot_small_triangles(api, r):
    """sub-pixel and few-pixel triangles (the direct path's kind) drawn between a snapshot and its restore, on a frame
    whose last tile row and column are ragged: the tile-granular snapshot (k_snap_save / k_snap_restore) has to save every
    tile such a draw changes, whichever raster path takes it, and nothing may leak through the restore"""
    w, h = 100, 70
    rng = np.random.default_rng(77)

    def cloud(n, zlo, zhi, size):
        c = rng.uniform([1, 1], [w - 1, h - 1], size=(n, 2))
        tris = []
        for (cx, cy), z in zip(c, rng.uniform(zlo, zhi, size=n)):
            d = rng.uniform(-size, size, size=(3, 2))
            tris.append(_screen_tri([(cx + d[0, 0], cy + d[0, 1]), (cx + d[1, 0], cy + d[1, 1]), (cx + d[2, 0], cy + d[2, 1])],
                                    float(z), w, h)[0])
        return np.array(tris, dtype=np.float64)

    r.begin_frame(w, h)
    r.submit_clip_triangles(_screen_tri([(2, 2), (97, 6), (40, 68)], 0.5, w, h))
    r.submit_clip_triangles(cloud(150, 0.3, 0.7, 2.5))
    r.depth_snapshot()
    r.submit_clip_triangles(cloud(300, -0.2, 0.9, 1.5))       # nearer and farther than the saved state, all over the frame
    r.submit_clip_triangles(_screen_tri([(60, 30), (99.5, 40), (70, 69.5)], 0.1, w, h))
    r.depth_restore()
    r.submit_clip_triangles(cloud(200, 0.2, 0.8, 3.0))        # tested against the restored depths
    r.depth_restore()
    r.submit_clip_triangles(cloud(50, 0.45, 0.55, 6.0))
    r.end_frame()
    return _grab(r)


def sub_range_draws(api, r):
    """drawing a mesh as three triangle ranges == drawing it at once (config-4 sharding unit)"""
    m = scenes.icosphere(3)
    h = r.upload_mesh(m.pos, m.nrm, m.uv, m.idx)
    mv = api.lookat([0, 0, 2.2], [0, 0, 0], [0, 1, 0])
    pr = api.perspective(60, 1.5, 0.1, 10)
    r.begin_frame(300, 200)
    n = m.ntris
    for a, b in ((0, n // 3), (n // 3, n // 2), (n // 2, n)):
        r.draw(h, mv, pr, first_tri=a, ntris=b - a)
    r.end_frame()
    return _grab(r)


def soup_duplicates_lit(api, r):
    """a lit SOUP (no index buffer: vertex 3t + k) that lists every triangle of a small sphere three times, shuffled, drawn
    as two triangle ranges with PhongShader: every covered pixel is a depth tie between three ids, the first submitted must
    win it, and its normals / texture coordinates must be those of ITS vertices - also when the backend keeps the soup's
    vertex arrays in a processing order of its own (mesh_order.cu, TRB_MESH_ORDER_MIN_TRIS)"""
    sc = scenes.head_scene(200, 160, 16, 12, tex_size=64)
    m = sc.items[0].mesh
    tri = m.idx.reshape(-1, 3)
    rng = np.random.Generator(np.random.PCG64(12))
    tri = np.concatenate([tri, tri, tri], axis=0)[rng.permutation(3 * tri.shape[0])].reshape(-1)
    flat = scenes.MeshData(m.pos[tri], m.nrm[tri], m.uv[tri], None, "soup")
    sc.items[0].mesh = flat
    up = scenes.UploadedScene(r, sc)
    view = scenes.head_view(api)[None]
    pr = api.perspective(sc.fov, 200 / 160, sc.znear, sc.zfar)
    r.begin_frame(sc.width, sc.height)
    mvs, uni = up._item_params(sc.items[0], view, None)
    n = tri.size // 3
    for a, b in ((0, n // 4), (n // 4, n)):
        r.draw(up.mesh_h[id(flat)], mvs, pr, kind=sc.items[0].kind, uniforms=uni, first_tri=a, ntris=b - a)
    r.end_frame()
    return _grab(r)


def indexed_duplicates(api, r):
    """an INDEXED mesh whose index buffer lists every triangle of a small sphere three times, shuffled, drawn as two
    triangle ranges: every covered pixel is a depth tie between three ids from different parts of the buffer, and
    the first submitted must win it (our_gl.cpp:160-166) - also when the backend visits the triangles in another
    order (mesh processing order, TRB_MESH_ORDER_MIN_TRIS)"""
    m = scenes.icosphere(3)
    tri = m.idx.reshape(-1, 3)
    rng = np.random.Generator(np.random.PCG64(11))
    tri = np.concatenate([tri, tri, tri], axis=0)[rng.permutation(3 * tri.shape[0])]
    h = r.upload_mesh(m.pos, m.nrm, m.uv, np.ascontiguousarray(tri.reshape(-1)).astype(m.idx.dtype))
    mv = api.lookat([0.3, 0.2, 2.4], [0, 0, 0], [0, 1, 0])
    pr = api.perspective(55, 1.25, 0.1, 10)
    r.begin_frame(320, 256)
    n = tri.shape[0]
    for a, b in ((0, n // 2), (n // 2, n)):
        r.draw(h, mv, pr, first_tri=a, ntris=b - a)
    r.end_frame()
    return _grab(r)


def lit_clip_triangles(api, r):
    """immediate mode behind rasterize(clip, PhongShader, fb): clip + varyings supplied by the caller"""
    m = scenes.uv_sphere(12, 9)
    mv = scenes.head_view(api)
    pr = api.perspective(60, 1.0, 0.1, 100)
    pos = m.pos.astype(np.float64)
    nrm = m.nrm.astype(np.float64)
    n = m.ntris
    clip = np.empty((n, 3, 4))
    vary = np.empty((n, 3, 8))
    for t in range(n):
        for k in range(3):
            vi = m.idx[3 * t + k]
            p4 = np.array([pos[vi, 0], pos[vi, 1], pos[vi, 2], 1.0])
            n4 = np.array([nrm[vi, 0], nrm[vi, 1], nrm[vi, 2], 0.0])
            pe = np.array([_dot4(mv[i], p4) for i in range(4)])
            ne = np.array([_dot4(mv[i], n4) for i in range(4)])
            clip[t, k] = [_dot4(pr[i], pe) for i in range(4)]
            vary[t, k] = [m.uv[vi, 0], m.uv[vi, 1], pe[0], pe[1], pe[2], ne[0], ne[1], ne[2]]
    tex = scenes.texture_diffuse(64, 11)
    out = {}
    for kind, name in ((SHADER_PHONG, "phong"), (SHADER_EYE, "eye")):
        r.begin_frame(150, 150)
        u = PhongUniforms()
        u.key_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.KEY_LIGHT))
        u.fill_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.FILL_LIGHT))
        u.rim_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.RIM_LIGHT))
        u.normal_map_strength = 0.7
        u.diffuse = r.upload_texture(tex)
        r.submit_clip_triangles(clip, vary, mv, kind, u)
        r.end_frame()
        g = _grab(r, stats=False)
        out["z_" + name], out["bgr_" + name] = g["z"], g["bgr"]
    return out


def shadow(size, shadow_size, tex, res, ground, kind=SHADER_SHADOW_PHONG):
    """config 2: depth pass from the light kept as a shadow map, then the shadow-mapped colour pass"""
    def run(api, r):
        sc = scenes.shadow_scene(size, size, body_res=res, ground_quads=ground, tex_size=tex)
        up = scenes.UploadedScene(r, sc)
        scenes.render_shadowed(up, scenes.head_view(api), api.perspective(sc.fov, 1.0, sc.znear, sc.zfar),
                               shadow_size=(shadow_size, shadow_size), kind=kind)
        return _grab(r, post=True)
    return run


shadow_small = shadow(256, 200, 64, (20, 14), 12)
gouraud_small = shadow(200, 64, 64, (20, 14), 8, kind=SHADER_GOURAUD)
shadow_c2 = shadow(2048, 2048, 1024, (42, 60), 32)


def _dot4(row, v):
    s = 0.0
    for i in range(4):
        s = s + row[i] * v[i]
    return s


# name -> (function, tier): "fast" cases run everywhere; "slow" only in the full-size tests
CASES = {
    "k1": k1, "k2": k2, "k2_flush_between": k2_flush_between, "k3": k3, "k4": k4,
    "k5_far_near": k5_far_near, "k5_near_far": k5_near_far, "k6": k6,
    "k7a_small": k7a_small, "k7b_small": k7b_small, "k7c_small": k7c_small,
    "empty_frame": empty_frame, "one_pixel": one_pixel, "rejects": rejects,
    "signed_zero_ties": signed_zero_ties, "duplicate_triangles": duplicate_triangles,
    "big_triangles": big_triangles, "queue_overflow": queue_overflow, "dense_tile": dense_tile,
    "soup_mesh_fp32": soup_mesh_fp32, "head_small": head_small, "orbit_small": orbit_small,
    "depth_only_then_color": depth_only_then_color, "sub_range_draws": sub_range_draws,
    "snapshot_restore_twice": snapshot_restore_twice, "snapshot_signed_zero": snapshot_signed_zero, "snapshot_small_triangles": snapshot_small_triangles,
    "lit_clip_triangles": lit_clip_triangles, "shadow_small": shadow_small, "gouraud_small": gouraud_small,
    "orbit_culled": orbit_culled, "indexed_duplicates": indexed_duplicates, "soup_duplicates_lit": soup_duplicates_lit,
}
FULL_SIZE_CASES = {"k7a": k7a, "k7b": k7b, "k7c": k7c, "head_c1": head_c1, "orbit_mid": orbit_mid,
                   "shadow_c2": shadow_c2, "orbit_c3": orbit_c3, "sphere_c4": sphere_c4}
# too large for an oracle run inside the GPU test session: checked against SHA-256 digests of the reference's
# own output (tests/golden/golden_fullsize.json, made by tests/golden/make_golden_fullsize.py)
DIGEST_ONLY_CASES = {"soup_c5": soup_c5}

"""Deterministic synthetic inputs for the five configs and the host-side frame
driver that mirrors the draw section of the reference's main() (main.cpp:606-730).

The reference's assets (african_head, diablo3_pose, sponza) and its Assimp loader are absent
(SURVEY F4), so configs run on stand-in meshes + procedural TGA-style textures.  Everything here
is plain numpy on the host; all values handed to a backend are fp32-representable, like
Assimp's output (model.cpp:160-185).  The same arrays go to the CUDA backend and to the CPU
oracle, which is what makes the parity tests meaningful.
"""
import ctypes as C
import math
import os

import numpy as np

from .capi import (PhongUniforms, ShadowUniforms, SHADER_DEPTH, SHADER_EYE, SHADER_FLAT_BARY, SHADER_GOURAUD, SHADER_PHONG,
                   SHADER_SHADOW_PHONG)

_PKG = os.path.dirname(os.path.abspath(__file__))
SCENEGEN_LIB = os.path.join(_PKG, "libtrb_scenegen.so")

# main.cpp:615-617 (normalised by the caller with the backend's own helper order: the
# reference normalises with `normalized(make_vec3(..))`, i.e. v / sqrt(dot(v,v)))
KEY_LIGHT = (1.0, 1.4, 1.0)
FILL_LIGHT = (-0.3, 0.5, 0.2)
RIM_LIGHT = (-1.0, 0.8, -1.5)


# numpy mirror of TrbPhongUniforms (include/trb.h) for bulk fills
_PHONG_DTYPE = np.dtype([("key", "<f8", 3), ("fill", "<f8", 3), ("rim", "<f8", 3), ("nms", "<f8"),
                         ("diffuse", "<u8"), ("normal", "<u8"), ("specular", "<u8")])
assert _PHONG_DTYPE.itemsize == C.sizeof(PhongUniforms)


def normalized(v):
    """normalized(), geometry.h:136-140, in the reference's operation order."""
    v = np.asarray(v, dtype=np.float64)
    s = 0.0
    for c in v:
        s = s + c * c
    n = math.sqrt(s)
    return v if n == 0 else v / n


# --------------------------------------------------------------------------------------------
# meshes
# --------------------------------------------------------------------------------------------
class MeshData:
    def __init__(self, pos, nrm, uv, idx, name="mesh"):
        self.pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        # nrm / uv / idx may be None: the accessors' fallbacks (0,0,1) / (0,0) and an implicit soup
        self.nrm = None if nrm is None else np.ascontiguousarray(nrm, dtype=np.float32).reshape(-1, 3)
        self.uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        self.idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
        self.name = name

    @property
    def ntris(self):
        return (self.pos.shape[0] if self.idx is None else self.idx.size) // 3

    @property
    def nverts(self):
        return self.pos.shape[0]


def uv_sphere(n_around=36, n_stacks=35, radius=1.0, center=(0.0, 0.0, 0.0), name="sphere"):
    """UV sphere, 2*n_around*n_stacks triangles (2520 for the C1 head stand-in), CCW seen
    from outside, uv in [0,1]."""
    i = np.arange(n_stacks + 1)
    j = np.arange(n_around + 1)
    theta = np.pi * i / n_stacks            # 0 = north pole
    phi = 2.0 * np.pi * j / n_around
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    n = np.stack([st * np.cos(phi)[None, :], np.broadcast_to(ct, (n_stacks + 1, n_around + 1)),
                  st * np.sin(phi)[None, :]], axis=-1)
    pos = n * radius + np.asarray(center, dtype=np.float64)
    uv = np.stack([np.broadcast_to(j[None, :] / n_around, (n_stacks + 1, n_around + 1)),
                   np.broadcast_to(1.0 - i[:, None] / n_stacks, (n_stacks + 1, n_around + 1))], axis=-1)
    a = (i[:-1, None] * (n_around + 1) + j[None, :-1]).reshape(-1)
    b = a + 1
    c = a + (n_around + 1)
    d = c + 1
    # outside-facing CCW: (a, b, c) with phi increasing to +x->+z ... checked by test_scenes
    idx = np.stack([a, b, c, b, d, c], axis=-1).reshape(-1)
    return MeshData(pos.reshape(-1, 3), n.reshape(-1, 3), uv.reshape(-1, 2), idx, name)


def _grid_face(origin, du, dv, nu, nv, normal):
    """(nu x nv) quads spanning origin + s*du + t*dv; CCW about `normal`."""
    s = np.arange(nu + 1) / nu
    t = np.arange(nv + 1) / nv
    S, T = np.meshgrid(s, t, indexing="xy")
    pos = (np.asarray(origin)[None, None, :] + S[..., None] * np.asarray(du)[None, None, :]
           + T[..., None] * np.asarray(dv)[None, None, :])
    uv = np.stack([S, T], axis=-1)
    nrm = np.broadcast_to(np.asarray(normal, dtype=np.float64), pos.shape)
    a = (np.arange(nv)[:, None] * (nu + 1) + np.arange(nu)[None, :]).reshape(-1)
    b, c = a + 1, a + (nu + 1)
    d = c + 1
    if np.dot(np.cross(du, dv), normal) > 0:
        idx = np.stack([a, b, d, a, d, c], axis=-1)
    else:
        idx = np.stack([a, d, b, a, c, d], axis=-1)
    return pos.reshape(-1, 3), nrm.reshape(-1, 3), uv.reshape(-1, 2), idx.reshape(-1)


def box_room(size=(40.0, 15.0, 20.0), scale=1.0 / 0.014, quads=((256, 128), (256, 64), (128, 128)),
             name="room"):
    """'sponza' stand-in: an axis-aligned room seen from inside, 262 144 triangles by default
    (floor/ceiling 256x128, long walls 256x64, short walls 128x128 quads), given in MODEL units
    so that main.cpp:507's scale(0.014) maps it to `size` world units; floor at y=0."""
    sx, sy, sz = (np.asarray(size) * scale)
    hx, hz = sx / 2, sz / 2
    faces = [
        ((-hx, 0, -hz), (sx, 0, 0), (0, 0, sz), quads[0], (0, 1, 0)),    # floor, normal up
        ((-hx, sy, -hz), (sx, 0, 0), (0, 0, sz), quads[0], (0, -1, 0)),  # ceiling
        ((-hx, 0, -hz), (sx, 0, 0), (0, sy, 0), quads[1], (0, 0, 1)),    # wall z=-hz
        ((-hx, 0, hz), (sx, 0, 0), (0, sy, 0), quads[1], (0, 0, -1)),    # wall z=+hz
        ((-hx, 0, -hz), (0, 0, sz), (0, sy, 0), quads[2], (1, 0, 0)),    # wall x=-hx
        ((hx, 0, -hz), (0, 0, sz), (0, sy, 0), quads[2], (-1, 0, 0)),    # wall x=+hx
    ]
    P, N, U, I = [], [], [], []
    base = 0
    for origin, du, dv, (nu, nv), nrm in faces:
        p, n, u, i = _grid_face(np.array(origin, float), np.array(du, float), np.array(dv, float), nu, nv,
                                np.array(nrm, float))
        P.append(p); N.append(n); U.append(u); I.append(i + base)
        base += p.shape[0]
    return MeshData(np.concatenate(P), np.concatenate(N), np.concatenate(U), np.concatenate(I), name)


def icosphere(level, radius=1.0, name="icosphere"):
    """Subdivided icosahedron: 20*4^level triangles, 10*4^level+2 vertices (level 10 = the 20 971 520
    triangle sphere of config 4).  Vertices are normalised in f64, then rounded to fp32."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t],
                  [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
                  [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
        key = lo * nv + hi
        uniq, inv = np.unique(key, return_inverse=True)
        mid = v[uniq // nv] + v[uniq % nv]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = f.shape[0]
        m01, m12, m20 = inv[:nf] + nv, inv[nf:2 * nf] + nv, inv[2 * nf:] + nv
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    uv = np.stack([0.5 + np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi), 0.5 + np.arcsin(np.clip(v[:, 1], -1, 1)) / np.pi],
                  axis=1)
    return MeshData(v * radius, v, uv, f.reshape(-1), name)


# --------------------------------------------------------------------------------------------
# K7 / config-5 triangle soup through std::mt19937_64
# --------------------------------------------------------------------------------------------
def _scenegen():
    lib = C.CDLL(SCENEGEN_LIB)
    lib.trb_gen_soup_clip.restype = C.c_int
    lib.trb_gen_soup_clip.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p]
    return lib


def triangle_soup(ntris, width, height, r, seed, round_fp32, want_clip=True):
    """K7 generator: per triangle cx=U*W, cy=U*H, z=2U-1 from uniform_real_distribution<double>
    (0,1) on mt19937_64(seed); verts (cx-r,cy-r),(cx+r,cy-r),(cx,cy+r); NDC = p/(W/2)-1, w=1.
    Returns (clip f64 [n,3,4], pos f32 [3n,3]); with round_fp32 every coordinate is rounded to
    fp32 first (config 5) so both forms describe the same triangles."""
    clip = np.empty((ntris, 3, 4), dtype=np.float64) if want_clip else None
    pos = np.empty((ntris * 3, 3), dtype=np.float32)
    rc = _scenegen().trb_gen_soup_clip(seed, ntris, width, height, r, 1 if round_fp32 else 0,
                                       clip.ctypes.data_as(C.c_void_p) if want_clip else None,
                                       pos.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("trb_gen_soup_clip failed")
    return clip, pos


# --------------------------------------------------------------------------------------------
# procedural textures (TGAImage memory order: [y][x][BGR])
# --------------------------------------------------------------------------------------------
def _value_noise(size, cells, seed, channels):
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.random((cells + 1, cells + 1, channels))
    g[-1] = g[0]
    g[:, -1] = g[:, 0]
    t = np.arange(size) * (cells / size)
    i0 = np.floor(t).astype(np.int64)
    f = t - i0
    f = f * f * (3 - 2 * f)
    a = g[i0][:, i0] * (1 - f)[None, :, None] + g[i0][:, i0 + 1] * f[None, :, None]
    b = g[i0 + 1][:, i0] * (1 - f)[None, :, None] + g[i0 + 1][:, i0 + 1] * f[None, :, None]
    return a * (1 - f)[:, None, None] + b * f[:, None, None]


def texture_diffuse(size=1024, seed=11):
    n = 0.6 * _value_noise(size, 16, seed, 3) + 0.4 * _value_noise(size, 64, seed + 100, 3)
    return np.ascontiguousarray(np.clip(40 + 200 * n, 0, 255).astype(np.uint8))


def texture_normal(size=1024, seed=12, strength=6.0):
    hgt = (0.7 * _value_noise(size, 16, seed, 1) + 0.3 * _value_noise(size, 64, seed + 100, 1))[:, :, 0]
    gx = (np.roll(hgt, -1, axis=1) - np.roll(hgt, 1, axis=1)) * strength * size / 64
    gy = (np.roll(hgt, -1, axis=0) - np.roll(hgt, 1, axis=0)) * strength * size / 64
    n = np.stack([-gx, -gy, np.ones_like(gx)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    rgb = np.clip((n + 1) * 0.5 * 255, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(rgb[:, :, ::-1])  # stored BGR; Model::normal reads c[2] as x


def texture_specular(size=1024, seed=13):
    n = _value_noise(size, 32, seed, 1)
    g = np.clip(255 * n, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(np.repeat(g, 3, axis=2))


# --------------------------------------------------------------------------------------------
# matrices of main.cpp:365-420 (host scalar glue, plain numpy: inputs, not the path under test)
# --------------------------------------------------------------------------------------------
def scale_matrix(s):
    m = np.eye(4)
    m[0, 0] = m[1, 1] = m[2, 2] = s
    return m


def translation_matrix(tx, ty, tz):
    m = np.eye(4)
    m[0, 3], m[1, 3], m[2, 3] = tx, ty, tz
    return m


def rotation_y_matrix(angle_rad):
    m = np.eye(4)
    c, s = math.cos(angle_rad), math.sin(angle_rad)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


# --------------------------------------------------------------------------------------------
# scenes
# --------------------------------------------------------------------------------------------
class DrawItem:
    """One model of the frame: what main.cpp:647-669 sets up around a face loop."""

    def __init__(self, mesh, model_matrix, kind, textures=None, normal_map_strength=1.0,
                 snapshot_before=False, restore_after=False, cull_with=None):
        self.mesh = mesh
        self.model_matrix = np.asarray(model_matrix, dtype=np.float64)
        self.kind = kind
        self.textures = textures or {}
        self.normal_map_strength = normal_map_strength
        self.snapshot_before = snapshot_before
        self.restore_after = restore_after
        self.cull_with = cull_with     # index of the item whose box decides this item's visibility (None: its own)


class Scene:
    def __init__(self, name, width, height, items, fov, znear, zfar):
        self.name = name
        self.width, self.height = width, height
        self.items = items
        self.fov, self.znear, self.zfar = fov, znear, zfar

    @property
    def ntris(self):
        return sum(it.mesh.ntris for it in self.items)


# --------------------------------------------------------------------------------------------
# model-level frustum culling, bug-for-bug (our_gl.cpp:212-280, geometry.h:297-327, model.cpp:15-40):
# host scalars, plain Python floats in the reference's operation order
# --------------------------------------------------------------------------------------------
def local_aabb(mesh):
    """Model::computeAABB: min / max over the vertices widened by 1 % of the extent on every side"""
    pos = np.asarray(mesh.pos, dtype=np.float64)
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    margin = (hi - lo) * 0.01
    return lo - margin, hi + margin


def visible_items(scene, views, perspective, api):
    """Which models main() would draw for each camera (main.cpp:623-624, 647, 680, 706): bool array (items, views),
    through the backend's host helpers (trb_aabb_transform / trb_cull_batch: the reference's Frustum, bug-for-bug).
    An item with `cull_with` set is tested with THAT item's box and only drawn when that item is: main.cpp:706 tests
    the head's box for the eyes (sic) inside the head's own block."""
    out = []
    for it in scene.items:
        src = scene.items[it.cull_with] if it.cull_with is not None else it
        if getattr(src, "_world_aabb", None) is None:
            lo, hi = local_aabb(src.mesh)
            src._world_aabb = api.aabb_transform(lo, hi, src.model_matrix)     # Model::getWorldAABB, model.h:99
        out.append(api.cull_batch(perspective, views, *src._world_aabb))
    return np.stack(out)


class UploadedScene:
    """Meshes and textures of a Scene resident in one backend context (upload once)."""

    def __init__(self, renderer, scene):
        self.r = renderer
        self.scene = scene
        self.mesh_h = {}
        self.tex_h = {}
        self.h2d_bytes = 0
        for it in scene.items:
            if id(it.mesh) not in self.mesh_h:
                m = it.mesh
                self.mesh_h[id(m)] = renderer.upload_mesh(m.pos, m.nrm, m.uv, m.idx)
                self.h2d_bytes += sum(a.nbytes for a in (m.pos, m.nrm, m.uv, m.idx) if a is not None)
            for t in it.textures.values():
                if id(t) not in self.tex_h:
                    self.tex_h[id(t)] = renderer.upload_texture(t)
                    self.h2d_bytes += t.nbytes

    def free(self):
        for h in self.mesh_h.values():
            self.r.free_mesh(h)
        for h in self.tex_h.values():
            self.r.free_texture(h)
        self.mesh_h, self.tex_h = {}, {}

    def _item_params(self, it, views, seen_i):
        """ModelViews and the uniform blocks of one model for every camera: main.cpp:653 (ModelView = view * model),
        main.cpp:55-69 (light directions through the upper-left 3x3)"""
        api = self.r.api
        n = views.shape[0]
        key, fill, rim = normalized(KEY_LIGHT), normalized(FILL_LIGHT), normalized(RIM_LIGHT)
        mvs = api.mat4_mul_batch(views, it.model_matrix)
        if seen_i is not None and not seen_i.all():
            # a batch draws into every frame: cameras that cull the model get a ModelView of zeros, which sends
            # every vertex to w = 0 and every triangle to the reject of our_gl.cpp:94 - nothing is drawn there
            mvs = np.where(seen_i[:, None, None], mvs, 0.0)
        uni = None
        if it.kind in (SHADER_PHONG, SHADER_EYE):
            # PhongUniforms for every view, filled through a numpy view of the ctypes array
            arr = (PhongUniforms * n)()
            rec = np.frombuffer(arr, dtype=_PHONG_DTYPE)
            rec["key"] = api.light_dir_eye_batch(mvs, key)
            rec["fill"] = api.light_dir_eye_batch(mvs, fill)
            rec["rim"] = api.light_dir_eye_batch(mvs, rim)
            rec["nms"] = it.normal_map_strength
            rec["diffuse"] = self.tex_h.get(id(it.textures.get("diffuse")), 0)
            rec["normal"] = self.tex_h.get(id(it.textures.get("normal")), 0)
            rec["specular"] = self.tex_h.get(id(it.textures.get("specular")), 0)
            uni = arr
        return mvs, uni

    def render(self, views, perspective, cull=True):
        """views: (n,4,4) view matrices.  Mirrors main.cpp:606-730 for every view: begin frame, the model-level
        frustum test (main.cpp:623-624, 647, 680, 706; bug-for-bug), per model ModelView = view*model
        (main.cpp:653), light directions through the upper-left 3x3 (main.cpp:55-69), face loop -> one draw
        call, z snapshot/restore around the eyes (main.cpp:700,730)."""
        r, sc, api = self.r, self.scene, self.r.api
        views = np.asarray(views, dtype=np.float64).reshape(-1, 4, 4)
        n = views.shape[0]
        r.begin_frame(sc.width, sc.height, nviews=n)
        seen = visible_items(sc, views, perspective, api) if cull else None
        self.culled = 0 if seen is None else int((~seen).sum())
        self.drawn_items = []
        for i, it in enumerate(sc.items):
            if seen is not None and not seen[i].any():
                continue                                  # culled for every camera of the batch: the block is skipped
            if it.snapshot_before:
                r.depth_snapshot()
            mvs, uni = self._item_params(it, views, None if seen is None else seen[i])
            r.draw(self.mesh_h[id(it.mesh)], mvs, perspective, kind=it.kind, uniforms=uni,
                   ntris=it.mesh.ntris)
            self.drawn_items.append(i)
            if it.restore_after:
                r.depth_restore()
        r.end_frame()

    def record(self, views, perspective, cull=True):
        """The frame loop body above as a frame recording (trb_record_begin / trb_record_end: one CUDA graph launch per
        frame instead of ~40 kernel launches).  The frame is rendered once by plain calls first, so that every buffer
        has its size.  Returns the recording; the context holds the rendered frame."""
        self.render(views, perspective, cull)
        self.r.record_begin()
        try:
            self.render(views, perspective, cull)
        finally:
            rec = self.r.record_end()
        self.recorded_items = list(self.drawn_items)
        return rec

    def replay(self, recording, views, perspective, cull=True):
        """The recorded frame for new cameras: same draw calls, new ModelViews / light directions.  A model the new
        camera culls but the recorded one drew gets a ModelView of zeros (nothing is drawn); a model the recorded camera
        culled cannot come back - record with cull=False if the orbit needs every model."""
        sc, api = self.scene, self.r.api
        views = np.asarray(views, dtype=np.float64).reshape(-1, 4, 4)
        seen = visible_items(sc, views, perspective, api) if cull else None
        draws = []
        for i in self.recorded_items:
            it = sc.items[i]
            mvs, uni = self._item_params(it, views, None if seen is None else seen[i])
            draws.append({"modelview": mvs, "perspective": perspective, "uniforms": uni})
        self.r.replay(recording, draws)


def head_scene(width=800, height=800, n_around=36, n_stacks=35, tex_size=1024, name="c1_head"):
    """Config 1 stand-in: 2520-triangle UV sphere, three procedural 1024^2 maps, PhongShader
    with normal_map_strength 1.0, lookat((1,1,3),(0,0,0),(0,1,0)), fov 60."""
    mesh = uv_sphere(n_around, n_stacks, 1.0, name="head")
    tex = {"diffuse": texture_diffuse(tex_size, 11), "normal": texture_normal(tex_size, 12),
           "specular": texture_specular(tex_size, 13)}
    item = DrawItem(mesh, np.eye(4), SHADER_PHONG, tex, 1.0)
    return Scene(name, width, height, [item], 60.0, 0.1, 100.0)


def head_view(api):
    return api.lookat([1.0, 1.0, 3.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])


ORBIT_TARGET = (1.3555, 1.5116, -0.9686)  # main.cpp:588


def orbit_scene(width=1920, height=1080, room_quads=((256, 128), (256, 64), (128, 128)), head_res=(36, 35),
                eye_res=(16, 12), tex_size=1024, name="c3_orbit"):
    """Config 3 stand-in for main.cpp:483-489 + 506-513: room ('sponza', scale 0.014, Phong with
    normal_map_strength 0.5), head (T(0,1.6815,0)*Ry(-112.82 deg), Phong 1.0), eyes (same matrix,
    EyeShader, z snapshot before / restore after as main.cpp:700,730)."""
    room = box_room(quads=room_quads)
    head = uv_sphere(head_res[0], head_res[1], 1.0, name="head")
    e1 = uv_sphere(eye_res[0], eye_res[1], 0.12, (0.3, 0.2, 0.85))
    e2 = uv_sphere(eye_res[0], eye_res[1], 0.12, (-0.3, 0.2, 0.85))
    eyes = MeshData(np.concatenate([e1.pos, e2.pos]), np.concatenate([e1.nrm, e2.nrm]),
                    np.concatenate([e1.uv, e2.uv]), np.concatenate([e1.idx, e2.idx + e1.nverts]), "eyes")
    head_m = translation_matrix(0.0, 1.6815, 0.0) @ rotation_y_matrix(-112.82 * math.pi / 180.0)
    t_room = {"diffuse": texture_diffuse(tex_size, 21), "normal": texture_normal(tex_size, 22),
              "specular": texture_specular(tex_size, 23)}
    t_head = {"diffuse": texture_diffuse(tex_size, 11), "normal": texture_normal(tex_size, 12),
              "specular": texture_specular(tex_size, 13)}
    t_eye = {"diffuse": texture_diffuse(256, 31)}
    items = [DrawItem(room, scale_matrix(0.014), SHADER_PHONG, t_room, 0.5),
             DrawItem(head, head_m, SHADER_PHONG, t_head, 1.0),
             DrawItem(eyes, head_m, SHADER_EYE, t_eye, 1.0, snapshot_before=True, restore_after=True, cull_with=1)]
    return Scene(name, width, height, items, 70.0, 0.05, 500.0)


def orbit_views(api, frames, total=1024):
    """eye = target + 4*(cos t, 0.17, sin t), t = 2*pi*k/total; frames: iterable of k"""
    out = []
    tgt = np.asarray(ORBIT_TARGET)
    for k in frames:
        th = 2.0 * math.pi * k / total
        eye = tgt + 4.0 * np.array([math.cos(th), 0.17, math.sin(th)])
        out.append(api.lookat(eye, tgt, [0.0, 1.0, 0.0]))
    return np.stack(out)


def sphere_scene(level=10, width=3840, height=2160, name="c4_sphere"):
    """Config 4: icosphere level 10, camera (0,0,2.2), fov 60, near 0.1, far 10, flat shader."""
    mesh = icosphere(level)
    return Scene(name, width, height, [DrawItem(mesh, np.eye(4), SHADER_FLAT_BARY)], 60.0, 0.1, 10.0)


def sphere_view(api):
    return api.lookat([0.0, 0.0, 2.2], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])


# --------------------------------------------------------------------------------------------
# config 2: two-pass shadow-mapped, normal-mapped shader (shaders authored in
# the test oracle (ref_harness.cpp) because the fork has none)
# --------------------------------------------------------------------------------------------
LIGHT_POS = tuple(5.0 * c for c in normalized((1.0, 1.4, 1.0)))


def shadow_scene(width=2048, height=2048, body_res=(42, 60), ground_quads=32, tex_size=1024, name="c2_shadow"):
    """'diablo3_pose' stand-in: a 5040-triangle sphere standing on a subdivided ground plane that
    receives its shadow; both drawn with SHADOW_PHONG."""
    body = uv_sphere(body_res[0], body_res[1], 1.0, name="body")
    p, n, u, i = _grid_face(np.array([-4.0, -1.0, 4.0]), np.array([8.0, 0.0, 0.0]), np.array([0.0, 0.0, -8.0]),
                            ground_quads, ground_quads, np.array([0.0, 1.0, 0.0]))
    ground = MeshData(p, n, u, i, "ground")
    t_body = {"diffuse": texture_diffuse(tex_size, 41), "normal": texture_normal(tex_size, 42),
              "specular": texture_specular(tex_size, 43)}
    t_ground = {"diffuse": texture_diffuse(min(tex_size, 256), 44)}
    items = [DrawItem(ground, np.eye(4), SHADER_SHADOW_PHONG, t_ground, 0.0),
             DrawItem(body, np.eye(4), SHADER_SHADOW_PHONG, t_body, 1.0)]
    return Scene(name, width, height, items, 60.0, 0.1, 100.0)


def render_shadowed(up, view, perspective, shadow_size=None, bias=2e-3, darkening=0.35, kind=SHADER_SHADOW_PHONG, release=True):
    """Pass 1: depth-only draw of every model from the light (TRB_SHADER_DEPTH) kept as the shadow map.
    Pass 2: the camera frame with `kind` (SHADOW_PHONG, or GOURAUD / PHONG for the plain variants)."""
    r, sc, api = up.r, up.scene, up.r.api
    sw, sh = shadow_size or (sc.width, sc.height)
    light_view = api.lookat(LIGHT_POS, [0.0, 0.0, 0.0], [0.0, 1.0, 0.0])
    light_proj = api.perspective(40.0, sw / sh, 1.0, 20.0)
    light_vp = api.viewport(0, 0, sw, sh)
    if release:            # (a frame recording releases before it begins: trb_release_shadow_maps is refused inside)
        r.release_shadow_maps()
    r.begin_frame(sw, sh)
    for it in sc.items:
        r.draw(up.mesh_h[id(it.mesh)], api.mat4_mul(light_view, it.model_matrix), light_proj, kind=SHADER_DEPTH,
               ntris=it.mesh.ntris)
    r.end_frame()
    smap = r.keep_depth_as_shadow_map()
    key, fill, rim = normalized(KEY_LIGHT), normalized(FILL_LIGHT), normalized(RIM_LIGHT)
    r.begin_frame(sc.width, sc.height)
    for it in sc.items:
        mv = api.mat4_mul(view, it.model_matrix)
        ph = PhongUniforms()
        ph.key_dir_eye[:] = api.light_dir_eye(mv, key)
        ph.fill_dir_eye[:] = api.light_dir_eye(mv, fill)
        ph.rim_dir_eye[:] = api.light_dir_eye(mv, rim)
        ph.normal_map_strength = it.normal_map_strength
        ph.diffuse = up.tex_h.get(id(it.textures.get("diffuse")), 0)
        ph.normal = up.tex_h.get(id(it.textures.get("normal")), 0)
        ph.specular = up.tex_h.get(id(it.textures.get("specular")), 0)
        if kind == SHADER_SHADOW_PHONG:
            u = ShadowUniforms()
            u.phong = ph
            u.light_modelview[:] = api.mat4_mul(light_view, it.model_matrix).reshape(-1)
            u.light_perspective[:] = np.asarray(light_proj).reshape(-1)
            u.light_viewport[:] = np.asarray(light_vp).reshape(-1)
            u.shadow_bias = bias
            u.shadow_darkening = darkening
            u.shadow_map = smap
            u.shadow_w, u.shadow_h = sw, sh
        else:
            u = ph
        r.draw(up.mesh_h[id(it.mesh)], mv, perspective, kind=kind, uniforms=u, ntris=it.mesh.ntris)
    r.end_frame()

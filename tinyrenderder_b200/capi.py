"""ctypes binding of the C ABI declared in include/trb.h.

`Api(path, prefix)` binds one shared library exporting `<prefix>_*` with the signatures of
include/trb.h.  The product uses `load_cuda()` (libtrb.so, prefix ``trb``); the test-suite and
bench.py's cpu_baseline leg bind the CPU oracle (prefix ``orc``) through the same class, which
is what makes the parity tests read the same for both sides.  Nothing in this package imports
or falls back to the oracle: if libtrb.so is missing, `load_cuda()` raises.
"""
import contextlib
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB = os.path.join(_PKG, "libtrb.so")

SHADER_FLAT_BARY = 0
SHADER_PHONG = 1
SHADER_EYE = 2
SHADER_DEPTH = 3
SHADER_SHADOW_PHONG = 4
SHADER_GOURAUD = 5

IMAGE_COLOR = 0
IMAGE_DEPTH = 1
IMAGE_SSAO = 2
IMAGE_FINAL = 3

VIS_NONE = 0xFFFFFFFF
VIS_SHADED = 0


class TrbError(RuntimeError):
    pass


class PhongUniforms(C.Structure):
    _fields_ = [
        ("key_dir_eye", C.c_double * 3),
        ("fill_dir_eye", C.c_double * 3),
        ("rim_dir_eye", C.c_double * 3),
        ("normal_map_strength", C.c_double),
        ("diffuse", C.c_uint64),
        ("normal", C.c_uint64),
        ("specular", C.c_uint64),
    ]


class ShadowUniforms(C.Structure):
    _fields_ = [
        ("phong", PhongUniforms),
        ("light_modelview", C.c_double * 16),
        ("light_perspective", C.c_double * 16),
        ("light_viewport", C.c_double * 16),
        ("shadow_bias", C.c_double),
        ("shadow_darkening", C.c_double),
        ("shadow_map", C.c_int32),
        ("shadow_w", C.c_int32),
        ("shadow_h", C.c_int32),
        ("_pad", C.c_int32),
    ]


class ReplayDraw(C.Structure):
    _fields_ = [
        ("modelview", C.c_void_p),
        ("perspective", C.c_void_p),
        ("uniforms", C.c_void_p),
        ("uniform_bytes", C.c_size_t),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("triangles_submitted", C.c_uint64),
        ("triangles_binned", C.c_uint64),
        ("tile_entries", C.c_uint64),
        ("fragments_covered", C.c_uint64),
        ("pixels_shaded", C.c_uint64),
        ("visible_triangles", C.c_uint64),
        ("bbox_min_x", C.c_int32),
        ("bbox_min_y", C.c_int32),
        ("bbox_max_x", C.c_int32),
        ("bbox_max_y", C.c_int32),
        ("z_min", C.c_double),
        ("z_max_covered", C.c_double),
        ("fragments_drawn_ref", C.c_uint64),
        ("z_max_ref", C.c_double),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_uint64), ("ms", C.c_double)]


_P = C.c_void_p
_D = C.POINTER(C.c_double)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against trb.h
SIGNATURES = {
    "create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "destroy": (C.c_int, [_P]),
    "last_error": (C.c_char_p, [_P]),
    "backend_name": (C.c_char_p, []),
    "upload_mesh": (C.c_int, [_P, _P, _P, _P, C.c_uint32, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "free_mesh": (C.c_int, [_P, C.c_uint64]),
    "upload_texture": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "free_texture": (C.c_int, [_P, C.c_uint64]),
    "begin_frame": (C.c_int, [_P, C.c_int, C.c_int]),
    "begin_batch": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "set_clear_color": (C.c_int, [_P, C.c_uint8, C.c_uint8, C.c_uint8]),
    "set_viewport": (C.c_int, [_P, _P]),
    "draw": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_int, _P, C.c_size_t, C.c_uint64, C.c_uint64]),
    "draw_batch": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_int, _P, C.c_size_t, C.c_uint64, C.c_uint64]),
    "draw_shard": (C.c_int, [_P, C.c_uint64, _P, _P, C.c_int, _P, C.c_size_t, C.c_int, C.c_int]),
    "submit_clip_triangles": (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_int, _P, C.c_size_t]),
    "depth_snapshot": (C.c_int, [_P]),
    "depth_restore": (C.c_int, [_P]),
    "keep_depth_as_shadow_map": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "release_shadow_maps": (C.c_int, [_P]),
    "flush": (C.c_int, [_P]),
    "end_frame": (C.c_int, [_P]),
    "record_begin": (C.c_int, [_P]),
    "record_end": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "replay": (C.c_int, [_P, C.c_uint64, _P, C.c_int]),
    "recording_free": (C.c_int, [_P, C.c_uint64]),
    "ssao": (C.c_int, [_P, C.c_int, _P]),
    "depth_image": (C.c_int, [_P, C.c_int, _P]),
    "composite_ao": (C.c_int, [_P, C.c_int, _P]),
    "encode_tga": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P]),
    "encode_tga_async": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P]),
    "read_color": (C.c_int, [_P, C.c_int, _P]),
    "write_color": (C.c_int, [_P, C.c_int, _P]),
    "read_depth": (C.c_int, [_P, C.c_int, _P]),
    "read_visibility": (C.c_int, [_P, C.c_int, _P]),
    "readback_async": (C.c_int, [_P, _P, _P]),
    "readback_wait": (C.c_int, [_P]),
    "get_stats": (C.c_int, [_P, C.c_int, C.POINTER(Stats)]),
    "synchronize": (C.c_int, [_P]),
    "timer_start": (C.c_int, [_P]),
    "timer_stop_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "profile_enable": (C.c_int, [_P, C.c_int]),
    "profile_read": (C.c_int, [_P, C.POINTER(KernelTime), C.c_int, C.POINTER(C.c_int), C.c_int]),
    "launch_count": (C.c_uint64, [_P]),
    "device_planes": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "set_triangle_id_base": (C.c_int, [_P, C.c_uint64]),
    "composite_save_local_depth": (C.c_int, [_P]),
    "composite_mask": (C.c_int, [_P]),
    "composite_finish": (C.c_int, [_P]),
    "set_shade_rows": (C.c_int, [_P, C.c_int, C.c_int]),
    "ipc_export_planes": (C.c_int, [_P, _P, _P]),
    "ipc_open_peers": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "open_peers_raw": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "ipc_close_peers": (C.c_int, [_P]),
    "composite_shade_p2p": (C.c_int, [_P, C.c_int, C.c_int]),
    "comm_init": (C.c_int, [_P, C.c_int]),
    "comm_export": (C.c_int, [_P, _P, C.c_size_t]),
    "comm_open": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "comm_close": (C.c_int, [_P]),
    "comm_shard": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "comm_rows": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "composite": (C.c_int, [_P]),
    "composite_group": (C.c_int, [_P, C.c_int]),
    "light_dir_eye": (None, [_P, _P, _P]),
    "lookat": (None, [_P, _P, _P, _P]),
    "perspective": (None, [C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "viewport": (None, [C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mat4_mul": (None, [_P, _P, _P]),
    "mat4_mul_batch": (None, [_P, C.c_int, _P, _P]),
    "light_dir_eye_batch": (None, [_P, C.c_int, _P, _P]),
    "frustum_planes": (None, [_P, _P]),
    "frustum_intersects": (C.c_int, [_P, _P, _P]),
    "aabb_transform": (None, [_P, _P, _P, _P, _P]),
    "cull_batch": (None, [_P, _P, C.c_int, _P, _P, _P]),
}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("expected %d doubles, got %d" % (n, a.size))
    return a


class Api:
    """One loaded backend library (CUDA product or CPU oracle)."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise TrbError("backend library not found: %s" % path)
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        self.fn = {}
        for name, (res, args) in SIGNATURES.items():
            f = getattr(self.lib, "%s_%s" % (prefix, name))
            f.restype = res
            f.argtypes = args
            self.fn[name] = f

    def backend_name(self):
        return self.fn["backend_name"]().decode()

    # host helpers (reference operation order) ------------------------------------------
    def lookat(self, eye, center, up):
        out = np.empty(16)
        self.fn["lookat"](_ptr(_f64(eye, 3)), _ptr(_f64(center, 3)), _ptr(_f64(up, 3)), _ptr(out))
        return out.reshape(4, 4)

    def perspective(self, fov_deg, aspect, znear, zfar):
        out = np.empty(16)
        self.fn["perspective"](fov_deg, aspect, znear, zfar, _ptr(out))
        return out.reshape(4, 4)

    def viewport(self, x, y, w, h):
        out = np.empty(16)
        self.fn["viewport"](x, y, w, h, _ptr(out))
        return out.reshape(4, 4)

    def mat4_mul(self, a, b):
        out = np.empty(16)
        self.fn["mat4_mul"](_ptr(_f64(a, 16)), _ptr(_f64(b, 16)), _ptr(out))
        return out.reshape(4, 4)

    def mat4_mul_batch(self, a, b):
        a = _f64(a)
        n = a.size // 16
        out = np.empty((n, 4, 4))
        self.fn["mat4_mul_batch"](_ptr(a), n, _ptr(_f64(b, 16)), _ptr(out))
        return out

    def light_dir_eye_batch(self, modelviews, dir_world):
        mv = _f64(modelviews)
        n = mv.size // 16
        out = np.empty((n, 3))
        self.fn["light_dir_eye_batch"](_ptr(mv), n, _ptr(_f64(dir_world, 3)), _ptr(out))
        return out

    # model-level frustum culling, bug-for-bug (our_gl.cpp:212-280, geometry.h:297-327) ----
    def frustum_planes(self, view_projection):
        out = np.empty(24)
        self.fn["frustum_planes"](_ptr(_f64(view_projection, 16)), _ptr(out))
        return out.reshape(6, 4)

    def frustum_intersects(self, planes, box_min, box_max):
        return bool(self.fn["frustum_intersects"](_ptr(_f64(planes, 24)), _ptr(_f64(box_min, 3)), _ptr(_f64(box_max, 3))))

    def aabb_transform(self, box_min, box_max, m):
        lo, hi = np.empty(3), np.empty(3)
        self.fn["aabb_transform"](_ptr(_f64(box_min, 3)), _ptr(_f64(box_max, 3)), _ptr(_f64(m, 16)), _ptr(lo), _ptr(hi))
        return lo, hi

    def cull_batch(self, perspective, views, box_min, box_max):
        v = _f64(views)
        n = v.size // 16
        out = np.empty(n, dtype=np.uint8)
        self.fn["cull_batch"](_ptr(_f64(perspective, 16)), _ptr(v), n, _ptr(_f64(box_min, 3)), _ptr(_f64(box_max, 3)), _ptr(out))
        return out.astype(bool)

    def light_dir_eye(self, modelview, dir_world):
        out = np.empty(3)
        self.fn["light_dir_eye"](_ptr(_f64(modelview, 16)), _ptr(_f64(dir_world, 3)), _ptr(out))
        return out


class Renderer:
    """A context of one backend; thin, 1:1 over the C ABI."""

    def __init__(self, api, device=0):
        self.api = api
        self._fn = api.fn
        h = _P()
        rc = self._fn["create"](device, C.byref(h))
        if rc != 0 or not h:
            raise TrbError("%s_create(device=%d) failed: rc=%d (no CPU fallback)" % (api.prefix, device, rc))
        self.h = h
        self.width = self.height = 0
        self.nviews = 0
        self._keep = []
        self._collect = None

    def close(self):
        if self.h:
            self._fn["destroy"](self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            msg = self._fn["last_error"](self.h)
            raise TrbError("%s_%s -> %d: %s" % (self.api.prefix, what, rc, msg.decode() if msg else ""))

    # resources ---------------------------------------------------------------------------
    def upload_mesh(self, pos, nrm=None, uv=None, idx=None):
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        nv = pos.shape[0]
        nrm = None if nrm is None else np.ascontiguousarray(nrm, dtype=np.float32).reshape(nv, 3)
        uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float32).reshape(nv, 2)
        if idx is None:
            nidx = nv
        else:
            idx = np.ascontiguousarray(idx, dtype=np.uint32).reshape(-1)
            nidx = idx.size
        out = C.c_uint64(0)
        self._ck(self._fn["upload_mesh"](self.h, _ptr(pos), _ptr(nrm), _ptr(uv), nv, _ptr(idx), nidx,
                                         C.byref(out)), "upload_mesh")
        return out.value

    def free_mesh(self, mesh):
        self._ck(self._fn["free_mesh"](self.h, mesh), "free_mesh")

    def upload_texture(self, texels):
        t = np.ascontiguousarray(texels, dtype=np.uint8)
        if t.ndim == 2:
            t = t[:, :, None]
        h, w, bpp = t.shape
        out = C.c_uint64(0)
        self._ck(self._fn["upload_texture"](self.h, _ptr(t), w, h, bpp, C.byref(out)), "upload_texture")
        return out.value

    def free_texture(self, tex):
        self._ck(self._fn["free_texture"](self.h, tex), "free_texture")

    # frame -------------------------------------------------------------------------------
    def begin_frame(self, w, h, nviews=1, viewport=None):
        if nviews == 1:
            self._ck(self._fn["begin_frame"](self.h, w, h), "begin_frame")
        else:
            self._ck(self._fn["begin_batch"](self.h, w, h, nviews), "begin_batch")
        self.width, self.height, self.nviews = w, h, nviews
        self.set_viewport(self.api.viewport(0, 0, w, h) if viewport is None else viewport)

    def set_clear_color(self, b, g, r):
        self._ck(self._fn["set_clear_color"](self.h, b, g, r), "set_clear_color")

    def set_viewport(self, m):
        self._ck(self._fn["set_viewport"](self.h, _ptr(_f64(m, 16))), "set_viewport")

    def draw(self, mesh, modelview, perspective, kind=SHADER_FLAT_BARY, uniforms=None, first_tri=0,
             ntris=None, mesh_ntris=None):
        """modelview/perspective: (4,4) or (nviews,4,4); uniforms: one struct or a ctypes array."""
        mv = _f64(modelview, 16 * self.nviews)
        pr = np.ascontiguousarray(perspective, dtype=np.float64)
        if pr.size == 16 and self.nviews > 1:
            pr = np.tile(pr.reshape(1, 16), (self.nviews, 1))
        pr = _f64(pr, 16 * self.nviews)
        if ntris is None:
            if mesh_ntris is None:
                raise ValueError("ntris required")
            ntris = mesh_ntris
        if uniforms is None:
            up, ub = None, 0
        else:
            if not isinstance(uniforms, C.Array):
                arr = (type(uniforms) * self.nviews)(*([uniforms] * self.nviews))
            else:
                arr = uniforms
            up, ub = C.cast(arr, _P), C.sizeof(arr._type_)
            self._keep.append(arr)
        if self._collect is not None:       # collect_draws(): the parameters a replay of this frame would pass
            self._collect.append({"modelview": mv, "perspective": pr, "uniforms": None if uniforms is None else arr})
            return
        name = "draw" if self.nviews == 1 else "draw_batch"
        self._ck(self._fn[name](self.h, mesh, _ptr(mv), _ptr(pr), kind, up, ub, first_tri, ntris), name)

    def draw_shard(self, mesh, modelview, perspective, shard_rank, shard_count, kind=SHADER_FLAT_BARY, uniforms=None):
        """this context's share of a mesh that `shard_count` contexts draw into one picture (trb_draw_shard)"""
        mv, pr = _f64(modelview, 16), _f64(perspective, 16)
        up, ub = (None, 0) if uniforms is None else (C.cast(C.pointer(uniforms), _P), C.sizeof(uniforms))
        if uniforms is not None:
            self._keep.append(uniforms)
        self._ck(self._fn["draw_shard"](self.h, mesh, _ptr(mv), _ptr(pr), kind, up, ub, shard_rank, shard_count), "draw_shard")

    def submit_clip_triangles(self, clip, varyings=None, modelview=None, kind=SHADER_FLAT_BARY, uniforms=None):
        clip = np.ascontiguousarray(clip, dtype=np.float64).reshape(-1, 12)
        n = clip.shape[0]
        vr = None if varyings is None else _f64(varyings, n * 24)
        mv = None if modelview is None else _f64(modelview, 16)
        up, ub = (None, 0) if uniforms is None else (C.cast(C.pointer(uniforms), _P), C.sizeof(uniforms))
        self._ck(self._fn["submit_clip_triangles"](self.h, _ptr(clip), _ptr(vr), n, _ptr(mv), kind, up, ub),
                 "submit_clip_triangles")

    def depth_snapshot(self):
        self._ck(self._fn["depth_snapshot"](self.h), "depth_snapshot")

    def depth_restore(self):
        self._ck(self._fn["depth_restore"](self.h), "depth_restore")

    def keep_depth_as_shadow_map(self):
        out = C.c_int32(-1)
        self._ck(self._fn["keep_depth_as_shadow_map"](self.h, C.byref(out)), "keep_depth_as_shadow_map")
        return out.value

    def release_shadow_maps(self):
        self._ck(self._fn["release_shadow_maps"](self.h), "release_shadow_maps")

    def flush(self):
        self._ck(self._fn["flush"](self.h), "flush")

    def end_frame(self):
        self._ck(self._fn["end_frame"](self.h), "end_frame")
        self._keep.clear()

    # frame recordings (CUDA graphs) --------------------------------------------------------
    @contextlib.contextmanager
    def collect_draws(self):
        """Run a frame's call sequence WITHOUT touching the device and collect what its draw calls pass: the `draws`
        list for replay() of a recording of the same sequence with other cameras / lights.
            with r.collect_draws() as draws: render_frame(r, new_camera)
            r.replay(rec, draws)
        Inside, begin_frame / snapshot / restore / flush / end_frame do nothing and keep_depth_as_shadow_map returns the
        index the recorded frame got (0, 1, ... after a release_shadow_maps)."""
        real, state = self._fn, (self.width, self.height, self.nviews)
        maps = [0]

        def stub(name):
            if name == "last_error":
                return real[name]
            if name == "keep_depth_as_shadow_map":
                def keep(h, out):
                    out._obj.value = maps[0]
                    maps[0] += 1
                    return 0
                return keep
            if name == "release_shadow_maps":
                def release(h):
                    maps[0] = 0
                    return 0
                return release
            return lambda *a: 0

        class Table(dict):
            def __missing__(self, name):
                self[name] = stub(name)
                return self[name]

        self._collect, self._fn = [], Table()
        try:
            yield self._collect
        finally:
            self._fn, self._collect = real, None
            self.width, self.height, self.nviews = state

    def record_begin(self):
        self._ck(self._fn["record_begin"](self.h), "record_begin")

    def record_end(self):
        out = C.c_uint64(0)
        self._ck(self._fn["record_end"](self.h, C.byref(out)), "record_end")
        self._keep.clear()
        return out.value

    def replay(self, recording, draws=None):
        """draws: None (replay unchanged) or one dict per recorded draw call with optional keys
        modelview / perspective ((nviews,4,4) or (4,4)) and uniforms (a struct or a ctypes array)"""
        if draws is None:
            self._ck(self._fn["replay"](self.h, recording, None, 0), "replay")
            return
        arr = (ReplayDraw * len(draws))()
        keep = []
        for i, d in enumerate(draws):
            d = d or {}
            for key in ("modelview", "perspective"):
                if d.get(key) is not None:
                    m = np.ascontiguousarray(d[key], dtype=np.float64)
                    if m.size == 16 and self.nviews > 1:
                        m = np.tile(m.reshape(1, 16), (self.nviews, 1))
                    m = _f64(m, 16 * self.nviews)
                    keep.append(m)
                    setattr(arr[i], key, m.ctypes.data)
            u = d.get("uniforms")
            if u is not None:
                if not isinstance(u, C.Array):
                    u = (type(u) * self.nviews)(*([u] * self.nviews))
                keep.append(u)
                arr[i].uniforms = C.cast(u, _P).value
                arr[i].uniform_bytes = C.sizeof(u._type_)
        self._ck(self._fn["replay"](self.h, recording, C.cast(arr, _P), len(draws)), "replay")

    def recording_free(self, recording):
        self._ck(self._fn["recording_free"](self.h, recording), "recording_free")

    # readback ----------------------------------------------------------------------------
    def read_color(self, view=0, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        self._ck(self._fn["read_color"](self.h, view, _ptr(out)), "read_color")
        return out

    def read_depth(self, view=0, out=None):
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.float64)
        self._ck(self._fn["read_depth"](self.h, view, _ptr(out)), "read_depth")
        return out

    def write_color(self, bgr, view=0):
        a = np.ascontiguousarray(bgr, dtype=np.uint8).reshape(self.height, self.width, 3)
        self._ck(self._fn["write_color"](self.h, view, _ptr(a)), "write_color")

    def readback_async(self, colors=None, depths=None):
        """queue the device->host copy of every view into the given lists of (pinned) numpy arrays"""
        def table(arrs):
            if arrs is None:
                return None
            t = (C.c_void_p * self.nviews)(*[a.ctypes.data for a in arrs])
            self._keep.append(t)
            return C.cast(t, _P)
        self._ck(self._fn["readback_async"](self.h, table(colors), table(depths)), "readback_async")

    def readback_wait(self):
        self._ck(self._fn["readback_wait"](self.h), "readback_wait")

    def read_visibility(self, view=0):
        out = np.empty((self.height, self.width), dtype=np.uint32)
        self._ck(self._fn["read_visibility"](self.h, view, _ptr(out)), "read_visibility")
        return out

    def ssao(self, view=0):
        out = np.empty((self.height, self.width), dtype=np.uint8)
        self._ck(self._fn["ssao"](self.h, view, _ptr(out)), "ssao")
        return out

    def encode_tga(self, which=0):
        """TGA file images (18-byte header + RLE packets, tgaimage.cpp:160-242) of every view: list of bytes"""
        n = self.nviews
        cap = self.width * self.height * 3 + self.width * self.height // 2 + 64
        bufs = [np.empty(cap, dtype=np.uint8) for _ in range(n)]
        table = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_uint64 * n)()
        self._ck(self._fn["encode_tga"](self.h, which, C.cast(table, _P), cap, C.cast(sizes, _P)), "encode_tga")
        return [bufs[v][:sizes[v]].tobytes() for v in range(n)]

    def encode_tga_async(self, bufs, sizes, which=0):
        """queue the TGA files of every view into the (pinned) uint8 arrays `bufs`; `sizes` is a uint64 array of
        nviews entries.  Both are valid after readback_wait()."""
        n = self.nviews
        table = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        self._keep.append(table)
        self._ck(self._fn["encode_tga_async"](self.h, which, C.cast(table, _P), bufs[0].size, _ptr(sizes)), "encode_tga_async")

    def depth_image(self, view=0):
        out = np.empty((self.height, self.width), dtype=np.uint8)
        self._ck(self._fn["depth_image"](self.h, view, _ptr(out)), "depth_image")
        return out

    def composite_ao(self, view=0):
        out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        self._ck(self._fn["composite_ao"](self.h, view, _ptr(out)), "composite_ao")
        return out

    def stats(self, view=0):
        s = Stats()
        self._ck(self._fn["get_stats"](self.h, view, C.byref(s)), "get_stats")
        return s.as_dict()

    def synchronize(self):
        self._ck(self._fn["synchronize"](self.h), "synchronize")

    # timing ------------------------------------------------------------------------------
    def timer_start(self):
        self._ck(self._fn["timer_start"](self.h), "timer_start")

    def timer_stop_ms(self):
        ms = C.c_float(0)
        self._ck(self._fn["timer_stop_ms"](self.h, C.byref(ms)), "timer_stop_ms")
        return ms.value

    def profile_enable(self, on=True):
        self._ck(self._fn["profile_enable"](self.h, 1 if on else 0), "profile_enable")

    def profile_read(self, reset=True):
        buf = (KernelTime * 64)()
        n = C.c_int(0)
        self._ck(self._fn["profile_read"](self.h, buf, 64, C.byref(n), 1 if reset else 0), "profile_read")
        return {buf[i].name.decode(): (buf[i].launches, buf[i].ms) for i in range(n.value)}

    def launch_count(self):
        return int(self._fn["launch_count"](self.h))

    # multi-GPU composite -------------------------------------------------------------------
    def device_planes(self):
        a, b, n = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._ck(self._fn["device_planes"](self.h, C.byref(a), C.byref(b), C.byref(n)), "device_planes")
        return a.value, b.value, n.value

    def set_triangle_id_base(self, base):
        self._ck(self._fn["set_triangle_id_base"](self.h, base), "set_triangle_id_base")

    def composite_save_local_depth(self):
        self._ck(self._fn["composite_save_local_depth"](self.h), "composite_save_local_depth")

    def composite_mask(self):
        self._ck(self._fn["composite_mask"](self.h), "composite_mask")

    def composite_finish(self):
        self._ck(self._fn["composite_finish"](self.h), "composite_finish")

    def ipc_export_planes(self):
        k, v = C.create_string_buffer(64), C.create_string_buffer(64)
        self._ck(self._fn["ipc_export_planes"](self.h, k, v), "ipc_export_planes")
        return k.raw, v.raw

    def ipc_open_peers(self, key_handles, vis_handles, my_rank):
        n = len(key_handles)
        self._ck(self._fn["ipc_open_peers"](self.h, b"".join(key_handles), b"".join(vis_handles), n, my_rank),
                 "ipc_open_peers")

    def open_peers_raw(self, key_ptrs, vis_ptrs, my_rank):
        n = len(key_ptrs)
        k = (C.c_uint64 * n)(*key_ptrs)
        v = (C.c_uint64 * n)(*vis_ptrs)
        self._ck(self._fn["open_peers_raw"](self.h, k, v, n, my_rank), "open_peers_raw")

    def ipc_close_peers(self):
        self._ck(self._fn["ipc_close_peers"](self.h), "ipc_close_peers")

    def composite_shade_p2p(self, y0, y1):
        self._ck(self._fn["composite_shade_p2p"](self.h, y0, y1), "composite_shade_p2p")

    # composite groups (trb_comm_*): no host barrier anywhere ---------------------------------
    COMM_BLOB_BYTES = 256

    def comm_export(self):
        b = C.create_string_buffer(self.COMM_BLOB_BYTES)
        self._ck(self._fn["comm_export"](self.h, b, self.COMM_BLOB_BYTES), "comm_export")
        return b.raw

    def comm_open(self, blobs, rank):
        self._ck(self._fn["comm_open"](self.h, b"".join(blobs), len(blobs), rank), "comm_open")

    def comm_close(self):
        self._ck(self._fn["comm_close"](self.h), "comm_close")

    def comm_shard(self, total):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._fn["comm_shard"](self.h, total, C.byref(a), C.byref(b)), "comm_shard")
        return a.value, b.value

    def comm_rows(self):
        a, b = C.c_int(0), C.c_int(0)
        self._ck(self._fn["comm_rows"](self.h, C.byref(a), C.byref(b)), "comm_rows")
        return a.value, b.value

    def composite(self):
        self._ck(self._fn["composite"](self.h), "composite")
        self._keep.clear()

    def set_shade_rows(self, y0, y1):
        self._ck(self._fn["set_shade_rows"](self.h, y0, y1), "set_shade_rows")


def comm_init(renderers):
    """trb_comm_init: the contexts of ONE process become the ranks of a composite group (rank = position)"""
    arr = (C.c_void_p * len(renderers))(*[r.h for r in renderers])
    rc = renderers[0]._fn["comm_init"](C.cast(arr, _P), len(renderers))
    if rc != 0:
        raise TrbError("trb_comm_init -> %d: %s" % (rc, " | ".join((r._fn["last_error"](r.h) or b"").decode() for r in renderers)))


def composite_group(renderers):
    """trb_composite_group: composite + shade the owned rows of every member, streams ordered by events"""
    arr = (C.c_void_p * len(renderers))(*[r.h for r in renderers])
    rc = renderers[0]._fn["composite_group"](C.cast(arr, _P), len(renderers))
    if rc != 0:
        raise TrbError("trb_composite_group -> %d: %s" % (rc, " | ".join((r._fn["last_error"](r.h) or b"").decode() for r in renderers)))
    for r in renderers:
        r._keep.clear()


_cuda_api = None


def load_cuda():
    """The product backend.  Raises (never falls back) when the CUDA library is not built."""
    global _cuda_api
    if _cuda_api is None:
        lib = os.environ.get("TRB_CUDA_LIB", CUDA_LIB)   # tuning builds of the same sources (profiles/ sweeps)
        if not os.path.exists(lib):
            raise TrbError("CUDA backend %s is not built; run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)" % lib)
        _cuda_api = Api(lib, "trb")
    return _cuda_api

"""The C++ host layer (tinyrenderder_b200/host: our_gl.h mirror, Model OBJ/TGA loader, ModelManager,
PhongShader/EyeShader) through the example program that mirrors the reference's main().

CPU: the ORACLE build of the example (reference headers + the reference's own our_gl.cpp/tgaimage.cpp)
must produce the z-buffer the Python-driven reference oracle produces for the same meshes, and our TGA
writer must be byte-identical to the reference's.  GPU: the DEVICE build of the same source must match the
oracle build's outputs."""
import os
import struct
import subprocess

import numpy as np
import pytest

import tinyrenderder_b200 as trb
from tinyrenderder_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "tinyrenderder_b200", "host", "bin", "example")
TGA_TOOL = os.path.join(ROOT, "tinyrenderder_b200", "host", "bin", "tga_tool")
EXAMPLE_REF = os.path.join(ROOT, "oracle", "_ref", "example_ref")
TGA_TOOL_REF = os.path.join(ROOT, "oracle", "_ref", "tga_tool_ref")
W, H = 300, 200


def write_obj(path, mesh):
    with open(path, "w") as f:
        for p in mesh.pos:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for t in mesh.uv:
            f.write("vt %.9g %.9g\n" % (t[0], np.float32(1.0) - t[1]))  # the loader flips v (aiProcess_FlipUVs)
        for n in mesh.nrm:
            f.write("vn %.9g %.9g %.9g\n" % tuple(n))
        for a, b, c in mesh.idx.reshape(-1, 3) + 1:
            f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, c, c, c))


def write_tga(path, bgr):
    h, w, bpp = bgr.shape
    with open(path, "wb") as f:  # type 2, top-left origin: read_tga_file keeps the rows as they are
        f.write(struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, w, h, bpp * 8, 0x20))
        f.write(np.ascontiguousarray(bgr).tobytes())


def read_tga(path):
    raw = open(path, "rb").read()
    idl, _, typ, _, _, _, _, _, w, h, bits, desc = struct.unpack("<BBBHHBHHHHBB", raw[:18])
    bpp = bits // 8
    body = raw[18 + idl:]
    if typ in (2, 3):
        px = np.frombuffer(body[:w * h * bpp], dtype=np.uint8)
    else:
        out = bytearray()
        i = 0
        while len(out) < w * h * bpp:
            hd = body[i]
            i += 1
            n = (hd & 127) + 1
            if hd < 128:
                out += body[i:i + n * bpp]
                i += n * bpp
            else:
                out += body[i:i + bpp] * n
                i += bpp
        px = np.frombuffer(bytes(out), dtype=np.uint8)
    return px.reshape(h, w, bpp)  # rows in file order (descriptor 0x00: row 0 = bottom = framebuffer y 0)


@pytest.fixture(scope="module")
def assets(tmp_path_factory, built):
    d = str(tmp_path_factory.mktemp("assets"))
    sc = scenes.orbit_scene(W, H, room_quads=((24, 12), (24, 6), (12, 12)), head_res=(20, 14), eye_res=(10, 8),
                            tex_size=64)
    names = {"room": "sponza", "head": "head", "eyes": "eyes"}
    for it in sc.items:
        stem = os.path.join(d, names[it.mesh.name])
        write_obj(stem + ".obj", it.mesh)
        for key, sfx in (("diffuse", "_diffuse"), ("normal", "_nm"), ("specular", "_spec")):
            if key in it.textures:
                write_tga(stem + sfx + ".tga", it.textures[key])
    return d, sc


def run_example(exe, assets_dir, outdir, *extra):
    os.makedirs(outdir, exist_ok=True)
    cmd = [exe, os.path.join(assets_dir, "head.obj"), os.path.join(assets_dir, "eyes.obj"),
           os.path.join(assets_dir, "sponza.obj"), str(W), str(H), outdir] + list(extra)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    out = {n: read_tga(os.path.join(outdir, n + ".tga")) for n in ("phong", "zbuffer", "ao", "final")}
    out["z"] = np.fromfile(os.path.join(outdir, "zbuffer.bin"), dtype=np.float64).reshape(H, W)
    out["stderr"] = res.stderr
    out["stdout"] = res.stdout
    out["dir"] = outdir
    return out


def need(path):
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference)" % os.path.relpath(path, ROOT))


def test_oracle_build_matches_python_driven_reference(assets, tmp_path, ref_api):
    need(EXAMPLE_REF)
    d, sc = assets
    got = run_example(EXAMPLE_REF, d, str(tmp_path / "ref"))
    # same scene through the Python frame driver + the reference's rasterize (arrays, no OBJ/TGA files)
    with trb.Renderer(ref_api) as r:
        up = scenes.UploadedScene(r, sc)
        view = ref_api.lookat([-3.4019, 2.2001, 1.8026], [1.3555, 1.5116, -0.9686], [0, 1, 0])
        up.render(view[None], ref_api.perspective(70.0, W / H, 0.05, 500.0))
        z = r.read_depth()
        ao = r.ssao()
        zi = r.depth_image()
    assert np.array_equal(got["z"].view(np.uint64), z.view(np.uint64))     # OBJ loader + draw flow + z restore
    assert np.array_equal(got["ao"][:, :, 0], ao) and np.array_equal(got["zbuffer"][:, :, 0], zi)
    assert "triangles=" in got["stderr"]


# the camera looks straight at the head and the reference's frustum (planes of the transposed matrix) still drops it
CULL_CAMERA = ["--eye", "-1.869", "0.582", "-0.816", "--target", "0.412", "2.723", "-0.129"]


def test_oracle_build_culls_like_the_python_driver(assets, tmp_path, ref_api):
    """a camera for which main()'s frustum test (the reference's own Frustum in the oracle build) drops the head and
    the eyes: the Python frame driver (trb_cull_batch) must skip the same models"""
    need(EXAMPLE_REF)
    d, sc = assets
    got = run_example(EXAMPLE_REF, d, str(tmp_path / "ref"), *CULL_CAMERA)
    assert "frustum sponza 1 head 0" in got["stdout"] and "models rendered 1 culled 1" in got["stdout"]
    with trb.Renderer(ref_api) as r:
        up = scenes.UploadedScene(r, sc)
        view = ref_api.lookat([-1.869, 0.582, -0.816], [0.412, 2.723, -0.129], [0, 1, 0])
        up.render(view[None], ref_api.perspective(70.0, W / H, 0.05, 500.0))
        assert up.culled == 2
        z = r.read_depth()
    assert np.array_equal(got["z"].view(np.uint64), z.view(np.uint64))


def test_tga_writer_is_byte_identical_to_the_reference(assets, tmp_path, built):
    need(TGA_TOOL_REF)
    d, _ = assets
    rng = np.random.default_rng(1)
    img = np.zeros((40, 300, 3), np.uint8)
    img[:, :100] = rng.integers(0, 255, (40, 100, 3))            # raw packets
    img[:, 100:260] = rng.integers(0, 255, (40, 1, 3))            # runs longer than 128
    img[::3, 260:] = 7                                            # short runs / alternations
    grey = rng.integers(0, 3, (33, 77, 1)).astype(np.uint8) * 100
    for name, arr in (("rgb", img), ("grey", grey), ("diffuse", None)):
        src = os.path.join(d, "head_diffuse.tga") if arr is None else str(tmp_path / (name + ".tga"))
        if arr is not None:
            write_tga(src, arr)
        for mode in ([], ["raw"]):
            a, b = str(tmp_path / "ours.tga"), str(tmp_path / "ref.tga")
            assert subprocess.run([TGA_TOOL, src, a] + mode).returncode == 0
            assert subprocess.run([TGA_TOOL_REF, src, b] + mode).returncode == 0
            assert open(a, "rb").read() == open(b, "rb").read(), (name, mode)
            # and our reader understands what was written (RLE round trip)
            c = str(tmp_path / "again.tga")
            assert subprocess.run([TGA_TOOL, a, c, "raw"]).returncode == 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [[], ["--immediate"], CULL_CAMERA])
def test_device_build_matches_oracle_build(assets, tmp_path, mode):
    """the same example source: device build (libtrb.so) vs oracle build (the reference's our_gl.cpp); the third
    variant uses a camera whose frustum test drops the head and the eyes (main.cpp:680, 706)"""
    need(EXAMPLE_REF)
    d, _ = assets
    cam = mode if mode == CULL_CAMERA else []
    want = run_example(EXAMPLE_REF, d, str(tmp_path / "ref"), *cam)
    got = run_example(EXAMPLE, d, str(tmp_path / "dev"), *mode)
    assert [l for l in got["stdout"].splitlines() if l.startswith(("frustum", "models"))] == \
           [l for l in want["stdout"].splitlines() if l.startswith(("frustum", "models"))]
    assert np.array_equal(got["z"].view(np.uint64), want["z"].view(np.uint64))
    assert np.array_equal(got["ao"], want["ao"]) and np.array_equal(got["zbuffer"], want["zbuffer"])
    for k in ("phong", "final"):
        diff = np.abs(got[k].astype(int) - want[k].astype(int)).max(axis=-1)
        assert (diff <= 1).mean() >= 0.999, k
    # ao.tga / zbuffer.tga files written by our TGA writer are byte-identical to the reference's
    for n in ("ao", "zbuffer"):
        assert open(os.path.join(got["dir"], n + ".tga"), "rb").read() == open(os.path.join(want["dir"], n + ".tga"), "rb").read()
    # gl_write_tga_file: the same four files packetised on the device (trb_encode_tga) are byte-identical
    # to what the host-side TGAImage::write_tga_file wrote from the read-back pixels
    for n in ("phong", "zbuffer", "ao", "final"):
        a = open(os.path.join(got["dir"], n + "_dev.tga"), "rb").read()
        b = open(os.path.join(got["dir"], n + ".tga"), "rb").read()
        assert a == b, n


@pytest.mark.gpu
def test_unknown_shader_is_an_error_not_a_fallback(tmp_path, built):
    src = tmp_path / "unknown.cpp"
    src.write_text('''#include <our_gl.h>
#include <iostream>
struct Mine : IShader { std::pair<bool, TGAColor> fragment(const vec3) const override { return {false, TGAColor()}; } };
int main() { TGAImage fb(8, 8, TGAImage::RGB); init_zbuffer(8, 8); Mine s; vec4 clip[3];
  try { rasterize(clip, s, fb); } catch (const std::exception& e) { std::cout << "threw: " << e.what() << std::endl; return 0; }
  return 1; }
''')
    host = os.path.join(ROOT, "tinyrenderder_b200", "host")
    exe = str(tmp_path / "unknown")
    subprocess.check_call(["g++", "-std=c++17", "-I", host, str(src), os.path.join(host, "our_gl.cpp"),
                           os.path.join(host, "model.cpp"), os.path.join(host, "tgaimage.cpp"), "-L",
                           os.path.join(ROOT, "tinyrenderder_b200"), "-ltrb",
                           "-Wl,-rpath," + os.path.join(ROOT, "tinyrenderder_b200"), "-o", exe])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0 and "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_views_api_equals_frame_by_frame(assets, tmp_path, built):
    """gl_begin_views / gl_draw_model_views (several cameras in one launch set, the C++ face of config 3)
    give the frames the classic one-frame-at-a-time calls give, and gl_write_tga_files writes the same
    files as TGAImage::write_tga_file on the read-back pixels"""
    d, _ = assets
    src = tmp_path / "views.cpp"
    src.write_text(r'''#include <our_gl.h>
#include <model.h>
#include <model_manager.h>
#include <shaders.h>
#include <fstream>
#include <iostream>
static void dump(const std::string& name, const TGAImage& fb) {
    std::ofstream c(name + ".bgr", std::ios::binary);
    c.write((const char*)const_cast<TGAImage&>(fb).buffer(), (std::streamsize)fb.width() * fb.height() * 3);
    std::ofstream z(name + ".z", std::ios::binary);
    z.write((const char*)zbuffer.data(), (std::streamsize)zbuffer.size() * sizeof(double));
}
int main(int argc, char** argv) {
    const std::string dir = argv[1], out = argv[2];
    const int W = 320, H = 200;
    auto head = ModelManager::getInstance().loadModel(dir + "/head.obj");
    auto eyes = ModelManager::getInstance().loadModel(dir + "/eyes.obj");
    if (!head || !eyes) return 2;
    const vec3 key{1.0, 1.2, 1.0}, fill{-1.0, 0.3, 0.5}, rim{0.0, 0.8, -1.0}, center{0.0, 0.0, 0.0}, up{0.0, 1.0, 0.0};
    const vec3 cams[3] = {vec3{1.0, 1.0, 3.0}, vec3{-2.0, 0.5, 2.5}, vec3{0.3, 2.2, 2.0}};
    init_perspective(60.0, (double)W / H, 0.1, 100.0);
    init_viewport(0, 0, W, H);
    std::vector<mat<4, 4>> views;
    for (int k = 0; k < 3; ++k) {                       // classic: one frame at a time
        lookat(cams[k], center, up);
        views.push_back(ModelView);
        TGAImage fb(W, H, TGAImage::RGB);
        init_zbuffer(W, H);
        PhongShader sh(head.get());
        sh.initLightDirections(key, fill, rim);
        sh.normal_map_strength = 1.0;
        gl_draw_model(*head, sh, fb);
        EyeShader eye(eyes.get());
        eye.initLightDirections(key, rim);
        gl_draw_model(*eyes, eye, fb);
        gl_flush(fb);
        dump(out + "/single" + std::to_string(k), fb);
        fb.write_tga_file(out + "/single" + std::to_string(k) + ".tga");
    }
    gl_begin_views(views, W, H);                        // the same three frames in one launch set
    gl_draw_model_views(*head, 1, mat<4, 4>::identity(), key, fill, rim, 1.0);
    gl_draw_model_views(*eyes, 2, mat<4, 4>::identity(), key, fill, rim, 1.0);
    gl_write_tga_files(0, {out + "/batch0.tga", out + "/batch1.tga", out + "/batch2.tga"});
    for (int k = 0; k < 3; ++k) {
        TGAImage fb;
        gl_read_view(k, fb);
        dump(out + "/batch" + std::to_string(k), fb);
    }
    std::cout << "ok" << std::endl;
    return 0;
}
''')
    host = os.path.join(ROOT, "tinyrenderder_b200", "host")
    exe = str(tmp_path / "views")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", host, str(src)] +
                          [os.path.join(host, f) for f in ("our_gl.cpp", "model.cpp", "model_manager.cpp", "tgaimage.cpp")] +
                          ["-L", os.path.join(ROOT, "tinyrenderder_b200"), "-ltrb",
                           "-Wl,-rpath," + os.path.join(ROOT, "tinyrenderder_b200"), "-o", exe])
    out = tmp_path / "views_out"
    out.mkdir()
    res = subprocess.run([exe, d, str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    for k in range(3):
        for ext in ("bgr", "z", "tga"):
            a = open(out / ("single%d.%s" % (k, ext)), "rb").read()
            b = open(out / ("batch%d.%s" % (k, ext)), "rb").read()
            assert a == b, (k, ext)


@pytest.mark.gpu
def test_views_recording_replays_for_other_cameras(assets, tmp_path, built):
    """gl_record_views_begin / gl_record_views_end / gl_replay_views: a batch (head, z snapshot, eyes, restore) captured as
    one CUDA graph per context and re-run for other cameras gives the frames the plain calls give - on one context and on
    two (cameras split in blocks)"""
    d, _ = assets
    src = tmp_path / "record.cpp"
    src.write_text(r'''#include <our_gl.h>
#include <model.h>
#include <model_manager.h>
#include <shaders.h>
#include <cstdlib>
#include <fstream>
#include <iostream>
static void dump(const std::string& name, int nviews) {
    for (int k = 0; k < nviews; ++k) {
        TGAImage fb;
        gl_read_view(k, fb);
        std::ofstream c(name + std::to_string(k) + ".bgr", std::ios::binary);
        c.write((const char*)fb.buffer(), (std::streamsize)fb.width() * fb.height() * 3);
        std::ofstream z(name + std::to_string(k) + ".z", std::ios::binary);
        z.write((const char*)zbuffer.data(), (std::streamsize)zbuffer.size() * sizeof(double));
    }
}
int main(int argc, char** argv) {
    const std::string dir = argv[1], out = argv[2];
    const int ndev = atoi(argv[3]);
    if (ndev > 1) gl_set_devices(std::vector<int>((size_t)ndev, 0));
    const int W = 320, H = 200;
    auto head = ModelManager::getInstance().loadModel(dir + "/head.obj");
    auto eyes = ModelManager::getInstance().loadModel(dir + "/eyes.obj");
    if (!head || !eyes) return 2;
    const vec3 key{1.0, 1.2, 1.0}, fill{-1.0, 0.3, 0.5}, rim{0.0, 0.8, -1.0}, center{0.0, 0.0, 0.0}, up{0.0, 1.0, 0.0};
    const vec3 cams[2][3] = {{vec3{1.0, 1.0, 3.0}, vec3{-2.0, 0.5, 2.5}, vec3{0.3, 2.2, 2.0}},
                             {vec3{2.0, 0.2, 2.0}, vec3{-1.0, 1.5, 2.8}, vec3{0.1, -1.0, 3.0}}};
    init_perspective(60.0, (double)W / H, 0.1, 100.0);
    init_viewport(0, 0, W, H);
    std::vector<mat<4, 4>> views[2];
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < 3; ++k) { lookat(cams[s][k], center, up); views[s].push_back(ModelView); }
    auto batch = [&](const std::vector<mat<4, 4>>& v) {
        gl_begin_views(v, W, H);
        gl_draw_model_views(*head, 1, mat<4, 4>::identity(), key, fill, rim, 1.0);
        gl_zbuffer_snapshot();
        gl_draw_model_views(*eyes, 2, mat<4, 4>::identity(), key, fill, rim, 1.0);
        TGAImage none;
        gl_zbuffer_restore(none);
    };
    batch(views[1]); dump(out + "/plainB", 3);
    batch(views[0]); dump(out + "/plainA", 3);          // also the warm-up of the recording
    gl_record_views_begin();
    batch(views[0]);
    const int rec = gl_record_views_end();
    dump(out + "/recA", 3);
    gl_replay_views(rec, views[1]); dump(out + "/repB", 3);
    gl_replay_views(rec, views[0]); dump(out + "/repA", 3);
    std::cout << "ok" << std::endl;
    return 0;
}
''')
    host = os.path.join(ROOT, "tinyrenderder_b200", "host")
    exe = str(tmp_path / "record")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", host, str(src)] +
                          [os.path.join(host, f) for f in ("our_gl.cpp", "model.cpp", "model_manager.cpp", "tgaimage.cpp")] +
                          ["-L", os.path.join(ROOT, "tinyrenderder_b200"), "-ltrb",
                           "-Wl,-rpath," + os.path.join(ROOT, "tinyrenderder_b200"), "-o", exe])
    for ndev in (1, 2):
        out = tmp_path / ("record_out%d" % ndev)
        out.mkdir()
        res = subprocess.run([exe, d, str(out), str(ndev)], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        for a, b in (("plainA", "recA"), ("plainB", "repB"), ("plainA", "repA")):
            for k in range(3):
                for ext in ("bgr", "z"):
                    x = open(out / ("%s%d.%s" % (a, k, ext)), "rb").read()
                    y = open(out / ("%s%d.%s" % (b, k, ext)), "rb").read()
                    assert x == y and len(x) > 0, (ndev, a, b, k, ext)


@pytest.mark.gpu
def test_two_contexts_from_cpp_equal_one(assets, tmp_path, built):
    """gl_set_devices: the C++ host layer driving two contexts (both on GPU 0 here).  A batch of cameras is split over
    the contexts in blocks (config 3, nothing exchanged); one picture is split by triangle ranges and put together by
    the composite group of the C ABI (config 4: trb_comm_init + trb_composite_group).  Both must equal what a single
    context renders - z-buffer bits, colours, TGA files."""
    d, _ = assets
    src = tmp_path / "multi.cpp"
    src.write_text(r'''#include <our_gl.h>
#include <model.h>
#include <model_manager.h>
#include <shaders.h>
#include <cstdlib>
#include <fstream>
#include <iostream>
static void dump(const std::string& name, const TGAImage& fb) {
    std::ofstream c(name + ".bgr", std::ios::binary);
    c.write((const char*)const_cast<TGAImage&>(fb).buffer(), (std::streamsize)fb.width() * fb.height() * 3);
    std::ofstream z(name + ".z", std::ios::binary);
    z.write((const char*)zbuffer.data(), (std::streamsize)zbuffer.size() * sizeof(double));
}
int main(int argc, char** argv) {
    const std::string dir = argv[1], out = argv[2];
    const int ndev = atoi(argv[3]);
    if (ndev > 1) gl_set_devices(std::vector<int>(ndev, 0));
    if (gl_device_count() != ndev) return 3;
    const int W = 320, H = 200;
    auto head = ModelManager::getInstance().loadModel(dir + "/head.obj");
    auto eyes = ModelManager::getInstance().loadModel(dir + "/eyes.obj");
    if (!head || !eyes) return 2;
    const vec3 key{1.0, 1.2, 1.0}, fill{-1.0, 0.3, 0.5}, rim{0.0, 0.8, -1.0}, center{0.0, 0.0, 0.0}, up{0.0, 1.0, 0.0};
    const vec3 cams[5] = {vec3{1.0, 1.0, 3.0}, vec3{-2.0, 0.5, 2.5}, vec3{0.3, 2.2, 2.0}, vec3{2.5, -0.4, 1.0}, vec3{0.0, 0.2, -3.0}};
    init_perspective(60.0, (double)W / H, 0.1, 100.0);
    init_viewport(0, 0, W, H);
    std::vector<mat<4, 4>> views;
    for (int k = 0; k < 5; ++k) { lookat(cams[k], center, up); views.push_back(ModelView); }
    // config 3: five cameras, blocks of 3 + 2 when there are two contexts
    gl_begin_views(views, W, H);
    gl_draw_model_views(*head, 1, mat<4, 4>::identity(), key, fill, rim, 1.0);
    gl_zbuffer_snapshot();
    gl_draw_model_views(*eyes, 2, mat<4, 4>::identity(), key, fill, rim, 1.0);
    TGAImage unused;
    gl_zbuffer_restore(unused);
    std::vector<std::string> names;
    for (int k = 0; k < 5; ++k) names.push_back(out + "/view" + std::to_string(k) + ".tga");
    if (!gl_write_tga_files(0, names)) return 4;
    for (int k = 0; k < 5; ++k) { TGAImage fb; gl_read_view(k, fb); dump(out + "/view" + std::to_string(k), fb); }
    // config 4: one picture, two models, triangle ranges per context + sort-last composite
    for (int k = 0; k < 2; ++k) {                 // twice: the composite group is reused, frames follow each other without a host barrier
        TGAImage fb(W, H, TGAImage::RGB);
        init_zbuffer(W, H);
        lookat(cams[k], center, up);
        PhongShader sh(head.get());
        sh.initLightDirections(key, fill, rim);
        sh.normal_map_strength = 1.0;
        gl_draw_model(*head, sh, fb);
        EyeShader eye(eyes.get());
        eye.initLightDirections(key, rim);
        gl_draw_model(*eyes, eye, fb);
        gl_flush(fb);
        dump(out + "/picture" + std::to_string(k), fb);
    }
    std::cout << "ok" << std::endl;
    return 0;
}
''')
    host = os.path.join(ROOT, "tinyrenderder_b200", "host")
    exe = str(tmp_path / "multi")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", host, str(src)] +
                          [os.path.join(host, f) for f in ("our_gl.cpp", "model.cpp", "model_manager.cpp", "tgaimage.cpp")] +
                          ["-L", os.path.join(ROOT, "tinyrenderder_b200"), "-ltrb",
                           "-Wl,-rpath," + os.path.join(ROOT, "tinyrenderder_b200"), "-o", exe])
    outs = {}
    for ndev in (1, 2, 3):
        out = tmp_path / ("multi_out%d" % ndev)
        out.mkdir()
        res = subprocess.run([exe, d, str(out), str(ndev)], capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
        outs[ndev] = out
    files = sorted(f for f in os.listdir(outs[1]))
    assert len(files) == 5 * 3 + 2 * 2
    for ndev in (2, 3):
        for f in files:
            a, b = open(outs[1] / f, "rb").read(), open(outs[ndev] / f, "rb").read()
            assert a == b, (ndev, f)


MODEL_TOOL = os.path.join(ROOT, "tinyrenderder_b200", "host", "bin", "model_tool")


def _load_with_tool(path):
    res = subprocess.run([MODEL_TOOL, path], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    out = {"sub": [], "mat": [], "v": []}
    for line in res.stdout.splitlines():
        f = line.split()
        if f[0] == "vertices":
            out["counts"] = (int(f[1]), int(f[3]), int(f[5]), int(f[7]))
        elif f[0] == "submesh":
            out["sub"].append({"name": f[3], "start": int(f[5]), "count": int(f[7]), "material": int(f[9]), "vertexStart": int(f[11]),
                               "normals": int(f[13]), "uvs": int(f[15])})
        elif f[0] == "material":
            out["mat"].append(tuple(int(f[k]) for k in (3, 5, 7, 9)))
        elif f[0] == "indices":
            out["idx"] = [int(x) for x in f[1:]]
        elif f[0] == "v":
            out["v"].append([float(f[k]) for k in (1, 2, 3, 5, 6, 7, 9, 10)])
    return out


def test_loader_flattens_submeshes_and_reads_materials(tmp_path, built):
    """model.cpp:143-205 / 207-267 / 269-312 without Assimp: sub-meshes per usemtl / o / g with their own vertex numbering
    flattened with vertexStart, materials in newmtl order with textures from the .mtl or the <stem>_*.tga fallback, normals
    regenerated for ALL vertices as soon as one is missing"""
    d = tmp_path
    rng = np.random.default_rng(0)
    for name, size in (("skin.tga", 8), ("cloth_nm.tga", 4), ("thing_spec.tga", 2), ("thing_diffuse.tga", 16)):
        write_tga(str(d / name), rng.integers(0, 255, (size, size, 3)).astype(np.uint8))
    (d / "thing.mtl").write_text("# two materials\nnewmtl skin\nKd 1 1 1\nmap_Kd skin.tga\n\nnewmtl cloth\nmap_Bump -bm 1.0 cloth_nm.tga\n"
                                 "map_Kd missing_file.tga\n")
    (d / "thing.obj").write_text("""mtllib thing.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
v 1 0 1
vt 0 0
vt 1 0
vt 1 1
vn 0 0 1
o body
usemtl skin
f 1/1/1 2/2/1 3/3/1 4/1/1
usemtl cloth
f 1/1/1 2/2/1 6/3/1
g tail
f 5 6 3
usemtl skin
f 5/1 1/2 4/3
""")
    m = _load_with_tool(str(d / "thing.obj"))
    # four sub-meshes: (body, skin) quad = 2 triangles; (body, cloth) 1; (tail, cloth) 1 without vt/vn; (tail, skin) 1 without vn
    assert [(s["name"], s["material"], s["count"]) for s in m["sub"]] == [("body", 0, 6), ("body", 1, 3), ("tail", 1, 3), ("tail", 0, 3)]
    assert [s["start"] for s in m["sub"]] == [0, 6, 9, 12]
    assert [s["vertexStart"] for s in m["sub"]] == [0, 4, 7, 10]         # vertices are NOT shared across sub-meshes
    assert m["counts"] == (13, 15, 4, 2)
    assert m["idx"] == [0, 1, 2, 0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]   # fan triangulation + vertexStart
    assert [(s["normals"], s["uvs"]) for s in m["sub"]] == [(1, 1), (1, 1), (0, 0), (0, 1)]
    # material 0: diffuse from the .mtl, specular from the fallback name; material 1: normal map from the .mtl (options
    # skipped), its diffuse names a file that does not exist -> fallback thing_diffuse.tga
    assert m["mat"] == [(8, 0, 2, 0), (16, 4, 2, 0)]
    # uv.y is flipped (aiProcess_FlipUVs); one vertex lacked a normal, so ALL were regenerated from the faces, unit length
    assert m["v"][1][6:] == [1.0, 1.0] and m["v"][2][6:] == [1.0, 0.0]
    nrm = np.array([v[3:6] for v in m["v"]])
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-6)
    assert np.allclose(nrm[:4], [[0, 0, 1]] * 4)                            # the quad keeps +z; computed, not copied


def test_grouped_obj_renders_like_the_flat_one(assets, tmp_path, ref_api):
    """the same head, written once as a single mesh and once split into groups / materials in the middle of its face list:
    the flattened arrays differ (vertices are duplicated at the seams) but the reference's rasterize() draws the same picture"""
    need(EXAMPLE_REF)
    d, sc = assets
    import shutil
    g = tmp_path / "grouped"
    shutil.copytree(d, g)
    head = open(os.path.join(d, "head.obj")).read().splitlines()
    faces = [i for i, l in enumerate(head) if l.startswith("f ")]
    cut1, cut2 = faces[len(faces) // 3], faces[2 * len(faces) // 3]
    head.insert(cut2, "usemtl b\ng lower")
    head.insert(cut1, "usemtl a")
    head.insert(faces[0], "mtllib head.mtl\no head\nusemtl b")
    (g / "head.obj").write_text("\n".join(head) + "\n")
    (g / "head.mtl").write_text("newmtl b\nmap_Kd head_diffuse.tga\nnewmtl a\n")
    flat = run_example(EXAMPLE_REF, d, str(tmp_path / "flat"))
    grouped = run_example(EXAMPLE_REF, str(g), str(tmp_path / "grp"))
    assert np.array_equal(flat["z"].view(np.uint64), grouped["z"].view(np.uint64))
    assert np.array_equal(flat["phong"], grouped["phong"])
    m = _load_with_tool(str(g / "head.obj"))
    assert m["counts"][2] == 3 and m["counts"][3] == 2

"""TGA RLE encoding: TGAImage::write_tga_file / unload_rle_data
(tgaimage.cpp:160-242).  Three implementations must agree byte for byte on the same images:
the reference's own writer (oracle/_ref, compiled from /root/reference/tgaimage.cpp), the
self-contained restatement (oracle/libtrb_port.so) and the device encoder (tga_rle.cuh, -m gpu)."""
import numpy as np
import pytest

import tinyrenderder_b200 as trb
from tinyrenderder_b200 import capi, scenes


def _seq_image(values):
    """1-row image whose pixel i is the colour number values[i] (distinct numbers = distinct colours)"""
    v = np.asarray(values, dtype=np.int64)
    img = np.empty((1, v.size, 3), dtype=np.uint8)
    img[0, :, 0] = v & 255
    img[0, :, 1] = (v >> 8) & 255
    img[0, :, 2] = (v * 7 + 3) & 255
    return img


def _runs(lengths):
    """pixel sequence made of runs of the given lengths, neighbouring runs always differ"""
    out = []
    for k, n in enumerate(lengths):
        out += [k] * n
    return _seq_image(out)


def rle_images():
    rng = np.random.default_rng(42)
    imgs = {}
    imgs["one_pixel"] = _runs([1])
    imgs["two_equal"] = _runs([2])
    imgs["two_different"] = _runs([1, 1])
    for n in (127, 128, 129, 130, 255, 256, 257, 258, 384, 385, 1000):
        imgs["single_run_%d" % n] = _runs([n])
        imgs["all_different_%d" % n] = _runs([1] * n)
    # a long run reached inside an open raw packet (first pixel swallowed) at every interesting phase
    for lead in (0, 1, 2, 126, 127, 128, 129, 255, 256):
        for run in (2, 3, 4, 128, 129, 130, 131, 257):
            imgs["lead%d_run%d" % (lead, run)] = _runs([1] * lead + [run] + [1] * 5 + [3] + [1] * 130 + [2])
    imgs["pairs"] = _runs([2] * 300)
    imgs["pair_single"] = _runs([2, 1] * 300)
    imgs["single_pair"] = _runs([1, 2] * 300)
    imgs["triples_singles"] = _runs([3, 1, 1] * 200)
    imgs["ends_with_lone_pixel_after_128"] = _runs([129])
    imgs["ends_in_raw_128"] = _runs([1] * 128)
    for p in (0.1, 0.5, 0.9, 0.99):
        n = 40000
        change = rng.random(n) > p
        imgs["random_p%.2f" % p] = _seq_image(np.cumsum(change))
    lens = rng.integers(1, 400, 300)
    imgs["random_runs"] = _runs(list(lens))
    # 2-D: runs continue across scanlines (the encoder never restarts at a row)
    img = np.zeros((37, 53, 3), dtype=np.uint8)
    img[5:20, 10:40] = (10, 200, 30)
    img[18:30, 0:53] = rng.integers(0, 255, (12, 53, 3), dtype=np.uint8)
    imgs["picture_2d"] = img
    imgs["noise_2d"] = rng.integers(0, 4, (64, 64, 3), dtype=np.uint8) * 80
    return imgs


def decode_tga(data):
    """independent decoder (TGA spec): returns (h, w, bpp) array in file order"""
    b = np.frombuffer(data, dtype=np.uint8)
    assert b[0] == 0 and b[1] == 0 and b[17] == 0
    w, h, bpp = int(b[12]) | int(b[13]) << 8, int(b[14]) | int(b[15]) << 8, int(b[16]) // 8
    assert b[2] == (11 if bpp == 1 else 10)
    out = np.empty((w * h, bpp), dtype=np.uint8)
    px, i = 0, 18
    while px < w * h:
        head = int(b[i])
        i += 1
        n = (head & 127) + 1
        if head & 128:
            out[px:px + n] = b[i:i + bpp]
            i += bpp
        else:
            out[px:px + n] = b[i:i + n * bpp].reshape(n, bpp)
            i += n * bpp
        px += n
    assert px == w * h and i == len(b)
    return out.reshape(h, w, bpp)


def encode_with(api, images):
    """one batch frame per image shape: every image becomes a view's framebuffer, then encode"""
    out = {}
    by_shape = {}
    for name, img in images.items():
        by_shape.setdefault(img.shape, []).append(name)
    with trb.Renderer(api) as r:
        for (h, w, _), names in by_shape.items():
            r.begin_frame(w, h, nviews=len(names))
            for v, name in enumerate(names):
                r.write_color(images[name], view=v)
            files = r.encode_tga(capi.IMAGE_COLOR)
            for name, f in zip(names, files):
                out[name] = f
    return out


IMAGES = rle_images()


def test_port_encoder_round_trips(port_api):
    files = encode_with(port_api, IMAGES)
    for name, img in IMAGES.items():
        assert np.array_equal(decode_tga(files[name]), img), name


def test_port_encoder_equals_the_reference_writer(port_api, ref_api):
    """pins the restated packetiser to the reference's own TGAImage::write_tga_file"""
    a, b = encode_with(port_api, IMAGES), encode_with(ref_api, IMAGES)
    for name in IMAGES:
        assert a[name] == b[name], name


def test_known_packets(port_api):
    """hand-checked against tgaimage.cpp:193-242"""
    f = encode_with(port_api, {"x": _runs([1, 1, 3, 1])})["x"][18:]
    # raw packet swallows the first pixel of the run of three: raw(3) = A B C, run(2) = C, raw(1) = D
    c = [bytes(_seq_image([k])[0, 0]) for k in range(4)]
    assert f == bytes([2]) + c[0] + c[1] + c[2] + bytes([129]) + c[2] + bytes([0]) + c[3]


@pytest.mark.gpu
def test_device_encoder_equals_the_oracle(cuda_api, port_api):
    a, b = encode_with(cuda_api, IMAGES), encode_with(port_api, IMAGES)
    for name in IMAGES:
        assert a[name] == b[name], "%s: %d vs %d bytes" % (name, len(a[name]), len(b[name]))


@pytest.mark.gpu
@pytest.mark.parametrize("which", [capi.IMAGE_COLOR, capi.IMAGE_DEPTH, capi.IMAGE_SSAO, capi.IMAGE_FINAL])
def test_device_encodes_rendered_frames_like_the_oracle(cuda_api, port_api, which):
    """the four files of main.cpp (framebuffer / zbuffer / ssao / final .tga) of a small orbit batch"""
    sc = scenes.orbit_scene(320, 180, room_quads=((16, 8), (16, 4), (8, 8)), tex_size=64)
    files = []
    for api in (cuda_api, port_api):
        views = scenes.orbit_views(api, [3, 400, 900])
        pr = api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
        with trb.Renderer(api) as r:
            up = scenes.UploadedScene(r, sc)
            up.render(views, pr)
            if which in (capi.IMAGE_COLOR, capi.IMAGE_FINAL):
                # colours may differ by one code between device and oracle (fp32 lighting): encode the same pixels
                if api is cuda_api:
                    colors = [r.read_color(v).copy() for v in range(3)]
                else:
                    for v in range(3):
                        r.write_color(colors[v], view=v)
            files.append(r.encode_tga(which))
    for v in range(3):
        assert files[0][v] == files[1][v], "view %d" % v
        assert len(files[0][v]) > 18


@pytest.mark.gpu
def test_device_encoder_full_hd_batch_round_trips(cuda_api):
    """BASELINE-size frames: 8 views of 1920x1080, decode(encode(x)) == x and sizes stay below raw"""
    rng = np.random.default_rng(7)
    w, h = 1920, 1080
    imgs = []
    for v in range(8):
        img = np.zeros((h, w, 3), dtype=np.uint8)
        img[100 + 20 * v:900, 200:1700] = rng.integers(0, 3, (800 - 20 * v, 1500, 3), dtype=np.uint8) * (60 + v)
        imgs.append(img)
    with trb.Renderer(cuda_api) as r:
        r.begin_frame(w, h, nviews=8)
        for v in range(8):
            r.write_color(imgs[v], view=v)
        files = r.encode_tga(capi.IMAGE_COLOR)
    for v in range(8):
        assert np.array_equal(decode_tga(files[v]), imgs[v])
        assert len(files[v]) < w * h * 3


def test_parallel_formulation_equals_the_sequential_packetiser(port_api):
    """the decomposition tga_rle.cuh runs on the device (segments, 2-bit maps, scans), modelled in numpy"""
    import rle_parallel_model as model
    want = encode_with(port_api, IMAGES)
    names = [n for n, img in IMAGES.items() if img.shape[0] == 1 and img.shape[1] <= 2000]
    for name in names:
        img = IMAGES[name]
        got = model.encode_views(img.reshape(-1, 3), img.shape[1], 1, 3)[0]
        assert got == want[name][18:], name
    # a batch: images of one shape back to back, packets must not leak across the image boundary
    same = [n for n in names if IMAGES[n].shape == (1, 1000, 3)]
    assert len(same) >= 3
    px = np.concatenate([IMAGES[n].reshape(-1, 3) for n in same])
    got = model.encode_views(px, 1000, len(same), 3)
    for n, g in zip(same, got):
        assert g == want[n][18:], n


@pytest.mark.gpu
def test_encode_tga_argument_errors(cuda_api):
    """a too small output buffer is an error (nothing is truncated silently); a NULL buffer is a size query"""
    import ctypes as C
    img = _runs([1, 3, 1, 130, 2])
    h, w = img.shape[:2]
    with trb.Renderer(cuda_api) as r:
        r.begin_frame(w, h)
        r.write_color(img)
        want = r.encode_tga(capi.IMAGE_COLOR)[0]
        sizes = (C.c_uint64 * 1)()
        table = (C.c_void_p * 1)(None)
        rc = cuda_api.fn["encode_tga"](r.h, 0, C.cast(table, C.c_void_p), 0, C.cast(sizes, C.c_void_p))
        assert rc == 0 and sizes[0] == len(want)
        small = np.empty(len(want) - 1, dtype=np.uint8)
        table = (C.c_void_p * 1)(small.ctypes.data)
        rc = cuda_api.fn["encode_tga"](r.h, 0, C.cast(table, C.c_void_p), small.size, C.cast(sizes, C.c_void_p))
        assert rc < 0
        rc = cuda_api.fn["encode_tga"](r.h, 7, C.cast(table, C.c_void_p), small.size, C.cast(sizes, C.c_void_p))
        assert rc < 0


@pytest.mark.gpu
def test_device_encoder_holds_its_packet_offset_invariants(built, port_api):
    """same images through the -DTRB_DEBUG_CHECKS build: every pixel asserts that it writes inside its segment"""
    import os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tinyrenderder_b200", "libtrb_checks.so")
    if not os.path.exists(p):
        pytest.skip("libtrb_checks.so not built")
    a, b = encode_with(trb.Api(p, "trb"), IMAGES), encode_with(port_api, IMAGES)
    for name in IMAGES:
        assert a[name] == b[name], name


@pytest.mark.gpu
def test_async_frame_writer_equals_the_blocking_one(cuda_api):
    """trb_encode_tga_async in a frame loop that never synchronises (two encodes in flight, packets travelling on the
    copy stream while the next batch renders): after readback_wait every file equals trb_encode_tga's"""
    import torch
    from tinyrenderder_b200 import scenes
    sc = scenes.orbit_scene(320, 180, room_quads=((16, 8), (16, 4), (8, 8)), tex_size=64)
    pr = cuda_api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
    steps = [[7 * k + 1, 7 * k + 300, 7 * k + 700] for k in range(5)]
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        want = []
        for ids in steps:
            up.render(scenes.orbit_views(cuda_api, ids), pr)
            want.append(r.encode_tga(capi.IMAGE_COLOR))
        cap = 320 * 180 * 3 + 320 * 180 // 2 + 64
        bufs = [[torch.empty(cap, dtype=torch.uint8).pin_memory().numpy() for _ in range(3)] for _ in steps]
        sizes = [np.zeros(3, dtype=np.uint64) for _ in steps]
        for k, ids in enumerate(steps):           # no synchronising call inside the loop
            up.render(scenes.orbit_views(cuda_api, ids), pr)
            r.encode_tga_async(bufs[k], sizes[k], capi.IMAGE_COLOR)
        r.readback_wait()
        for k in range(len(steps)):
            for v in range(3):
                assert bufs[k][v][:int(sizes[k][v])].tobytes() == want[k][v], (k, v)
        # the other images through the same path, and a wait with nothing in flight is harmless
        up.render(scenes.orbit_views(cuda_api, steps[0]), pr)
        for which in (capi.IMAGE_DEPTH, capi.IMAGE_SSAO, capi.IMAGE_FINAL):
            ref = r.encode_tga(which)
            r.encode_tga_async(bufs[0], sizes[0], which)
            r.readback_wait()
            for v in range(3):
                assert bufs[0][v][:int(sizes[0][v])].tobytes() == ref[v], (which, v)
        r.readback_wait()

"""GPU parity tests (run with -m gpu on a B200): every case goes through the C ABI of libtrb.so
and is compared with the CPU oracle on the same inputs - depth bit-exact, colour within 1 LSB on
>= 99.9 % of the pixels - and with the committed golden vectors made from the reference itself."""
import json
import os

import numpy as np
import pytest

import cases
import compare
import tinyrenderder_b200 as trb
from tinyrenderder_b200 import scenes

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    GOLDEN = json.load(f)


with open(os.path.join(HERE, "golden", "golden_fullsize.json")) as f:
    FULLSIZE = json.load(f)


def run_case(api, name):
    fn = cases.CASES.get(name) or cases.FULL_SIZE_CASES.get(name) or cases.DIGEST_ONLY_CASES[name]
    with trb.Renderer(api) as r:
        return fn(api, r)


def test_backend_is_cuda(cuda_api):
    assert cuda_api.backend_name() == "cuda-sm100a"
    with trb.Renderer(cuda_api) as r:
        assert r.launch_count() == 0
        cases.k1(cuda_api, r)
        assert r.launch_count() >= 6  # clear, vertex, setup, 3 scan, fill, raster, shade ...


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_matches_oracle(cuda_api, port_api, name):
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


@pytest.mark.parametrize("name", sorted(cases.FULL_SIZE_CASES))
def test_full_size_case_matches_oracle(cuda_api, port_api, name):
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


@pytest.mark.parametrize("name", sorted(set(cases.CASES) | set(cases.FULL_SIZE_CASES)))
def test_case_matches_golden_depth(cuda_api, name):
    """depth / ao / z-image digests made from the reference's own our_gl.cpp (tests/golden/make_golden.py)"""
    got = run_case(cuda_api, name)
    compare.assert_matches_golden(name, got, GOLDEN[name], skip=("bgr", "final"))


def test_c3_bench_step_matches_reference_digests(cuda_api):
    """BASELINE config 3 exactly as bench.py renders it: one 32-frame batch of the 1920x1080 orbit (and a second
    one across the k mod 1024 wrap); every frame's z-buffer must hash to what the reference's own rasterize()
    produced for that frame (tests/golden/golden_fullsize.json, all 1024 frames)"""
    want = FULLSIZE["c3_orbit"]["z_sha256"]
    sc = scenes.orbit_scene()
    assert FULLSIZE["c3_orbit"]["workload"] == "c3_orbit_%dx%d_%dtri" % (sc.width, sc.height, sc.ntris)
    pr = cuda_api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar)
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        for first in (0, 1008):
            frames = [(first + j) % 1024 for j in range(32)]
            up.render(scenes.orbit_views(cuda_api, frames), pr)
            for v, k in enumerate(frames):
                assert compare.sha(r.read_depth(v)) == want[k], "frame %d differs from the reference's z-buffer" % k


def test_c4_sphere_matches_reference_digests(cuda_api):
    """BASELINE config 4 at full size (20 971 520 triangles, 3840x2160): depth AND the flat-shaded colour"""
    got = run_case(cuda_api, "sphere_c4")
    g = FULLSIZE["c4_sphere"]
    assert int(np.isfinite(got["z"]).sum()) == g["pixels_shaded"]
    assert compare.sha(got["z"]) == g["z_sha256"] and compare.sha(got["bgr"]) == g["bgr_sha256"]


@pytest.mark.slow
def test_c5_soup_matches_reference_digests(cuda_api):
    """BASELINE config 5 at full size (100 000 000 sub-pixel triangles, 8192x8192): depth and colour digests of
    the reference's own output"""
    got = run_case(cuda_api, "soup_c5")
    g = FULLSIZE["c5_soup"]
    assert int(np.isfinite(got["z"]).sum()) == g["pixels_shaded"]
    assert compare.sha(got["z"]) == g["z_sha256"] and compare.sha(got["bgr"]) == g["bgr_sha256"]


def test_reference_library_agrees_when_present(cuda_api, ref_api):
    for name in ("k2", "rejects", "dense_tile", "head_small"):
        got = run_case(cuda_api, name)
        want = run_case(ref_api, name)
        compare.assert_outputs_match(name, got, want, skip=("stats_port",))


def test_survey_k7_counts_on_gpu(cuda_api):
    for name, px in (("k7a", 592743), ("k7b", 3571218)):
        got = run_case(cuda_api, name)
        assert int(np.isfinite(got["z"]).sum()) == px


def test_order_independence(cuda_api):
    """the (depth, id) resolve does not depend on how the hardware schedules CTAs: shuffling the
    submission order changes ids, so compare against the oracle for each order, and rendering the
    same order twice must be bit-identical"""
    clip, _ = scenes.triangle_soup(60000, 512, 512, 3.0, 9, False)
    with trb.Renderer(cuda_api) as r:
        r.begin_frame(512, 512)
        r.submit_clip_triangles(clip)
        a = (r.read_depth().copy(), r.read_color().copy())
        r.begin_frame(512, 512)
        r.submit_clip_triangles(clip)
        b = (r.read_depth(), r.read_color())
    assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64)) and np.array_equal(a[1], b[1])


def test_full_size_properties_without_oracle(cuda_api):
    """size-independent properties at a BASELINE-size frame (3840x2160, 1M triangles): splitting the
    draw in ranges, or drawing a prefix twice, leaves depth and colour unchanged (idempotence); the
    z-buffer equals the element-wise minimum of the two halves rendered separately (linearity of min)"""
    w, h = 3840, 2160
    _, pos = scenes.triangle_soup(1_000_000, w, h, 1.5, 12, True)
    n = pos.shape[0] // 3
    eye = np.eye(4)
    with trb.Renderer(cuda_api) as r:
        mesh = r.upload_mesh(pos)
        r.begin_frame(w, h)
        r.draw(mesh, eye, eye, ntris=n)
        z_all, c_all = r.read_depth().copy(), r.read_color().copy()
        s_all = r.stats()
        r.begin_frame(w, h)
        r.draw(mesh, eye, eye, first_tri=0, ntris=n // 2)
        z_a = r.read_depth().copy()
        r.begin_frame(w, h)
        r.draw(mesh, eye, eye, first_tri=n // 2, ntris=n - n // 2)
        z_b = r.read_depth().copy()
        r.begin_frame(w, h)
        r.draw(mesh, eye, eye, first_tri=0, ntris=n // 2)
        r.draw(mesh, eye, eye, first_tri=n // 2, ntris=n - n // 2)
        z_split, c_split = r.read_depth().copy(), r.read_color().copy()
        s_split = r.stats()
    assert np.array_equal(np.minimum(z_a, z_b).view(np.uint64), z_all.view(np.uint64))
    assert np.array_equal(z_split.view(np.uint64), z_all.view(np.uint64))
    assert np.array_equal(c_split, c_all)
    for k in ("triangles_submitted", "triangles_binned", "fragments_covered", "pixels_shaded", "tile_entries"):
        assert s_all[k] == s_split[k], k


def test_batch_equals_single_frames(cuda_api):
    """a batch of views rendered in one launch set equals the same views rendered one by one"""
    sc = scenes.orbit_scene(320, 180, room_quads=((16, 8), (16, 4), (8, 8)), tex_size=64)
    views = scenes.orbit_views(cuda_api, [3, 400, 900, 77])
    pr = cuda_api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        up.render(views, pr)
        batch = [(r.read_depth(v).copy(), r.read_color(v).copy()) for v in range(4)]
        for v in range(4):
            up.render(views[v:v + 1], pr)
            assert np.array_equal(r.read_depth(0).view(np.uint64), batch[v][0].view(np.uint64))
            assert np.array_equal(r.read_color(0), batch[v][1])


def test_pipelined_readback_equals_blocking_reads(cuda_api):
    """trb_readback_async (copy stream, staging area) returns what trb_read_color/depth return, also
    when the next frame is rendered while the copies are in flight"""
    import torch
    sc = scenes.orbit_scene(320, 180, room_quads=((16, 8), (16, 4), (8, 8)), tex_size=64)
    pr = cuda_api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        want, got = [], []
        for step in range(4):
            views = scenes.orbit_views(cuda_api, [10 * step + 1, 10 * step + 400, 10 * step + 800])
            up.render(views, pr)
            want.append([(r.read_depth(v).copy(), r.read_color(v).copy()) for v in range(3)])
        for step in range(4):
            views = scenes.orbit_views(cuda_api, [10 * step + 1, 10 * step + 400, 10 * step + 800])
            cs = [pin((180, 320, 3), torch.uint8) for _ in range(3)]
            ds = [pin((180, 320), torch.float64) for _ in range(3)]
            up.render(views, pr)
            r.readback_async(cs, ds)      # no wait: the next iteration renders while these copy
            got.append((cs, ds))
        r.readback_wait()
        for step in range(4):
            for v in range(3):
                assert np.array_equal(got[step][1][v].view(np.uint64), want[step][v][0].view(np.uint64))
                assert np.array_equal(got[step][0][v], want[step][v][1])


@pytest.mark.parametrize("warp_max", ["0", "8", "1000000000"])
@pytest.mark.parametrize("name", ["k2", "k7b_small", "signed_zero_ties", "duplicate_triangles", "big_triangles",
                                  "queue_overflow", "dense_tile", "soup_mesh_fp32", "orbit_small", "sub_range_draws"])
def test_both_raster_kernels_match_oracle(cuda_api, port_api, monkeypatch, name, warp_max):
    """TRB_WARP_MAX = 0 sends every bin to the CTA-per-tile kernel, a huge value sends every bin to the
    warp-per-tile kernel, 8 splits the tiles of one draw between the two: all must give the oracle's bits"""
    monkeypatch.setenv("TRB_WARP_MAX", warp_max)
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


@pytest.mark.parametrize("env", [{"TRB_WARP_MAX": "8", "TRB_SPLIT_S": "32"},
                                 {"TRB_WARP_MAX": "8", "TRB_SPLIT_S": "32", "TRB_SPLIT_CAP": "12"},
                                 {"TRB_WARP_MAX": "64", "TRB_SPLIT_S": "64", "TRB_MESH_ORDER_MIN_TRIS": "1"},
                                 {"TRB_WARP_MAX": "8", "TRB_SPLIT": "0"}])
@pytest.mark.parametrize("name", ["k2", "k7b_small", "signed_zero_ties", "duplicate_triangles", "big_triangles", "queue_overflow",
                                  "dense_tile", "soup_mesh_fp32", "orbit_small", "sub_range_draws", "indexed_duplicates",
                                  "snapshot_restore_twice", "k5_far_near"])
def test_split_bins_match_oracle(cuda_api, port_api, monkeypatch, name, env):
    """bins longer than TRB_WARP_MAX are cut into slices of TRB_SPLIT_S triangles, one warp each; every slice starts from
    an empty tile and the last one to finish folds the tile's slices into the frame.  Small values make almost every tile
    take that path (and, with a 12-entry item list, spill most of them over to the CTA-per-tile kernel): the oracle's
    bits and counters must come out, ties included."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


@pytest.mark.parametrize("env", [{}, {"TRB_LAZY_SNAPSHOT": "0"}, {"TRB_LAZY_SNAPSHOT_MAX_TRIS": "0"}, {"TRB_BIN_CAP": "16"}, {"TRB_COLLECT_BY_TILES": "0"},
                                 {"TRB_SYNC_DRAWS": "1"}, {"TRB_WARP_MAX": "0"}, {"TRB_WARP_MAX": "8", "TRB_SPLIT_S": "32"}])
@pytest.mark.parametrize("name", ["snapshot_restore_twice", "snapshot_signed_zero", "snapshot_small_triangles", "orbit_small",
                                  "orbit_culled"])
def test_depth_snapshot_variants_match_oracle(cuda_api, port_api, monkeypatch, name, env):
    """`zbuffer_before_eyes = zbuffer` / `zbuffer = zbuffer_before_eyes` (main.cpp:700, 730).  Default: the tile-granular
    snapshot - draws inside the window save the tiles their bins name (and give up the direct path for it);
    TRB_LAZY_SNAPSHOT_MAX_TRIS=0 makes every such draw save ALL tiles and keep the direct path (what a draw of more than
    2^20 triangles does), a 16-entry bin buffer makes it overflow into the unbinned kernels (again all tiles);
    TRB_LAZY_SNAPSHOT=0 is the whole-plane copy with the pointer-swap restore.  Flushes inside the window collect their
    pixel list from the saved tiles (TRB_COLLECT_BY_TILES=0: from the whole id plane).  Same bits, same counters, every time."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


LIT_CASES = ["head_small", "orbit_small", "shadow_small", "gouraud_small", "lit_clip_triangles"]


@pytest.mark.parametrize("name", LIT_CASES)
def test_exact_shading_mode_matches_oracle(cuda_api, port_api, monkeypatch, name):
    """TRB_SHADE_EXACT=1: all-fp64 lighting in the reference's operation order - depth bit-exact and
    the colour equal to the oracle's on (practically) every pixel"""
    monkeypatch.setenv("TRB_SHADE_EXACT", "1")
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)
    for key in want:
        if key.startswith("bgr"):
            same = float((got[key] == want[key]).all(axis=-1).mean())
            assert same >= 0.9999, "%s/%s: only %.6f of the pixels identical in exact mode" % (name, key, same)


@pytest.mark.parametrize("name", LIT_CASES)
def test_fp32_lighting_stays_within_one_code(cuda_api, port_api, monkeypatch, name):
    """default mode (fp32 lighting, fp64 coverage / depth / barycentrics / texel choice): no channel of
    any pixel is further than 1 LSB from the oracle - stricter than north_star's 99.9 % - because the
    lighting is continuous up to the final truncation"""
    monkeypatch.delenv("TRB_SHADE_EXACT", raising=False)
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)
    for key in want:
        if key.startswith("bgr"):
            d = np.abs(got[key].astype(np.int32) - want[key].astype(np.int32))
            frac = float((d.max(axis=-1) <= 1).mean())
            assert frac >= 0.9999, "%s/%s: %.6f within 1 LSB, max diff %d" % (name, key, frac, int(d.max()))


DRAW_PATH_CASES = ["k2", "k5_far_near", "k7b_small", "signed_zero_ties", "duplicate_triangles", "big_triangles",
                   "dense_tile", "soup_mesh_fp32", "head_small", "orbit_small", "sub_range_draws", "rejects",
                   "snapshot_restore_twice", "indexed_duplicates"]


@pytest.mark.parametrize("name", DRAW_PATH_CASES)
def test_bin_overflow_takes_the_unbinned_kernels(cuda_api, port_api, monkeypatch, name):
    """draws are enqueued without a host round trip, so the bin buffer is sized from an estimate;
    TRB_BIN_CAP=16 makes every draw with more than 16 bin entries overflow it: the unbinned kernels
    must then produce the oracle's bits (a performance cliff, never an error)"""
    monkeypatch.setenv("TRB_BIN_CAP", "16")
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


@pytest.mark.parametrize("name", DRAW_PATH_CASES)
def test_synchronous_draws_match_oracle(cuda_api, port_api, monkeypatch, name):
    """TRB_SYNC_DRAWS=1: bins sized exactly after one stream synchronisation per draw (the pre-async path)"""
    monkeypatch.setenv("TRB_SYNC_DRAWS", "1")
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


@pytest.mark.parametrize("env", [{}, {"TRB_WARP_MAX": "0"}, {"TRB_BIN_CAP": "16"}, {"TRB_SYNC_DRAWS": "1"}, {"TRB_DIRECT_AREA": "0"}])
@pytest.mark.parametrize("name", ["indexed_duplicates", "head_small", "orbit_small", "sub_range_draws", "orbit_culled",
                                  "snapshot_restore_twice", "shadow_small", "gouraud_small", "soup_mesh_fp32",
                                  "soup_duplicates_lit"])
def test_mesh_processing_order_matches_oracle(cuda_api, port_api, monkeypatch, name, env):
    """TRB_MESH_ORDER_MIN_TRIS=1 gives EVERY indexed mesh the Morton processing order that only multi-million-triangle
    meshes get by default (mesh_order.cu): set-up, bin fill, both raster kernels, the direct path and the unbinned fallback
    then work on slots and map them to triangle ids through the permutation - the oracle's bits must come out, ties
    included (ids stay those of the index buffer)"""
    monkeypatch.setenv("TRB_MESH_ORDER_MIN_TRIS", "1")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run_case(cuda_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)


def test_pinned_uploads_and_block_recycling(cuda_api):
    """meshes / textures handed over in page-locked memory go to the device by DMA + an interleave kernel
    (no CPU staging); uploading, rendering and freeing the whole scene every frame recycles the
    event-tagged device blocks while earlier frames are still in flight.  Every frame must equal the
    frame rendered from pageable arrays uploaded once."""
    import torch
    sc = scenes.orbit_scene(320, 180, room_quads=((16, 8), (16, 4), (8, 8)), tex_size=64)
    pr = cuda_api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
    frames = [[5 * k + 1, 5 * k + 300, 5 * k + 700] for k in range(6)]
    with trb.Renderer(cuda_api) as r:
        up = scenes.UploadedScene(r, sc)
        want = []
        for ids in frames:
            up.render(scenes.orbit_views(cuda_api, ids), pr)
            want.append([(r.read_depth(v).copy(), r.read_color(v).copy()) for v in range(3)])
        up.free()
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
    pinbuf = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    with trb.Renderer(cuda_api) as r:
        got = []
        for ids in frames:                      # no synchronising call inside the loop
            up = scenes.UploadedScene(r, sc)
            up.render(scenes.orbit_views(cuda_api, ids), pr)
            cs = [pinbuf((180, 320, 3), torch.uint8) for _ in range(3)]
            ds = [pinbuf((180, 320), torch.float64) for _ in range(3)]
            r.readback_async(cs, ds)
            up.free()
            got.append((cs, ds))
        r.readback_wait()
    for k in range(len(frames)):
        for v in range(3):
            assert np.array_equal(got[k][1][v].view(np.uint64), want[k][v][0].view(np.uint64)), (k, v)
            assert np.array_equal(got[k][0][v], want[k][v][1]), (k, v)


@pytest.fixture(scope="module")
def checks_api(built):
    p = os.path.join(os.path.dirname(HERE), "tinyrenderder_b200", "libtrb_checks.so")
    if not os.path.exists(p):
        pytest.skip("libtrb_checks.so not built")
    return trb.Api(p, "trb")


@pytest.mark.parametrize("env", [{}, {"TRB_WARP_MAX": "0"}, {"TRB_BIN_CAP": "16"}, {"TRB_SYNC_DRAWS": "1"},
                                 {"TRB_MESH_ORDER_MIN_TRIS": "1"}, {"TRB_MESH_ORDER_MIN_TRIS": "1", "TRB_BIN_CAP": "16"},
                                 {"TRB_WARP_MAX": "8", "TRB_SPLIT_S": "32"}, {"TRB_WARP_MAX": "8", "TRB_SPLIT_S": "32", "TRB_SPLIT_CAP": "12"}])
@pytest.mark.parametrize("name", ["k7b_small", "big_triangles", "dense_tile", "soup_mesh_fp32", "orbit_small",
                                  "snapshot_restore_twice", "shadow_small"])
def test_kernels_hold_their_indexing_invariants(checks_api, port_api, monkeypatch, name, env):
    """the -DTRB_DEBUG_CHECKS build asserts bin slots, shared-memory tile indices and sample decoding inside
    the kernels (a failed assert traps the kernel and the next call reports the CUDA error); results must
    still equal the oracle's.  compute-sanitizer is not available on the GPU pool."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run_case(checks_api, name)
    want = run_case(port_api, name)
    compare.assert_outputs_match(name, got, want)

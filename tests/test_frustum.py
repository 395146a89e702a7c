"""Model-level frustum culling: the host helpers of the product library must reproduce the
reference's own Frustum / AABB code (our_gl.cpp:212-280, geometry.h:297-327 - including the planes it reads from the
transposed matrix) bit for bit; oracle/_ref answers with the reference's classes themselves."""
import numpy as np
import pytest

import tinyrenderder_b200 as trb
from tinyrenderder_b200 import scenes


@pytest.fixture(scope="module")
def product_api(built):
    return trb.load_cuda()      # host helpers only: no device is touched


def _matrices(api, rng, n):
    out = []
    for i in range(n):
        if i % 3 == 0:            # what main() builds: Perspective * ModelView
            eye, ctr = rng.normal(size=3) * rng.choice([1, 5, 40]), rng.normal(size=3) * 2
            pr = api.perspective(rng.uniform(20, 110), rng.uniform(0.5, 2.5), 0.05, rng.choice([10.0, 500.0]))
            out.append(api.mat4_mul(pr, api.lookat(eye, ctr, [0.0, 1.0, 0.0])))
        elif i % 3 == 1:          # anything at all
            out.append(rng.normal(size=(4, 4)) * rng.choice([1e-3, 1, 1e3]))
        else:                     # rows that vanish: a plane of length zero stays unnormalised (our_gl.cpp:254-258)
            m = rng.normal(size=(4, 4))
            m[:3, 3] = 0.0
            m[:3, rng.integers(0, 3)] = 0.0
            out.append(m)
    return out


def _boxes(rng, n):
    out = []
    for _ in range(n):
        c, h = rng.normal(size=3) * rng.choice([0.5, 5, 50]), np.abs(rng.normal(size=3)) * rng.choice([0.0, 0.1, 3, 30])
        out.append((c - h, c + h))
    return out


@pytest.mark.parametrize("other", ["ref", "port"])
def test_frustum_and_aabb_match_the_reference(product_api, ref_api, port_api, other):
    want_api = ref_api if other == "ref" else port_api
    rng = np.random.default_rng(7)
    mats, boxes = _matrices(product_api, rng, 300), _boxes(rng, 40)
    verdicts = [0, 0]
    for m in mats:
        pg, pw = product_api.frustum_planes(m), want_api.frustum_planes(m)
        assert np.array_equal(pg.view(np.uint64), pw.view(np.uint64))
        for lo, hi in boxes:
            g, w = product_api.frustum_intersects(pg, lo, hi), want_api.frustum_intersects(pw, lo, hi)
            assert g == w
            verdicts[int(g)] += 1
    assert min(verdicts) > 500, verdicts          # both outcomes are exercised
    for m in mats[:60]:
        if abs(m[3]).sum() == 0:
            continue
        for lo, hi in boxes[:10]:
            a, b = product_api.aabb_transform(lo, hi, m), want_api.aabb_transform(lo, hi, m)
            assert np.array_equal(a[0].view(np.uint64), b[0].view(np.uint64))
            assert np.array_equal(a[1].view(np.uint64), b[1].view(np.uint64))


def test_cull_batch_is_what_main_asks_per_model(product_api, ref_api):
    sc = scenes.orbit_scene(320, 180, room_quads=((8, 4), (8, 2), (4, 4)), tex_size=16)
    pr = product_api.perspective(sc.fov, 320 / 180, sc.znear, sc.zfar)
    rng = np.random.default_rng(3)
    cams = [(rng.normal(size=3) * 6, rng.normal(size=3) * 3) for _ in range(100)]
    # cameras for which the reference's (transposed-plane) frustum drops the head, found with the reference itself
    cams += [((0, 1.6, -5), (0, 1.6, -30)), ((0, 1.7, 0.5), (5, 1.7, 0.5)), ((5, 3, 5), (20, 3, 20)), ((0, 5, 0), (3, 5, 0)),
             ((-3.4019, 2.2001, 1.8026), (1.3555, 1.5116, -0.9686))]
    views = np.stack([product_api.lookat(e, c, [0.0, 1.0, 0.0]) for e, c in cams])
    got, want = scenes.visible_items(sc, views, pr, product_api), scenes.visible_items(sc, views, pr, ref_api)
    assert np.array_equal(got, want)
    assert got[1].any() and not got[1].all()             # the head is culled for some of these cameras and kept for others
    assert np.array_equal(got[2], got[1])                # the eyes follow the HEAD's box (main.cpp:706)
    # one camera at a time == Frustum::createFromMatrix(Perspective * ModelView).intersects(box) spelled out
    lo, hi = sc.items[1]._world_aabb
    for v in range(len(cams) - 25, len(cams)):
        planes = ref_api.frustum_planes(ref_api.mat4_mul(pr, views[v]))
        assert ref_api.frustum_intersects(planes, lo, hi) == bool(got[1][v])

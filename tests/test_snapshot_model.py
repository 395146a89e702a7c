"""CPU check of the tile-granular depth snapshot protocol (tests/snapshot_tile_model.py) against the plain
`zbuffer_before_eyes = zbuffer` / `zbuffer = zbuffer_before_eyes` of main.cpp:700, 730 on random call sequences.
The GPU side of the same protocol is tests/test_gpu_parity.py::test_depth_snapshot_variants_match_oracle."""
import numpy as np
import pytest

import snapshot_tile_model as model


def _rects(rng, n, w, h, size):
    out = []
    for _ in range(n):
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        x1, y1 = min(w - 1, x0 + int(rng.integers(0, size))), min(h - 1, y0 + int(rng.integers(0, size)))
        out.append((x0, y0, x1, y1, float(rng.uniform(-1, 1))))
    return out


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("variant", ["tiles", "save_all", "overflow", "whole_plane_collect"])
def test_tile_snapshot_equals_plane_copy(seed, variant):
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(17, 90)), int(rng.integers(17, 70))          # ragged last tile row / column most of the time
    kw = {"tiles": {}, "save_all": {"lazy_max_tris": 0}, "overflow": {}, "whole_plane_collect": {"collect_by_tiles": False}}[variant]
    lazy, plain = model.Frame(w, h, lazy=True, **kw), model.Frame(w, h, lazy=False)
    saved_keys = None
    for _ in range(int(rng.integers(6, 16))):
        op = rng.choice(["draw", "draw", "draw", "flush", "snapshot", "restore"])
        if op == "draw":
            rects = _rects(rng, int(rng.integers(1, 12)), w, h, int(rng.choice([2, 5, 40])))
            over = variant == "overflow" and bool(rng.integers(0, 2))
            lazy.draw(rects, overflow=over)
            plain.draw(rects)
        elif op == "flush":
            lazy.flush()
            plain.flush()
        elif op == "snapshot":
            lazy.snapshot()
            plain.snapshot()
            saved_keys = plain.snap.copy()
        elif saved_keys is not None:
            lazy.restore()
            plain.restore()
        if saved_keys is not None:
            lazy.check_invariant(saved_keys)
        assert np.array_equal(lazy.key, plain.key)
        assert np.array_equal(lazy.vis, plain.vis)
    lazy.flush()
    plain.flush()
    assert np.array_equal(lazy.shaded_by, plain.shaded_by)               # what was shaded, pixel by pixel: the colours


def test_flush_inside_the_window_uses_the_marked_tiles():
    f = model.Frame(64, 48)
    f.draw([(0, 0, 63, 47, 0.5)])
    f.snapshot()
    f.draw([(20, 20, 22, 21, 0.1)])
    f.flush()
    assert f.collected_from_tiles == 1 and int(f.saved.sum()) == 1
    f.restore()
    assert (f.key == 0.5).all() and f.shaded_by[20, 20] == 2

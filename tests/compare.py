"""Comparison rules of the parity tests (north_star): coverage and depth bit-exact, shaded colour
within 1 LSB per channel on at least 99.9 % of the pixels."""
import hashlib

import numpy as np

COLOR_LSB = 1
COLOR_FRACTION = 0.999


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_same(name, got, want, key):
    if isinstance(want, dict):
        for k, v in want.items():
            g = got[k]
            if isinstance(v, float):
                assert (g == v) or (np.isnan(g) and np.isnan(v)), "%s/%s.%s: %r != %r" % (name, key, k, g, v)
            else:
                assert int(g) == int(v), "%s/%s.%s: %r != %r" % (name, key, k, g, v)
        return
    assert got.shape == want.shape and got.dtype == want.dtype, "%s/%s: shape/dtype" % (name, key)
    if key.startswith(("bgr", "final")):
        d = np.abs(got.astype(np.int32) - want.astype(np.int32)).max(axis=-1)
        frac = float((d <= COLOR_LSB).mean()) if d.size else 1.0
        assert frac >= COLOR_FRACTION, "%s/%s: only %.5f of the pixels within %d LSB (max diff %d)" % (
            name, key, frac, COLOR_LSB, int(d.max()))
    else:  # depth (compared as bits so that -0.0 / +0.0 and inf count), ao, z image: exact
        a = got.view(np.uint64) if got.dtype == np.float64 else got
        b = want.view(np.uint64) if want.dtype == np.float64 else want
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s/%s: %d elements differ, first at %s: got %r want %r" % (
                name, key, len(bad), bad[0].tolist(), got[tuple(bad[0])], want[tuple(bad[0])]))


def assert_outputs_match(name, got, want, skip=()):
    for key, w in want.items():
        if any(key.startswith(s) for s in skip):
            continue
        assert key in got, "%s: missing output %s" % (name, key)
        assert_same(name, got[key], w, key)


def assert_matches_golden(name, got, gold, skip=()):
    for key, w in gold.items():
        if any(key.startswith(s) for s in skip):
            continue
        g = got[key]
        if isinstance(w, dict) and "sha256" in w:
            assert list(g.shape) == w["shape"] and str(g.dtype) == w["dtype"], "%s/%s: shape/dtype" % (name, key)
            assert sha(g) == w["sha256"], "%s/%s: differs from the committed golden vector" % (name, key)
        else:
            assert_same(name, g, w, key)

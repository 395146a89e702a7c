"""Integer identities the kernels rely on, checked exhaustively on the CPU (no GPU needed):
the reciprocal trick that turns a sample index into (row, column) in k_raster / k_raster_warp, the
window bookkeeping that maps a lane to its triangle, and the 2-bit map algebra of the RLE scan."""
import itertools

import numpy as np


def test_sample_index_to_row_by_reciprocal():
    """row = l / bw computed as (l * ceil(2^15 / bw)) >> 15 for every clipped-bbox width and sample index
    a 16x16 tile can produce (kernels.cuh: SpEntry.inv, the `pack` word of k_raster_warp)"""
    for bw in range(1, 17):
        inv = (32768 + bw - 1) // bw
        assert inv < (1 << 16)
        for l in range(256):
            assert (l * inv) >> 15 == l // bw, (bw, l)


def test_pack_word_round_trips():
    """x0 | y0 << 4 | (bw - 1) << 8 | inv << 12 fits 32 bits and unpacks to what went in"""
    for x0, y0, bw in itertools.product(range(16), range(16), range(1, 17)):
        inv = (32768 + bw - 1) // bw
        pk = x0 | (y0 << 4) | ((bw - 1) << 8) | (inv << 12)
        assert pk < (1 << 32)
        assert (pk & 15, (pk >> 4) & 15, ((pk >> 8) & 15) + 1, pk >> 12) == (x0, y0, bw, inv)


def test_lane_to_triangle_lookup():
    """k_raster_warp: with the samples of a batch laid end to end, lane L of the window starting at `bs`
    belongs to triangle  before + popc(starts & mask_le(L)) - 1,  where `starts` marks the triangles that
    begin inside the window and `before` counts those that began earlier"""
    rng = np.random.default_rng(1)
    for _ in range(300):
        n = int(rng.integers(1, 33))
        ns = rng.integers(1, 257, n)                      # every binned triangle has at least one sample
        first = np.concatenate([[0], np.cumsum(ns)[:-1]])
        total = int(ns.sum())
        owner = np.repeat(np.arange(n), ns)
        for bs in range(0, total, 32):
            starts = 0
            for f in first:
                if bs <= f < bs + 32:
                    starts |= 1 << int(f - bs)
            before = int((first < bs).sum())
            for lane in range(32):
                s = bs + lane
                if s >= total:
                    break
                e = before + bin(starts & (0xFFFFFFFF >> (31 - lane))).count("1") - 1
                assert e == owner[s]
            single = (starts & ~1) == 0                  # the fast path: the whole window is one triangle
            assert single == (len(set(owner[bs:min(bs + 32, total)])) == 1)


def compose(a, b):
    return ((b >> (a & 1)) & 1) | (((b >> ((a >> 1) & 1)) & 1) << 1)


def test_rle_map_composition_is_an_associative_monoid():
    """tga_rle.cuh MapCompose: 2-bit maps on {P, R}; the scan needs associativity and an identity"""
    maps = range(4)
    for a, b, c in itertools.product(maps, maps, maps):
        assert compose(compose(a, b), c) == compose(a, compose(b, c))
    for a in maps:
        assert compose(2, a) == a and compose(a, 2) == a
        for x in (0, 1):                                  # the encoding really is "a, then b"
            for b in maps:
                assert (compose(a, b) >> x) & 1 == (b >> ((a >> x) & 1)) & 1


def test_depth_key_is_order_preserving():
    """exact.cuh fragment_key: -0.0 is canonicalised to +0.0 (the reference's `<` treats them as equal), then
    unsigned order of the keys == numeric order of the doubles; +inf is the clear value"""
    rng = np.random.default_rng(2)
    v = np.concatenate([rng.normal(size=2000), [0.0, -0.0, np.inf, -np.inf, 1e-310, -1e-310, 1.0, -1.0]])
    v = v + 0.0                                           # -0.0 -> +0.0
    bits = v.view(np.uint64)
    sign = np.uint64(1) << np.uint64(63)
    keys = np.where(bits & sign, ~bits, bits | sign)
    order = np.argsort(v, kind="stable")
    ks = keys[order]
    vs = v[order]
    assert all(ks[i] <= ks[i + 1] for i in range(len(ks) - 1))
    assert all((ks[i] < ks[i + 1]) == (vs[i] < vs[i + 1]) or (vs[i] == vs[i + 1]) for i in range(len(ks) - 1))
    assert int(keys[list(v).index(np.inf)]) == 0xFFF0000000000000

"""numpy model of the parallel formulation in tinyrenderder_b200/csrc/tga_rle.cuh (same passes, same
formulas, loops instead of grids) - lets the CPU suite check the decomposition itself against the
sequential packetiser before any GPU is involved."""
import numpy as np


def seg_info(s, e, nxt, next_is_long, entry_r):
    L = e - s
    c = entry_r if L else 0
    X = L - c
    m = X & 127
    p0 = s + c
    rle_px = X - (1 if m == 1 else 0)
    rle_packets = (X >> 7) + (1 if m >= 2 else 0)
    r0 = p0 + rle_px
    cnt = nxt - r0
    exit_r = (cnt & 127) != 0
    nraw = cnt + (1 if (next_is_long and exit_r) else 0)
    return p0, r0, rle_px, rle_packets, nraw, exit_r


def compose(a, b):
    return ((b >> (a & 1)) & 1) | (((b >> ((a >> 1) & 1)) & 1) << 1)


def encode_views(px, npix, nviews, bpp):
    """px: (nviews*npix, bpp) uint8.  Returns list of per-view packet streams (bytes, no header)."""
    total = npix * nviews
    r = np.arange(total) % npix
    same_prev = np.zeros(total, dtype=bool)
    same_prev[1:] = (px[1:] == px[:-1]).all(axis=1)
    eq = same_prev & (r != 0)
    eq_next = np.zeros(total, dtype=bool)
    eq_next[:-1] = same_prev[1:]
    eq_next &= (r + 1 != npix)
    ls = (~eq) & eq_next
    ls_excl = np.concatenate([[0], np.cumsum(ls)[:-1]]).astype(np.int64)
    ls_total = int(ls.sum())
    ns = ls_total + nviews
    seg_start = np.full(ns + 1, -1, dtype=np.int64)
    long_end = np.full(ns + 1, -1, dtype=np.int64)
    for i in range(total):
        v = i // npix
        ex = ls_excl[i]
        if ls[i]:
            seg_start[ex + 1 + v] = i
        if r[i] == 0:
            seg_start[ex + v] = i
            long_end[ex + v] = i
        elif (not eq[i]) and eq[i - 1]:
            long_end[ex + v] = i
        if r[i] + 1 == npix and eq[i]:
            long_end[ex + v] = i + 1
    seg_start[ns] = total
    long_end[ns] = total
    assert (seg_start >= 0).all() and (long_end >= 0).all()
    maps = np.empty(ns, dtype=np.int64)
    for k in range(ns):
        nl = long_end[k + 1] != seg_start[k + 1]
        m0 = seg_info(seg_start[k], long_end[k], seg_start[k + 1], nl, 0)[5]
        m1 = seg_info(seg_start[k], long_end[k], seg_start[k + 1], nl, 1)[5]
        maps[k] = int(m0) | (int(m1) << 1)
    prefix = np.empty(ns, dtype=np.int64)
    acc = 2
    for k in range(ns):
        prefix[k] = acc
        acc = compose(acc, maps[k])
    nbytes = np.empty(ns + 1, dtype=np.int64)
    nbytes[ns] = 0
    for k in range(ns):
        nl = long_end[k + 1] != seg_start[k + 1]
        g = seg_info(seg_start[k], long_end[k], seg_start[k + 1], nl, prefix[k] & 1)
        nbytes[k] = g[3] * (1 + bpp) + (g[4] + 127) // 128 + g[4] * bpp
    base = np.concatenate([[0], np.cumsum(nbytes)[:-1]])
    out = np.full(int(nbytes.sum()), 0xEE, dtype=np.uint8)
    written = np.zeros(out.size, dtype=bool)
    for i in range(total):
        v = i // npix
        k = ls_excl[i] + int(ls[i]) + v
        s, e = seg_start[k], long_end[k]
        if i == s and e != s and (prefix[k] & 1):
            k -= 1
            s, e = seg_start[k], long_end[k]
        nxt = seg_start[k + 1]
        p0, r0, rle_px, rle_packets, nraw, _ = seg_info(s, e, nxt, long_end[k + 1] != nxt, prefix[k] & 1)
        o = base[k]
        if i < r0:
            d = i - p0
            if d & 127:
                continue
            q = d >> 7
            ln = min(128, rle_px - (q << 7))
            o += q * (1 + bpp)
            out[o] = 128 + ln - 1
            out[o + 1:o + 1 + bpp] = px[i]
            written[o:o + 1 + bpp] = True
        else:
            d = i - r0
            q, w = d >> 7, d & 127
            o += rle_packets * (1 + bpp) + q * (1 + 128 * bpp) + 1 + w * bpp
            if w == 0:
                out[o - 1] = min(128, nraw - (q << 7)) - 1
                written[o - 1] = True
            out[o:o + bpp] = px[i]
            written[o:o + bpp] = True
    assert written.all(), "holes in the packet stream"
    res = []
    for v in range(nviews):
        a = base[ls_excl[v * npix] + v]
        b = base[ls_excl[(v + 1) * npix] + v + 1] if v + 1 < nviews else base[ns]
        res.append(out[a:b].tobytes())
    return res

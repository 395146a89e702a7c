"""numpy model of the tile-granular depth snapshot in tinyrenderder_b200/csrc (trb.cu: trb_depth_snapshot /
raster_draw / do_flush / trb_depth_restore, kernels.cuh: k_snap_save / k_snap_restore / k_shade_collect_tiles): the same
protocol with loops instead of grids, so that the CPU suite can check the protocol itself - "a tile whose byte is 0
still holds the snapshot's keys", "every unshaded pixel of a flush inside the window lies in a marked tile" - against the
plain `saved = zbuffer` ... `zbuffer = saved` of main.cpp:700, 730 before any GPU is involved."""
import numpy as np

TILE = 16
NONE, SHADED = 0xFFFFFFFF, 0


class Frame:
    """one view; depth keys as floats (smaller wins), ids as in the id plane (NONE / SHADED / triangle id + 1)"""

    def __init__(self, w, h, lazy=True, lazy_max_tris=1 << 20, collect_by_tiles=True):
        self.w, self.h = w, h
        self.tw, self.th = (w + TILE - 1) // TILE, (h + TILE - 1) // TILE
        self.key = np.full((h, w), np.inf)
        self.vis = np.full((h, w), NONE, dtype=np.uint32)
        self.shaded_by = np.zeros((h, w), dtype=np.uint32)      # stands for the colour: id that was shaded last
        self.next_id = 0
        self.lazy, self.lazy_max_tris, self.collect_by_tiles = lazy, lazy_max_tris, collect_by_tiles
        self.have_snapshot = self.snap_lazy = self.snap_all = False
        self.snap = np.zeros((h, w))
        self.saved = np.zeros(self.tw * self.th, dtype=np.uint8)
        self.collected_from_tiles = 0

    # -- helpers --------------------------------------------------------------------------------
    def _tile_rect(self, t):
        tx, ty = t % self.tw, t // self.tw
        return slice(ty * TILE, min((ty + 1) * TILE, self.h)), slice(tx * TILE, min((tx + 1) * TILE, self.w))

    def _window(self):
        return self.have_snapshot and self.snap_lazy and not self.snap_all

    def _save(self, counts, overflow):      # k_snap_save
        every = counts is None or overflow
        for t in range(self.tw * self.th):
            if not self.saved[t] and (every or counts[t]):
                self.saved[t] = 1
                ys, xs = self._tile_rect(t)
                self.snap[ys, xs] = self.key[ys, xs]
        if counts is None:
            self.snap_all = True

    # -- the calls ------------------------------------------------------------------------------
    def draw(self, rects, overflow=False):
        """rects: (x0, y0, x1, y1, z) axis-aligned 'triangles' (inclusive pixel boxes); the bins count bbox tiles"""
        save_tiles = self._window()
        if save_tiles and len(rects) > self.lazy_max_tris:
            self._save(None, False)
            save_tiles = False
        counts = np.zeros(self.tw * self.th, dtype=np.uint32)
        for x0, y0, x1, y1, _ in rects:
            for ty in range(y0 // TILE, y1 // TILE + 1):
                for tx in range(x0 // TILE, x1 // TILE + 1):
                    counts[ty * self.tw + tx] += 1
        if save_tiles:
            self._save(counts, overflow)
        for x0, y0, x1, y1, z in rects:
            self.next_id += 1
            k, v = self.key[y0:y1 + 1, x0:x1 + 1], self.vis[y0:y1 + 1, x0:x1 + 1]
            win = z < k                       # a later fragment that ties loses
            k[win] = z
            v[win] = self.next_id

    def flush(self):
        pending = (self.vis != NONE) & (self.vis != SHADED)
        if self._window() and self.collect_by_tiles:     # k_shade_collect_tiles
            listed = np.zeros_like(pending)
            for t in np.flatnonzero(self.saved):
                ys, xs = self._tile_rect(t)
                listed[ys, xs] = pending[ys, xs]
            self.collected_from_tiles += 1
            assert (listed == pending).all(), "an unshaded pixel lies outside the marked tiles"
            pending = listed
        self.shaded_by[pending] = self.vis[pending]
        self.vis[pending] = SHADED

    def snapshot(self):
        self.flush()
        self.snap_lazy, self.snap_all = self.lazy, False
        if self.snap_lazy:
            self.saved[:] = 0
        else:
            self.snap[:] = self.key
        self.have_snapshot = True

    def restore(self):
        assert self.have_snapshot
        self.flush()
        if self.snap_lazy:                   # k_snap_restore
            for t in np.flatnonzero(self.saved):
                ys, xs = self._tile_rect(t)
                self.key[ys, xs] = self.snap[ys, xs]
        else:
            self.key[:] = self.snap

    def check_invariant(self, reference_snapshot):
        """a tile whose byte is 0 still holds the snapshot's keys; a marked tile has them in the spare plane"""
        if not (self.have_snapshot and self.snap_lazy):
            return
        for t in range(self.tw * self.th):
            ys, xs = self._tile_rect(t)
            got = self.snap[ys, xs] if self.saved[t] else self.key[ys, xs]
            assert np.array_equal(got, reference_snapshot[ys, xs]), "tile %d lost the snapshot" % t

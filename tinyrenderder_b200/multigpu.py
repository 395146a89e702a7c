"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collectives.

* config 3 (camera orbit): frames are independent -> `frame_shard` hands rank r its frames; no
  data-path collective at all.
* config 4 (huge mesh): triangle ranges per rank + sort-last composite of the (depth key, id)
  planes: two MIN all-reduces over NVLink, exact because keys are order preserving and ids are
  global submission indices.
"""
import numpy as np


def frame_shard(total_frames, rank, world):
    """Contiguous block of frames for `rank` (first ranks get the remainder)."""
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def triangle_shard(ntris, rank, world):
    """[first, first+count) of the mesh's triangles for `rank`; ids stay global."""
    base, rem = divmod(ntris, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def row_shard(height, rank, world):
    base, rem = divmod(height, world)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)


class _DevPtr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def plane_tensors(renderer):
    """torch views (zero copy) of the context's depth-key plane as int64 and id plane as int32."""
    import torch
    kp, vp, n = renderer.device_planes()
    key = torch.as_tensor(_DevPtr(kp, n, "<i8"), device="cuda")
    vid = torch.as_tensor(_DevPtr(vp, n, "<i4"), device="cuda")
    return key, vid


def composite(renderer, all_reduce_min):
    """Sort-last composite of the calling rank's unflushed frame with its peers.
    `all_reduce_min(tensor)` reduces in place over the ranks (dist.all_reduce(op=MIN) in production,
    an emulation in the single-GPU test)."""
    import torch
    renderer.composite_save_local_depth()
    key, vid = plane_tensors(renderer)
    all_reduce_min(key)
    torch.cuda.synchronize()
    renderer.composite_mask()
    all_reduce_min(vid)
    torch.cuda.synchronize()
    renderer.composite_finish()


class P2PComposite:
    """Fused NVLink composite + shade (config 4): peers' planes are opened once through CUDA IPC,
    then every frame costs two host barriers and ONE kernel per rank (trb_composite_shade_p2p)."""

    def __init__(self, renderer, dist, rank, world):
        self.r, self.dist, self.rank, self.world = renderer, dist, rank, world
        self.key = None

    def _open(self):
        planes = self.r.device_planes()[:2]
        # whether to re-exchange handles is decided COLLECTIVELY: a rank whose planes were reallocated
        # (DevBuf growth, depth restore by pointer swap) must not enter the all-gather alone
        changed = [None] * self.world
        self.dist.all_gather_object(changed, planes != self.key)
        if not any(changed):
            return
        mine = self.r.ipc_export_planes()
        every = [None] * self.world
        self.dist.all_gather_object(every, mine)
        self.r.ipc_open_peers([h[0] for h in every], [h[1] for h in every], self.rank)
        self.key = planes

    def run(self, height):
        """call after the rank's draws of the frame (no flush); shades the rows this rank owns"""
        self._open()                      # (re)exchange handles when the planes were (re)allocated
        self.r.synchronize()              # my draws are complete ...
        self.dist.barrier()               # ... and so are everybody else's
        y0, y1 = row_shard(height, self.rank, self.world)
        self.r.composite_shade_p2p(y0, y1)
        self.dist.barrier()               # nobody clears its planes while a peer still reads them
        return y0, y1


class CommComposite:
    """The composite group of the C ABI (trb_comm_export / trb_comm_open / trb_composite) for one process per rank:
    the IPC blobs are exchanged ONCE with the host-side collective `all_gather(obj) -> list`; after that a frame costs no
    host synchronisation at all - every rank's stream publishes and polls frame counters in device memory."""

    def __init__(self, renderer, all_gather, rank, world):
        self.r, self.all_gather, self.rank, self.world = renderer, all_gather, rank, world
        self.opened = False

    def open(self):
        """call once, after the first begin_frame of the final size (the planes are exported as they are then)"""
        blobs = self.all_gather(self.r.comm_export())
        self.r.comm_open(blobs, self.rank)
        self.opened = True

    def shard(self, ntris):
        return self.r.comm_shard(ntris) if self.opened else triangle_shard(ntris, self.rank, self.world)

    def run(self):
        """after the rank's draws of the frame (no flush): composite + shade the owned rows; returns them"""
        if not self.opened:
            self.open()
        self.r.composite()
        return self.r.comm_rows()


def composite_depth_color_cpu(z_list, bgr_list):
    """The same protocol on host arrays (used by the gloo tests with the CPU oracle): the winner of a
    pixel is the rank with the smallest depth, lowest rank on ties (lower ranks hold lower ids)."""
    z = np.stack(z_list)
    win = np.argmin(z, axis=0)  # first minimum = lowest rank
    zc = np.take_along_axis(z, win[None], 0)[0]
    c = np.stack(bgr_list)
    cc = np.take_along_axis(c, win[None, :, :, None], 0)[0]
    return zc, cc

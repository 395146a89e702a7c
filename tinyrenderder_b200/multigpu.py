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


def composite_depth_color_cpu(z_list, bgr_list):
    """The same protocol on host arrays (used by the gloo tests with the CPU oracle): the winner of a
    pixel is the rank with the smallest depth, lowest rank on ties (lower ranks hold lower ids)."""
    z = np.stack(z_list)
    win = np.argmin(z, axis=0)  # first minimum = lowest rank
    zc = np.take_along_axis(z, win[None], 0)[0]
    c = np.stack(bgr_list)
    cc = np.take_along_axis(c, win[None, :, :, None], 0)[0]
    return zc, cc

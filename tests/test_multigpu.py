"""Multi-GPU logic.  CPU: world_size-2 gloo processes exercise the frame sharding of
config 3 and the sort-last composite protocol of config 4 on the CPU oracle.  GPU: the composite
entry points of libtrb.so with two 'ranks' emulated as two contexts on one device."""
import os
import sys

import numpy as np
import pytest

import tinyrenderder_b200 as trb
from tinyrenderder_b200 import multigpu, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_the_work():
    for total, world in ((1024, 8), (1000, 7), (5, 8), (0, 3)):
        seen = []
        for r in range(world):
            seen += list(multigpu.frame_shard(total, r, world))
        assert seen == list(range(total))
    for n, world in ((20971520, 8), (13, 4), (3, 5)):
        first = 0
        for r in range(world):
            f, c = multigpu.triangle_shard(n, r, world)
            assert f == first
            first += c
        assert first == n
    rows = [multigpu.row_shard(2160, r, 8) for r in range(8)]
    assert rows[0][0] == 0 and rows[-1][1] == 2160 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


def _gloo_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    api = trb.Api(os.path.join(ROOT, "oracle", "libtrb_port.so"), "orc")
    # --- config 3: every rank renders its own frames, nothing is exchanged but the final gather
    sc = scenes.orbit_scene(160, 90, room_quads=((8, 4), (8, 2), (4, 4)), head_res=(10, 8), eye_res=(6, 4), tex_size=32)
    pr = api.perspective(sc.fov, 160 / 90, sc.znear, sc.zfar)
    frames = list(multigpu.frame_shard(6, rank, world))
    with trb.Renderer(api) as r:
        up = scenes.UploadedScene(r, sc)
        up.render(scenes.orbit_views(api, frames, total=6), pr)
        mine = np.stack([r.read_depth(v) for v in range(len(frames))])
    gathered = [torch.zeros(3, 90, 160, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    # --- config 4: triangle ranges + sort-last composite (MIN all-reduce of depth, lowest rank on ties)
    m = scenes.icosphere(3)
    mv, pr4 = scenes.sphere_view(api), api.perspective(60, 1.5, 0.1, 10)
    first, count = multigpu.triangle_shard(m.ntris, rank, world)
    with trb.Renderer(api) as r:
        h = r.upload_mesh(m.pos, m.nrm, m.uv, m.idx)
        r.begin_frame(150, 100)
        r.draw(h, mv, pr4, first_tri=first, ntris=count)
        r.end_frame()
        z, c = r.read_depth(), r.read_color()
    zt = torch.from_numpy(z.copy())
    dist.all_reduce(zt, op=dist.ReduceOp.MIN)
    owner = torch.from_numpy(np.where(z == zt.numpy(), rank, world).astype(np.int32))
    dist.all_reduce(owner, op=dist.ReduceOp.MIN)             # lowest rank among the winners
    col = torch.from_numpy(np.where((owner.numpy() == rank)[..., None], c, 0).astype(np.int32))
    dist.all_reduce(col, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), frames=torch.cat(gathered).numpy(), z=zt.numpy(),
                 c=col.numpy().astype(np.uint8))
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path, port_api):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out = np.load(tmp_path / "out.npz")
    api = port_api
    # config 3: the gathered frames are the frames one process renders alone
    sc = scenes.orbit_scene(160, 90, room_quads=((8, 4), (8, 2), (4, 4)), head_res=(10, 8), eye_res=(6, 4), tex_size=32)
    with trb.Renderer(api) as r:
        up = scenes.UploadedScene(r, sc)
        up.render(scenes.orbit_views(api, range(6), total=6), api.perspective(sc.fov, 160 / 90, sc.znear, sc.zfar))
        for v in range(6):
            assert np.array_equal(out["frames"][v].view(np.uint64), r.read_depth(v).view(np.uint64))
    # config 4: the composite equals the unsharded render, depth bits and colours
    m = scenes.icosphere(3)
    with trb.Renderer(api) as r:
        h = r.upload_mesh(m.pos, m.nrm, m.uv, m.idx)
        r.begin_frame(150, 100)
        r.draw(h, scenes.sphere_view(api), api.perspective(60, 1.5, 0.1, 10), ntris=m.ntris)
        r.end_frame()
        assert np.array_equal(out["z"].view(np.uint64), r.read_depth().view(np.uint64))
        assert np.array_equal(out["c"], r.read_color())


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 3])
def test_sort_last_composite_on_one_gpu(cuda_api, port_api, nranks):
    """N 'ranks' = N contexts on one GPU; the all-reduce is emulated by an element-wise minimum over
    the ranks' planes (same result as NCCL MIN).  Composite + per-rank row shading == one render."""
    import torch
    m = scenes.icosphere(5)
    w, h = 640, 400
    mv, pr = scenes.sphere_view(cuda_api), cuda_api.perspective(60, w / h, 0.1, 10)
    # duplicate the mesh's triangles so that exact depth ties ACROSS ranks exist
    idx = np.concatenate([m.idx, m.idx])
    ntris = idx.size // 3
    rs = [trb.Renderer(cuda_api) for _ in range(nranks)]
    planes = []
    for rank, r in enumerate(rs):
        mesh = r.upload_mesh(m.pos, m.nrm, m.uv, idx)
        first, count = multigpu.triangle_shard(ntris, rank, nranks)
        r.begin_frame(w, h)
        r.set_triangle_id_base(first)
        r.draw(mesh, mv, pr, first_tri=first, ntris=count)
        r.composite_save_local_depth()
        planes.append(multigpu.plane_tensors(r))
    gmin = torch.stack([p[0] for p in planes]).min(dim=0).values
    for p in planes:
        p[0].copy_(gmin)
    torch.cuda.synchronize()
    for r in rs:
        r.composite_mask()
    imin = torch.stack([p[1] for p in planes]).min(dim=0).values
    for p in planes:
        p[1].copy_(imin)
    torch.cuda.synchronize()
    color = np.zeros((h, w, 3), np.uint8)
    depth = np.zeros((h, w))
    for rank, r in enumerate(rs):
        r.composite_finish()
        y0, y1 = multigpu.row_shard(h, rank, nranks)
        r.set_shade_rows(y0, y1)
        r.flush()
        color[y0:y1] = r.read_color()[y0:y1]
        depth[y0:y1] = r.read_depth()[y0:y1]
    with trb.Renderer(port_api) as o:
        mesh = o.upload_mesh(m.pos, m.nrm, m.uv, idx)
        o.begin_frame(w, h)
        o.draw(mesh, mv, pr, ntris=ntris)
        o.end_frame()
        assert np.array_equal(depth.view(np.uint64), o.read_depth().view(np.uint64))
        assert np.array_equal(color, o.read_color())
    for r in rs:
        r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4])
def test_fused_p2p_composite_shade_on_one_gpu(cuda_api, port_api, nranks):
    """the fused composite+shade kernel (trb_composite_shade_p2p) with N contexts of one process as
    ranks (raw peer pointers instead of CUDA IPC): equals the unsharded render bit for bit"""
    m = scenes.icosphere(5)
    w, h = 640, 400
    mv, pr = scenes.sphere_view(cuda_api), cuda_api.perspective(60, w / h, 0.1, 10)
    idx = np.concatenate([m.idx, m.idx])          # duplicates: exact depth ties across ranks
    ntris = idx.size // 3
    rs = [trb.Renderer(cuda_api) for _ in range(nranks)]
    tex = scenes.texture_diffuse(64, 3)
    for rank, r in enumerate(rs):
        mesh = r.upload_mesh(m.pos, m.nrm, m.uv, idx)
        first, count = multigpu.triangle_shard(ntris, rank, nranks)
        u = trb.PhongUniforms()
        u.key_dir_eye[:] = cuda_api.light_dir_eye(mv, scenes.normalized(scenes.KEY_LIGHT))
        u.fill_dir_eye[:] = cuda_api.light_dir_eye(mv, scenes.normalized(scenes.FILL_LIGHT))
        u.rim_dir_eye[:] = cuda_api.light_dir_eye(mv, scenes.normalized(scenes.RIM_LIGHT))
        u.normal_map_strength = 0.0
        u.diffuse = r.upload_texture(tex)
        r.begin_frame(w, h)
        r.set_triangle_id_base(first)
        r.draw(mesh, mv, pr, kind=trb.SHADER_PHONG, uniforms=u, first_tri=first, ntris=count)
        r.synchronize()
    planes = [r.device_planes() for r in rs]
    color = np.zeros((h, w, 3), np.uint8)
    depth = np.zeros((h, w))
    for rank, r in enumerate(rs):
        r.open_peers_raw([p[0] for p in planes], [p[1] for p in planes], rank)
    for rank, r in enumerate(rs):     # every rank composites + shades its rows; nobody has cleared yet
        y0, y1 = multigpu.row_shard(h, rank, nranks)
        r.composite_shade_p2p(y0, y1)
        color[y0:y1] = r.read_color()[y0:y1]
        depth[y0:y1] = r.read_depth()[y0:y1]
    with trb.Renderer(port_api) as o:
        mesh = o.upload_mesh(m.pos, m.nrm, m.uv, idx)
        u = trb.PhongUniforms()
        u.key_dir_eye[:] = port_api.light_dir_eye(mv, scenes.normalized(scenes.KEY_LIGHT))
        u.fill_dir_eye[:] = port_api.light_dir_eye(mv, scenes.normalized(scenes.FILL_LIGHT))
        u.rim_dir_eye[:] = port_api.light_dir_eye(mv, scenes.normalized(scenes.RIM_LIGHT))
        u.normal_map_strength = 0.0
        u.diffuse = o.upload_texture(tex)
        o.begin_frame(w, h)
        o.draw(mesh, mv, pr, kind=trb.SHADER_PHONG, uniforms=u, ntris=ntris)
        o.end_frame()
        assert np.array_equal(depth.view(np.uint64), o.read_depth().view(np.uint64))
        assert np.abs(color.astype(int) - o.read_color().astype(int)).max() <= 1
    for r in rs:
        r.close()


def _shard_frame(api, r, mesh, ntris, first, count, mv, pr, tex_handle):
    u = trb.PhongUniforms()
    u.key_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.KEY_LIGHT))
    u.fill_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.FILL_LIGHT))
    u.rim_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.RIM_LIGHT))
    u.normal_map_strength = 0.0
    u.diffuse = tex_handle
    r.set_triangle_id_base(first)
    r.draw(mesh, mv, pr, kind=trb.SHADER_PHONG, uniforms=u, first_tri=first, ntris=count)


def _unsharded(port_api, m, idx, ntris, mv, pr, tex, w, h):
    with trb.Renderer(port_api) as o:
        mesh = o.upload_mesh(m.pos, m.nrm, m.uv, idx)
        o.begin_frame(w, h)
        _shard_frame(port_api, o, mesh, ntris, 0, ntris, mv, pr, o.upload_texture(tex))
        o.end_frame()
        return o.read_depth().copy(), o.read_color().copy()


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4])
def test_composite_group_on_one_gpu(cuda_api, port_api, nranks):
    """trb_comm_init + trb_composite_group: N contexts of one process (sharing the GPU) as ranks, three frames in a row
    with a moving camera and NO host synchronisation between them - the streams are ordered by events: a rank's clear of
    frame k+1 waits for every peer's composite of frame k.  Each frame equals the unsharded render."""
    m = scenes.icosphere(5)
    w, h = 640, 400
    pr = cuda_api.perspective(60, w / h, 0.1, 10)
    idx = np.concatenate([m.idx, m.idx])          # duplicates: exact depth ties across ranks
    ntris = idx.size // 3
    tex = scenes.texture_diffuse(64, 3)
    rs = [trb.Renderer(cuda_api) for _ in range(nranks)]
    meshes = [r.upload_mesh(m.pos, m.nrm, m.uv, idx) for r in rs]
    texes = [r.upload_texture(tex) for r in rs]
    for r in rs:
        r.begin_frame(w, h)
    trb.comm_init(rs)
    cams = [[0.0, 0.0, 2.2], [0.7, 0.3, 2.0], [-0.5, 0.9, 1.9]]
    outs = []
    for cam in cams:
        mv = cuda_api.lookat(cam, [0, 0, 0], [0, 1, 0])
        for rank, r in enumerate(rs):
            r.begin_frame(w, h)
            first, count = r.comm_shard(ntris)
            assert (first, count) == multigpu.triangle_shard(ntris, rank, nranks)
            _shard_frame(cuda_api, r, meshes[rank], ntris, first, count, mv, pr, texes[rank])
        trb.composite_group(rs)
        # read-backs are queued behind the composite on each rank's stream; nothing else synchronises
        color = np.zeros((h, w, 3), np.uint8)
        depth = np.zeros((h, w))
        for rank, r in enumerate(rs):
            y0, y1 = r.comm_rows()
            assert (y0, y1) == multigpu.row_shard(h, rank, nranks)
            color[y0:y1] = r.read_color()[y0:y1]
            depth[y0:y1] = r.read_depth()[y0:y1]
        outs.append((depth, color))
    for cam, (depth, color) in zip(cams, outs):
        zo, co = _unsharded(port_api, m, idx, ntris, port_api.lookat(cam, [0, 0, 0], [0, 1, 0]), pr, tex, w, h)
        assert np.array_equal(depth.view(np.uint64), zo.view(np.uint64))
        assert np.abs(color.astype(int) - co.astype(int)).max() <= 1
    with pytest.raises(trb.TrbError):             # contexts that share a device must be composited as a group
        rs[0].composite()
    with pytest.raises(trb.TrbError):             # the frame size is part of the group
        rs[0].begin_frame(w // 2, h)
    for r in rs:
        r.close()


def _phong(api, mv, tex_handle):
    u = trb.PhongUniforms()
    u.key_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.KEY_LIGHT))
    u.fill_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.FILL_LIGHT))
    u.rim_dir_eye[:] = api.light_dir_eye(mv, scenes.normalized(scenes.RIM_LIGHT))
    u.normal_map_strength = 0.0
    u.diffuse = tex_handle
    return u


@pytest.mark.gpu
@pytest.mark.parametrize("order_min_tris", ["1", "0"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_draw_shard_composite_group_on_one_gpu(cuda_api, port_api, monkeypatch, nranks, order_min_tris):
    """trb_draw_shard: every rank submits the WHOLE mesh and rasterises its share (blocks of the mesh's processing order
    when it has one - TRB_MESH_ORDER_MIN_TRIS=1 forces it for this small mesh, the last block is partial - else a contiguous
    range), ids global without trb_set_triangle_id_base, two draws per frame.  After trb_composite_group every frame equals
    the unsharded render of the oracle, ties across ranks included."""
    monkeypatch.setenv("TRB_MESH_ORDER_MIN_TRIS", order_min_tris)
    # the masked vertex stage of a share (and the composite's on-demand vertex records) normally starts at 4 ranks
    monkeypatch.setenv("TRB_SHARE_VERTEX_MIN_RANKS", "2")
    m = scenes.icosphere(5)
    w, h = 640, 400
    pr = cuda_api.perspective(60, w / h, 0.1, 10)
    idx = np.concatenate([m.idx, m.idx, m.idx[:300]])      # duplicates: exact depth ties across ranks; 41 060 triangles
    ntris = idx.size // 3
    tex = scenes.texture_diffuse(64, 3)
    shift = np.eye(4)
    shift[0, 3], shift[2, 3] = 0.9, -0.8                   # a second copy of the sphere, partly behind the first
    rs = [trb.Renderer(cuda_api) for _ in range(nranks)]
    meshes = [r.upload_mesh(m.pos, m.nrm, m.uv, idx) for r in rs]
    texes = [r.upload_texture(tex) for r in rs]
    for r in rs:
        r.begin_frame(w, h)
    trb.comm_init(rs)
    cams = [[0.0, 0.0, 2.6], [0.7, 0.3, 2.4]]
    for cam in cams:
        view = cuda_api.lookat(cam, [0, 0, 0], [0, 1, 0])
        mvs = [view, cuda_api.mat4_mul(view, shift)]
        for rank, r in enumerate(rs):
            r.begin_frame(w, h)
            for mv in mvs:
                r.draw_shard(meshes[rank], mv, pr, rank, nranks, kind=trb.SHADER_PHONG, uniforms=_phong(cuda_api, mv, texes[rank]))
        trb.composite_group(rs)
        color = np.zeros((h, w, 3), np.uint8)
        depth = np.zeros((h, w))
        for rank, r in enumerate(rs):
            y0, y1 = r.comm_rows()
            color[y0:y1] = r.read_color()[y0:y1]
            depth[y0:y1] = r.read_depth()[y0:y1]
        with trb.Renderer(port_api) as o:
            mesh, th = o.upload_mesh(m.pos, m.nrm, m.uv, idx), o.upload_texture(tex)
            o.begin_frame(w, h)
            ov = port_api.lookat(cam, [0, 0, 0], [0, 1, 0])
            for mv in (ov, port_api.mat4_mul(ov, shift)):
                o.draw(mesh, mv, pr, kind=trb.SHADER_PHONG, uniforms=_phong(port_api, mv, th), ntris=ntris)
            o.end_frame()
            assert np.array_equal(depth.view(np.uint64), o.read_depth().view(np.uint64))
            assert np.abs(color.astype(int) - o.read_color().astype(int)).max() <= 1
    for r in rs:
        r.close()


@pytest.mark.gpu
def test_draw_shard_shares_partition_the_mesh(cuda_api, monkeypatch):
    """the shares of trb_draw_shard are disjoint and cover the mesh: the ranks' submitted-triangle counts add up, and
    the union of the id planes of the ranks (each rendered alone, no composite) holds every id one context produces"""
    monkeypatch.setenv("TRB_MESH_ORDER_MIN_TRIS", "1")
    m = scenes.icosphere(5)
    w, h = 512, 512
    mv, pr = cuda_api.lookat([0, 0, 2.2], [0, 0, 0], [0, 1, 0]), cuda_api.perspective(60, 1.0, 0.1, 10)
    with trb.Renderer(cuda_api) as r:
        mesh = r.upload_mesh(m.pos, m.nrm, m.uv, m.idx)
        r.begin_frame(w, h)
        r.draw(mesh, mv, pr, ntris=m.ntris)
        whole = r.read_visibility(0).copy()
        whole_z = r.read_depth(0).copy()
        r.end_frame()
        seen = np.full((h, w), 0xFFFFFFFF, np.uint32)
        zmin = np.full((h, w), np.inf)
        total = 0
        for rank in range(3):
            r.begin_frame(w, h)
            r.draw_shard(mesh, mv, pr, rank, 3)
            total += r.stats(0)["triangles_submitted"]
            vis, z = r.read_visibility(0), r.read_depth(0)
            better = (z < zmin) | ((z == zmin) & (vis < seen))
            seen[better], zmin[better] = vis[better], z[better]
            r.end_frame()
        assert total == m.ntris
        assert np.array_equal(seen, whole) and np.array_equal(zmin.view(np.uint64), whole_z.view(np.uint64))


def _ipc_rank(rank, world, out_dir):
    """one process per GPU: blobs exchanged through files, then frames without any host barrier"""
    import time
    sys.path.insert(0, ROOT)
    api = trb.load_cuda()
    m = scenes.icosphere(6)
    w, h = 960, 600
    pr = api.perspective(60, w / h, 0.1, 10)
    idx = np.concatenate([m.idx, m.idx])
    ntris = idx.size // 3
    tex = scenes.texture_diffuse(64, 3)
    with trb.Renderer(api, rank) as r:
        mesh, th = r.upload_mesh(m.pos, m.nrm, m.uv, idx), r.upload_texture(tex)
        r.begin_frame(w, h)

        def all_gather(blob):
            with open(os.path.join(out_dir, "blob%d.tmp" % rank), "wb") as f:
                f.write(blob)
            os.rename(os.path.join(out_dir, "blob%d.tmp" % rank), os.path.join(out_dir, "blob%d" % rank))
            blobs = []
            for k in range(world):
                p = os.path.join(out_dir, "blob%d" % k)
                t0 = time.time()
                while not os.path.exists(p):
                    if time.time() - t0 > 60:
                        raise RuntimeError("rank %d never exported" % k)
                    time.sleep(0.01)
                blobs.append(open(p, "rb").read())
            return blobs

        comm = multigpu.CommComposite(r, all_gather, rank, world)
        comm.open()
        res = []
        for k, cam in enumerate([[0.0, 0.0, 2.2], [0.7, 0.3, 2.0], [-0.5, 0.9, 1.9], [0.1, -0.8, 2.1]]):
            if rank == 1 and k == 2:
                time.sleep(0.3)                   # a straggler: the peers' streams wait on the device, not the hosts
            mv = api.lookat(cam, [0, 0, 0], [0, 1, 0])
            r.begin_frame(w, h)
            if k % 2 == 0:
                first, count = comm.shard(ntris)
                _shard_frame(api, r, mesh, ntris, first, count, mv, pr, th)
            else:                                  # the backend picks the share (trb_draw_shard)
                r.draw_shard(mesh, mv, pr, rank, world, kind=trb.SHADER_PHONG, uniforms=_phong(api, mv, th))
            y0, y1 = comm.run()
            res.append((y0, r.read_depth()[y0:y1].copy(), r.read_color()[y0:y1].copy()))
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array(res, dtype=object), allow_pickle=True)
        r.synchronize()
        # nobody may tear its planes down while a peer still reads them: wait for every rank's result file
        t0 = time.time()
        while not all(os.path.exists(os.path.join(out_dir, "rank%d.npy" % k)) for k in range(world)):
            if time.time() - t0 > 60:
                break
            time.sleep(0.01)
        r.comm_close()


@pytest.mark.gpu
def test_composite_across_processes_without_host_barriers(port_api, tmp_path):
    """trb_comm_export / trb_comm_open / trb_composite with one PROCESS per GPU (CUDA IPC): four frames in a row, one
    rank deliberately late - frame counters in device memory keep the ranks in step.  Needs >= 2 GPUs."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_ipc_rank, args=(world, str(tmp_path)), nprocs=world, join=True)
    m = scenes.icosphere(6)
    w, h = 960, 600
    pr = port_api.perspective(60, w / h, 0.1, 10)
    idx = np.concatenate([m.idx, m.idx])
    tex = scenes.texture_diffuse(64, 3)
    parts = [np.load(tmp_path / ("rank%d.npy" % k), allow_pickle=True) for k in range(world)]
    for k, cam in enumerate([[0.0, 0.0, 2.2], [0.7, 0.3, 2.0], [-0.5, 0.9, 1.9], [0.1, -0.8, 2.1]]):
        zo, co = _unsharded(port_api, m, idx, idx.size // 3, port_api.lookat(cam, [0, 0, 0], [0, 1, 0]), pr, tex, w, h)
        for part in parts:
            y0, z, c = part[k]
            assert np.array_equal(z.view(np.uint64), zo[y0:y0 + z.shape[0]].view(np.uint64)), (k, y0)
            assert np.abs(c.astype(int) - co[y0:y0 + c.shape[0]].astype(int)).max() <= 1

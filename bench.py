#!/usr/bin/env python3
"""bench.py - the rasterization hot path on N B200s (one process per GPU), or the reference's
CPU implementation of the same path (--impl reference).

A "step" is one pass of the hot path over one batch of synthetic input:
  c3 (default)  F frames of the 1920x1080 camera orbit of the three-model scene (config 3,
                SURVEY 8d) per rank per step: vertex -> bin -> raster -> shade for every frame.
                Frames are independent, ranks share nothing: weak scaling, no collective.
  c4            the 20 971 520-triangle sphere at 3840x2160 (config 4), one frame per step.
  c5            tiny-triangle stress, --c5-tris sub-pixel triangles at 8192x8192 (config 5).
  c1            the 800x800 head (config 1), one frame per step (launch bound; reported only).

One JSON line on stdout (rank 0); see DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "triangles/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--frames-per-step", type=int, default=32, help="c3: frames per rank per step")
    ap.add_argument("--c4-level", type=int, default=10)
    ap.add_argument("--c5-tris", type=int, default=100_000_000)
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: worker processes (0 = min(nproc, 8))")
    ap.add_argument("--composite", default="p2p", choices=["p2p", "nccl"], help="c4 on N>1 GPUs: fused NVLink kernel or NCCL all-reduces")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--tga", action="store_true", help="also time trb_encode_tga (device RLE packetiser) on one step's frames")
    ap.add_argument("--no-also", action="store_true", help="c3 only: skip the short config-4 / config-5 runs appended as \"also\"")
    ap.add_argument("--also-steps", type=int, default=5)
    ap.add_argument("--no-exact-shade", action="store_true", help="skip the TRB_SHADE_EXACT=1 (all-fp64 lighting) timing")
    ap.add_argument("--parity-dump", default="", help="reference arm: directory that receives z / bgr of --parity-frames")
    ap.add_argument("--parity-frames", default="", help="reference arm: comma-separated c3 frame indices to dump")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
class Workload:
    """Synthetic inputs of one config plus how to draw one step of it through a Renderer."""

    def __init__(self, args, api):
        from tinyrenderder_b200 import scenes
        self.scenes = scenes
        self.name = args.workload
        self.args = args
        w = args.workload
        if w == "c3":
            self.scene = scenes.orbit_scene()
            self.frames = args.frames_per_step
            self.label = "c3_orbit_1920x1080_%dtri_x%dframes" % (self.scene.ntris, self.frames)
        elif w == "c1":
            self.scene = scenes.head_scene()
            self.frames = 1
            self.label = "c1_head_800x800_%dtri" % self.scene.ntris
        elif w == "c2":
            self.scene = scenes.shadow_scene()
            self.frames = 1
            self.label = "c2_shadow_2048x2048_%dtri_depth_pass+shadow_pass" % self.scene.ntris
        elif w == "c4":
            self.scene = scenes.sphere_scene(args.c4_level)
            self.frames = 1
            self.label = "c4_icosphere_l%d_3840x2160_%dtri" % (args.c4_level, self.scene.ntris)
        else:
            n = args.c5_tris
            _, pos = scenes.triangle_soup(n, 8192, 8192, 0.4, 5, True, want_clip=False)
            mesh = scenes.MeshData(pos, None, None, None, "soup")   # V = 3T, implicit indices
            self.scene = scenes.Scene("c5_soup", 8192, 8192, [scenes.DrawItem(mesh, np.eye(4), 0)], 60, 0.1, 10)
            self.frames = 1
            self.label = "c5_soup_8192x8192_%dtri_r0.4" % n
        sc = self.scene
        self.width, self.height = sc.width, sc.height
        if w == "c5":
            self.perspective = np.eye(4)
        else:
            self.perspective = api.perspective(sc.fov, sc.width / sc.height, sc.znear, sc.zfar)
        self.tris_per_frame = sc.ntris

    def views(self, api, step, rank, world):
        sc = self.scenes
        if self.name == "c3":
            first = ((step * world) + rank) * self.frames
            return sc.orbit_views(api, [(first + j) % 1024 for j in range(self.frames)])
        if self.name in ("c1", "c2"):
            return sc.head_view(api)[None]
        if self.name == "c4":
            return sc.sphere_view(api)[None]
        return np.eye(4)[None]


    def render(self, up, views):
        """one step through an UploadedScene: the frame loop of main.cpp; config 2 adds the light's depth pass"""
        if self.name == "c2":
            self.scenes.render_shadowed(up, views[0], self.perspective)
        else:
            up.render(views, self.perspective)


def algorithmic_bytes(V, T, P, R, T_vis, C, frames=1):
    """SURVEY 8(d) per-frame formula, split by kernel (DESIGN.md 'Algorithmic bytes')."""
    b = {
        "k_vertex_mesh": V * (12 + 32),
        "k_setup_count": T * (12 + 96),
        "k_fill": 4 * R,
        "k_raster": R * (4 + 96) + 12 * P,
        "k_shade": 12 * P + 96 * T_vis + 9 * C + 3 * P,
        "k_clear": 8 * P,
    }
    return {k: v * frames for k, v in b.items()}


# kernels that are flavours of one pass share that pass' algorithmic bytes
PASS_OF = {"k_raster_warp": "k_raster", "k_shade_dense": "k_shade"}
SYNC_KERNELS = ("k_comm_wait", "k_comm_publish")     # waiting for peers is not work: never the "dominant kernel"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.first = 0
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def mark(self):
        """samples before this point were taken before the timed region"""
        self.first = len(self.samples)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples[max(0, self.first - 1):]:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU rasterizer on the host cores
# ---------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init(libpath, workload_args):
    import tinyrenderder_b200 as trb
    api = trb.Api(libpath, "orc")
    args = argparse.Namespace(**workload_args)
    wl = Workload(args, api)
    r = trb.Renderer(api)
    up = wl.scenes.UploadedScene(r, wl.scene)
    _W.update(api=api, wl=wl, r=r, up=up)


def _ref_worker_frame(job):
    step, slot, world, tri_first, tri_count = job
    wl, api, up, r = _W["wl"], _W["api"], _W["up"], _W["r"]
    if wl.name == "c3":
        views = wl.scenes.orbit_views(api, [(step * world + slot) % 1024])
    else:
        views = wl.views(api, step, 0, 1)
    t0 = time.perf_counter()
    if tri_count is None:
        wl.render(up, views)
        ntri = wl.tris_per_frame
    else:  # bounded sample of a huge single frame: a contiguous triangle range (frame clear not timed)
        it = wl.scene.items[0]
        r.begin_frame(wl.width, wl.height)
        mv = api.mat4_mul(views[0], it.model_matrix)
        t0 = time.perf_counter()
        r.draw(up.mesh_h[id(it.mesh)], mv, wl.perspective, kind=it.kind, first_tri=tri_first, ntris=tri_count)
        r.end_frame()
        ntri = tri_count
    dt = time.perf_counter() - t0
    return ntri, r.stats()["fragments_covered"] if api.backend_name() == "oracle-port" else 0, dt


def _ref_worker_dump(job):
    """parity material for the GPU arm: the reference's z-buffer and framebuffer of one orbit frame"""
    k, out_dir = job
    wl, api, up, r = _W["wl"], _W["api"], _W["up"], _W["r"]
    wl.render(up, wl.scenes.orbit_views(api, [k]) if wl.name == "c3" else wl.views(api, 0, 0, 1))
    np.save(os.path.join(out_dir, "z_%d.npy" % k), r.read_depth(0))
    np.save(os.path.join(out_dir, "bgr_%d.npy" % k), r.read_color(0))
    return k


def oracle_library():
    ref = os.path.join(ROOT, "oracle", "_ref", "libtrb_ref.so")
    if os.path.exists(ref):
        return ref, "reference"
    return os.path.join(ROOT, "oracle", "libtrb_port.so"), "port"


def run_reference(args, steps, warmup):
    """The reference's CPU path (oracle/_ref = its own our_gl.cpp; else the port) on the host cores.
    It keeps its state in unsynchronised globals (our_gl.cpp:12-22), so parallelism is one process
    per frame (c3) / per triangle range (c4, c5), as SURVEY 8(d) prescribes."""
    import multiprocessing as mp
    lib, kind = oracle_library()
    procs = args.ref_procs or min(os.cpu_count() or 1, 8)
    wargs = dict(vars(args))
    if args.workload == "c5":
        wargs["c5_tris"] = min(args.c5_tris, 4_000_000 * procs)  # bounded sample: a prefix of the soup
    ctx = mp.get_context("fork")
    pool = ctx.Pool(procs, initializer=_ref_worker_init, initargs=(lib, wargs))
    try:
        tris_per_frame = None
        times = []
        tris_total = 0
        sample = ""
        for s in range(warmup + steps):
            if args.workload in ("c1", "c2", "c3"):
                jobs = [(s, p, procs, 0, None) for p in range(procs)]
                sample = "%d frame(s) per step, one per process" % procs
            else:
                # one frame is tens of seconds on a CPU: each process rasterises a 1/64 range
                import tinyrenderder_b200  # noqa: F401
                total = wargs["c5_tris"] if args.workload == "c5" else 20 * 4 ** args.c4_level
                chunk = max(1, total // 64 // procs) if args.workload == "c4" else total // procs
                jobs = [(s, p, procs, ((s * procs + p) * chunk) % max(1, total - chunk), chunk) for p in range(procs)]
                sample = "%d triangle ranges of %d per step (1/%d of a frame each)" % (procs, chunk, max(1, total // chunk))
            res = pool.map(_ref_worker_frame, jobs)
            dt = max(x[2] for x in res)  # the processes run side by side: the step ends with the slowest
            if s >= warmup:
                times.append(dt)
                tris_total += sum(r[0] for r in res)
            tris_per_frame = res[0][0]
        if args.parity_dump and args.parity_frames and args.workload in ("c1", "c2", "c3"):
            os.makedirs(args.parity_dump, exist_ok=True)
            pool.map(_ref_worker_dump, [(int(k), args.parity_dump) for k in args.parity_frames.split(",")])
    finally:
        pool.close()
        pool.join()
    total_t = sum(times)
    value = tris_total / total_t
    return {"value": value, "kind": kind, "cores": procs, "sample": sample, "ms_per_step": 1e3 * total_t / max(1, len(times)),
            "tris_per_step": tris_total / max(1, len(times)), "tris_per_unit": tris_per_frame}


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPU cores next to its GPU (NVML's ideal affinity) before any pinned host
    buffer is allocated: with 8 ranks the read-backs otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


# ---------------------------------------------------------------------------------------------
# parity of what was just timed
# ---------------------------------------------------------------------------------------------
GOLDEN_FULLSIZE = os.path.join(ROOT, "tests", "golden", "golden_fullsize.json")


def _sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parity_check(wl, r, frames, env, sharded_rows=None):
    """Compare the frames the timed region left in the context with the reference's own output: SHA-256 digests
    of the z-buffers (tests/golden/golden_fullsize.json, made by oracle/_ref = the reference's rasterize() compiled
    from /root/reference; every one of the 1024 orbit frames, the config-4 sphere, the config-5 soup).  The colour
    of the lit config-3 frames is compared later with the frames the cpu_baseline run renders (returns them)."""
    torch, dist, rank, world = env["torch"], env["dist"], env["rank"], env["world"]
    try:
        gold = json.load(open(GOLDEN_FULLSIZE))
    except Exception as e:
        return {"frames": 0, "depth": "unchecked: %r" % (e,)}, {}
    out = {"source": "tests/golden/golden_fullsize.json (digests of the reference's own rasterize() output)"}
    kept = {}
    if wl.name == "c3":
        g = gold.get("c3_orbit", {})
        if g.get("workload") != wl.label.split("_x")[0]:
            return {"frames": 0, "depth": "unchecked: no golden for %s" % wl.label}, {}
        ok = 0
        for v, k in enumerate(frames):
            ok += 1 if _sha(r.read_depth(v)) == g["z_sha256"][k] else 0
        for v in sorted({0, len(frames) - 1}):
            kept[frames[v]] = r.read_color(v).copy()
        t = torch.tensor([ok, len(frames)], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        ok, n = int(t[0].item()), int(t[1].item())
        out.update(frames=n, depth="bit-exact" if ok == n else "MISMATCH in %d of %d frames" % (n - ok, n),
                   frame_indices_rank0=[frames[0], frames[-1]])
        return out, kept
    key = {"c4": "c4_sphere", "c5": "c5_soup"}.get(wl.name)
    g = gold.get(key or "", {})
    if not g or g.get("workload") != wl.label:
        return {"frames": 0, "depth": "unchecked: no golden for %s" % wl.label}, {}
    z, c = r.read_depth(0), r.read_color(0)
    if sharded_rows is not None:    # sort-last composite: every rank owns rows [y0, y1) of the picture
        y0, y1 = sharded_rows
        parts = [None] * world
        dist.all_gather_object(parts, (y0, z[y0:y1].copy(), c[y0:y1].copy()))
        for py0, pz, pc in parts:
            z[py0:py0 + pz.shape[0]] = pz
            c[py0:py0 + pc.shape[0]] = pc
    zs, cs = _sha(z) == g["z_sha256"], _sha(c) == g["bgr_sha256"]
    out.update(frames=1, depth="bit-exact" if zs else "MISMATCH", colour="bit-exact (flat shader)" if cs else "MISMATCH",
               pixels_shaded=int(np.isfinite(z).sum()))
    if sharded_rows is not None:
        out["note"] = "composited picture of all ranks compared with the UNSHARDED reference render"
    return out, {}


# ---------------------------------------------------------------------------------------------
# one workload: diagnostics, warm-up, timed region, parity, roofline (and for the primary one e2e + cpu baseline)
# ---------------------------------------------------------------------------------------------
def measure(args, env, workload, steps, warmup, primary, composite="p2p"):
    import tinyrenderder_b200 as trb
    from tinyrenderder_b200 import multigpu
    torch, dist, rank, world, local_rank, api = (env[k] for k in ("torch", "dist", "rank", "world", "local_rank", "api"))
    wargs = argparse.Namespace(**dict(vars(args), workload=workload))
    r = trb.Renderer(api, local_rank)
    wl = Workload(wargs, api)
    up = wl.scenes.UploadedScene(r, wl.scene)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sharded_c4 = wl.name == "c4" and world > 1
    def gather_objects(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # composite group of the C ABI: IPC blobs exchanged once, then no host barrier per frame (trb_composite)
    p2p = multigpu.CommComposite(r, gather_objects, rank, world) if sharded_c4 and composite == "p2p" else None
    rows = multigpu.row_shard(wl.height, rank, world) if sharded_c4 else None

    def frames_of(s):
        if wl.name != "c3":
            return [0]
        first = ((s * world) + rank) * wl.frames
        return [(first + j) % 1024 for j in range(wl.frames)]

    def step(s):
        if not sharded_c4:
            wl.render(up, wl.views(api, s, rank, world))
            return
        # config 4 on N GPUs: every rank draws its share of the mesh (trb_draw_shard: blocks of the mesh's processing
        # order dealt out round robin, ids global), sort-last composite, shade own rows
        it = wl.scene.items[0]
        r.begin_frame(wl.width, wl.height)
        if p2p is not None and not p2p.opened:
            p2p.open()                    # once: the planes of this frame size are exported to the peers
        mv = api.mat4_mul(wl.views(api, s, rank, world)[0], it.model_matrix)
        r.draw_shard(up.mesh_h[id(it.mesh)], mv, wl.perspective, rank, world, kind=it.kind)
        if composite == "p2p":
            p2p.run()                     # fused NVLink composite + shade of the owned rows, stream-ordered against the peers
        else:
            multigpu.composite(r, lambda t: dist.all_reduce(t, op=dist.ReduceOp.MIN))
            r.set_shade_rows(*rows)
            r.end_frame()

    # ---- diagnostics pass (untimed): counters for the algorithmic-bytes formula ---------------------
    step(0)
    nviews = wl.frames
    R = sum(r.stats(v)["tile_entries"] for v in range(nviews))
    C = sum(r.stats(v)["pixels_shaded"] for v in range(nviews))
    frag = sum(r.stats(v)["fragments_covered"] for v in range(nviews))
    # distinct winning triangles: re-draw view 0 without the flush and look at the id plane
    views0 = wl.views(api, 0, rank, world)[:1]
    r.begin_frame(wl.width, wl.height)
    for it in wl.scene.items:
        mv = api.mat4_mul(views0[0], it.model_matrix)
        r.draw(up.mesh_h[id(it.mesh)], mv, wl.perspective, kind=0, ntris=it.mesh.ntris)
    vis = r.read_visibility(0)
    tvis = int(np.unique(vis[(vis != 0xFFFFFFFF) & (vis != 0)]).size) * nviews
    del vis
    r.end_frame()
    V = sum(it.mesh.nverts for it in wl.scene.items)
    T = wl.tris_per_frame
    P = wl.width * wl.height
    balg = algorithmic_bytes(V * nviews, T * nviews, P * nviews, R, tvis, C)
    if sharded_c4:
        # per RANK: its triangle range through set-up / bins / raster over the whole picture (R is this rank's own
        # counter), the unsharded vertex stage, and the shading of the rows it owns
        t_rank = multigpu.triangle_shard(T, rank, world)[1]
        balg = algorithmic_bytes(V, t_rank, P, R, tvis // world, C // world)
        balg["k_shade"] = (12 * P + 96 * tvis + 9 * C + 3 * P) // world
        if composite == "p2p":
            balg["k_composite_shade_p2p"] = 12 * P + balg.pop("k_shade")  # every rank reads world x 12 B for P / world pixels

    # ---- warm-up + timed region -----------------------------------------------------------------------
    for s in range(warmup):
        step(s)
    # the same K steps once without the per-kernel events (reported as ms_per_step_unprofiled: what a
    # caller sees), then the timed region proper with every launch bracketed by CUDA events
    barrier()
    r.timer_start()
    for s in range(steps):
        step(warmup + s)
    ms_unprofiled = r.timer_stop_ms()
    barrier()
    r.profile_enable(True)
    r.profile_read(reset=True)
    launches0 = r.launch_count()
    clocks = ClockSampler(local_rank) if primary else None
    if clocks:
        clocks.start()
        time.sleep(0.2)
    barrier()
    if clocks:
        clocks.mark()
    r.timer_start()
    t_wall0 = time.perf_counter()
    for s in range(steps):
        step(warmup + s)
    ms = r.timer_stop_ms()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if clocks else None
    prof = r.profile_read(reset=True)
    r.profile_enable(False)
    launches = r.launch_count() - launches0
    t = torch.tensor([ms, ms_unprofiled], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, ms_unprofiled = float(t[0].item()), float(t[1].item())

    tris_step_all = T * nviews * (1 if sharded_c4 else world)
    value = tris_step_all * steps / (ms_max * 1e-3)

    # ---- parity of the frames the timed region just produced (still resident in the context) -------------
    last_frames = frames_of(warmup + steps - 1)
    parity, kept_colour = parity_check(wl, r, last_frames, env, rows)

    # ---- all-fp64 lighting (TRB_SHADE_EXACT=1) timed beside the default fp32 lighting --------------------
    exact = None
    if primary and not args.no_exact_shade and not sharded_c4 and wl.name in ("c1", "c2", "c3"):
        os.environ["TRB_SHADE_EXACT"] = "1"
        try:
            with trb.Renderer(api, local_rank) as r2:
                up2 = wl.scenes.UploadedScene(r2, wl.scene)
                for s in range(2):
                    wl.render(up2, wl.views(api, s, rank, world))
                torch.cuda.synchronize()
                r2.timer_start()
                k2 = max(1, min(steps, 5))
                for s in range(k2):
                    wl.render(up2, wl.views(api, warmup + s, rank, world))
                exact = {"ms_per_step": r2.timer_stop_ms() / k2, "steps": k2,
                         "note": "same steps with PhongShader / EyeShader evaluated in fp64 in the reference's operation order"}
        finally:
            del os.environ["TRB_SHADE_EXACT"]

    tga = None
    if primary and args.tga and not sharded_c4:
        step(0)
        r.profile_enable(True)
        r.profile_read(reset=True)
        t0 = time.perf_counter()
        files = r.encode_tga(0)                                        # framebuffer.tga of every frame, main.cpp:743
        dt = time.perf_counter() - t0
        kt = r.profile_read(reset=True)
        r.profile_enable(False)
        tga = {"frames": len(files), "bytes": sum(len(f) for f in files), "raw_bytes": nviews * P * 3,
               "wall_ms": 1e3 * dt, "device_ms": sum(m for k, (n, m) in kt.items() if k.startswith("k_rle")),
               "note": "device-side packetiser of tgaimage.cpp:193-242 + D2H of the packets, one blocking call"}

    # ---- end-to-end: host buffers in, host buffers out, every step --------------------------------------
    e2e = None
    if primary and not args.no_e2e and not sharded_c4:
        e2e = measure_e2e(args, env, wl, r, steps, tris_step_all, nviews, P, barrier)

    if world > 1:
        barrier()
    r.close()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kern = {k: {"launches": int(n), "ms": float(m)} for k, (n, m) in prof.items()}
    total_k_ms = sum(v["ms"] for v in kern.values()) or 1.0
    work = {k: v for k, v in kern.items() if k not in SYNC_KERNELS}
    dom = max(work, key=lambda k: work[k]["ms"]) if work else None
    roof = None
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tr_path):
        try:
            traffic = json.load(open(tr_path)).get(wl.name, {}).get(dom)
        except Exception:
            traffic = None
    if dom:
        per_launch_ms = kern[dom]["ms"] / kern[dom]["launches"]
        # launches of the dominant kernel per step (one per draw call) share the step's algorithmic bytes
        bytes_per_launch = balg.get(PASS_OF.get(dom, dom), 0) * steps / kern[dom]["launches"]
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/ncu_traffic.json: dram bytes of one separate `ncu --set full` capture of this "
                                  "kernel, scaled to this launch size (not measured in this run)" if traffic else None,
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms,
                "share_of_kernel_time": kern[dom]["ms"] / total_k_ms}
    step_bytes = sum(balg.values())
    step_gbs = step_bytes * steps / (ms_max * 1e-3) / 1e9

    cpu = None
    if primary and world == 1 and not args.no_cpu_baseline:
        import shutil
        import tempfile
        dump = tempfile.mkdtemp(prefix="trb_parity_")
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
                   "--steps", "2", "--warmup", "0", "--c4-level", str(args.c4_level), "--c5-tris", str(args.c5_tris)]
            if kept_colour:
                cmd += ["--parity-dump", dump, "--parity-frames", ",".join(str(k) for k in sorted(kept_colour))]
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            ref_line = json.loads(out.stdout.strip().splitlines()[-1])
            cpu = ref_line["cpu_baseline"]
            # colour (and once more depth) of the frames both arms rendered: GPU frames kept from the timed step
            fr, zok = [], True
            for k, bgr in sorted(kept_colour.items()):
                want = np.load(os.path.join(dump, "bgr_%d.npy" % k))
                d = np.abs(bgr.astype(np.int32) - want.astype(np.int32)).max(axis=-1)
                fr.append(float((d <= 1).mean()))
                zok &= _sha(np.load(os.path.join(dump, "z_%d.npy" % k))) == json.load(open(GOLDEN_FULLSIZE))["c3_orbit"]["z_sha256"][k]
            if fr:
                parity["colour_within_1lsb"] = min(fr)
                parity["colour_frames"] = sorted(kept_colour)
                parity["colour_vs"] = "the same frames rendered by the cpu_baseline run (%s); its z-buffers %s the committed digests" % (
                    cpu.get("kind"), "match" if zok else "DO NOT match")
        except Exception as e:  # the baseline is a reported figure; never let it sink the GPU line
            cpu = {"value": None, "unit": "triangles/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        finally:
            shutil.rmtree(dump, ignore_errors=True)
    if wl.name == "c3":
        parity.setdefault("colour_within_1lsb", None)

    line = {
        "metric": METRIC, "value": value, "unit": "triangles/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True,
        "scaling": "strong" if sharded_c4 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "precision": "coverage, depth, barycentrics, texture coordinates and texel choice in f64 (bit-exact); "
                     "lighting of the lit shaders in f32 (within 1 LSB, asserted by tests and parity_check)",
        "config": {"workload": wl.label, "frames_per_step_per_gpu": nviews, "width": wl.width, "height": wl.height,
                   "triangles_per_frame": T, "l2": "inputs larger than L2 (depth+id+colour planes of one step = %d MB)"
                   % (nviews * P * 15 // 2 ** 20), "parallelism": ("triangle ranges + %s sort-last composite" % ("fused NVLink P2P (trb_composite: device-side frame counters, no host barrier)" if composite == "p2p" else "NCCL all-reduces")
                                   if sharded_c4 else
                                   "frames sharded, no collective") if world > 1 else "1 GPU"},
        "fragments_per_s": frag * (1 if sharded_c4 else world) * steps / (ms_max * 1e-3),
        "pixels_shaded_per_s": C * (1 if sharded_c4 else world) * steps / (ms_max * 1e-3),
        "frame_ms": ms_max / steps / nviews,
        "frames_per_s": nviews * (1 if sharded_c4 else world) * steps / (ms_max * 1e-3),
        "parity_check": parity,
        "roofline": roof,
        "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "achieved_gbs": step_gbs, "frac": step_gbs / peak,
                          "counters": {"V": V * nviews, "T": T * nviews, "P": P * nviews, "R": R, "T_vis": tvis, "C": C}},
        "kernels": kern,
        "gpu_launches": launches,
        "ms_per_step_unprofiled": ms_unprofiled / steps,
    }
    if primary:
        line.update({"cpu_baseline": cpu, "e2e": e2e, "clocks": clk, "wall_ms_per_step": 1e3 * t_wall / steps,
                     "shade_exact_f64": exact, "tga_encode": tga, "cpu_affinity_cores": env["numa"]})
    return line


def measure_e2e(args, env, wl, r, steps, tris_step_all, nviews, P, barrier):
    """The same metric through the C ABI with HOST buffers: every step uploads its inputs from pinned host memory
    and reads its frames back into pinned host memory (copies inside the timed region)."""
    torch, dist, rank, world, api = (env[k] for k in ("torch", "dist", "rank", "world", "api"))
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    # three sets of pinned host buffers (the library keeps up to three read-backs in flight): the read-back of step s
    # overlaps the rendering of steps s+1 and s+2
    NSETS = 3
    color_host = [[pin((wl.height, wl.width, 3), torch.uint8) for _ in range(nviews)] for _ in range(NSETS)]
    depth_host = [[pin((wl.height, wl.width), torch.float64) for _ in range(nviews)] for _ in range(NSETS)]
    e_steps = max(1, min(steps, 1 if wl.name == "c5" else (3 if wl.name == "c4" else steps)))
    host_ms = {"upload": 0.0, "render": 0.0, "readback": 0.0, "free": 0.0}

    def pinned(a):
        if a is None:
            return None
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    pinned_tex = {}
    # the step's inputs live in pinned host memory (the contract's H2D source): meshes and textures
    for it in wl.scene.items:
        m = it.mesh
        if not getattr(m, "_pinned", False):
            m.pos, m.nrm, m.uv, m.idx = pinned(m.pos), pinned(m.nrm), pinned(m.uv), pinned(m.idx)
            m._pinned = True
        for k in list(it.textures):
            t = it.textures[k]
            if id(t) not in pinned_tex:              # maps shared between models stay shared
                p = pinned(t)
                pinned_tex[id(t)] = p
                pinned_tex[id(p)] = p
            it.textures[k] = pinned_tex[id(t)]
    up_resident = wl.scenes.UploadedScene(r, wl.scene)
    per_step_uniform_bytes = nviews * sum(32 * 8 + (104 if it.kind in (1, 2) else 0) for it in wl.scene.items)

    def e2e_step(s, with_depth, resident=False):
        t = [time.perf_counter()]
        if resident:   # meshes and textures uploaded once (north_star); per step only matrices and uniforms go up
            wl.render(up_resident, wl.views(api, s, rank, world))
            r.readback_async(color_host[s % NSETS], depth_host[s % NSETS] if with_depth else None)
            return per_step_uniform_bytes
        up2 = wl.scenes.UploadedScene(r, wl.scene)                 # H2D: meshes + textures
        t.append(time.perf_counter())
        wl.render(up2, wl.views(api, s, rank, world))              # H2D: matrices, uniforms
        t.append(time.perf_counter())
        # D2H: the BGR framebuffer of every frame (what the reference writes out, main.cpp:743);
        # with_depth also brings back the f64 z-buffer the reference keeps in a host global
        r.readback_async(color_host[s % NSETS], depth_host[s % NSETS] if with_depth else None)
        t.append(time.perf_counter())
        up2.free()
        t.append(time.perf_counter())
        for k, a, b in zip(("upload", "render", "readback", "free"), t[:-1], t[1:]):
            host_ms[k] += 1e3 * (b - a)
        return up2.h2d_bytes

    def e2e_run(with_depth, resident=False):
        h2d = e2e_step(0, with_depth, resident)
        e2e_step(1, with_depth, resident)                          # the block cache settles after two or three frames
        e2e_step(2, with_depth, resident)
        for k in host_ms:
            host_ms[k] = 0.0
        r.readback_wait()
        barrier()
        t0 = time.perf_counter()
        for s in range(e_steps):
            e2e_step(3 + s, with_depth, resident)
        t_host = time.perf_counter() - t0                          # host time to enqueue the steps (nothing waited for)
        r.readback_wait()                                          # every host buffer is complete here
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"value": tris_step_all * e_steps / float(t.item()), "unit": "triangles/s",
                "h2d_bytes_per_step": int(h2d + (0 if resident else per_step_uniform_bytes)),
                "d2h_bytes_per_step": int(nviews * P * (3 + (8 if with_depth else 0))),
                "steps": e_steps, "ms_per_step": 1e3 * float(t.item()) / e_steps,
                "host_enqueue_ms_per_step": 1e3 * t_host / e_steps,
                "host_ms_per_step": {k: v / e_steps for k, v in host_ms.items()}}

    # ---- the same with the frames coming home as TGA FILES (what main.cpp:743 produces): RLE packets built on the
    #      device behind the frame (trb_encode_tga_async), only the packets cross PCIe
    cap = wl.width * wl.height * 3 + wl.width * wl.height // 2 + 64
    tga_host = [[pin((cap,), torch.uint8) for _ in range(nviews)] for _ in range(2)]
    tga_sizes = [np.zeros(nviews, dtype=np.uint64) for _ in range(2)]

    def tga_step(s):
        up2 = wl.scenes.UploadedScene(r, wl.scene)
        wl.render(up2, wl.views(api, s, rank, world))
        r.encode_tga_async(tga_host[s & 1], tga_sizes[s & 1], 0)
        up2.free()
        return up2.h2d_bytes

    def tga_run():
        for s in range(3):
            h2d = tga_step(s)
        r.readback_wait()
        barrier()
        t0 = time.perf_counter()
        sent = 0
        for s in range(e_steps):
            tga_step(3 + s)
        r.readback_wait()
        barrier()
        dt = time.perf_counter() - t0
        sent = int(tga_sizes[0].sum() + tga_sizes[1].sum()) // 2       # the last two steps' files
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"value": tris_step_all * e_steps / float(t.item()), "unit": "triangles/s",
                "h2d_bytes_per_step": int(h2d + per_step_uniform_bytes), "d2h_bytes_per_step": sent,
                "raw_bytes_per_step": int(nviews * P * 3), "steps": e_steps, "ms_per_step": 1e3 * float(t.item()) / e_steps,
                "note": "frames returned as complete RLE TGA files (byte-identical to TGAImage::write_tga_file), packetised on "
                        "the device; on a host whose ingest saturates (many GPUs) fewer bytes cross, on one GPU the encode "
                        "costs more SM time than the raw copy saves"}

    e2e = e2e_run(False)
    e2e["note"] = ("per step: upload meshes+textures, render, read back the BGR framebuffer of every frame into pinned "
                   "host memory; the z-buffer stays in HBM for the device-side post passes")
    if wl.name in ("c1", "c2", "c3"):
        e2e["tga_files"] = tga_run()
    e2e["with_depth_readback"] = e2e_run(True)   # same, plus the f64 z-buffer of every frame (PCIe bound)
    # the deployment north_star describes: meshes / textures go to HBM once, a step uploads matrices and uniforms only
    e2e["scene_resident"] = e2e_run(False, resident=True)
    e2e["scene_resident"].pop("host_ms_per_step", None)
    e2e["per_gpu_value"] = e2e["value"] / world
    return e2e


# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        import __graft_entry__ as g
        g.build_scenegen()
        steps, warmup = args.steps, args.warmup
        res = run_reference(args, steps, warmup)
        import tinyrenderder_b200 as trb
        wl_label = Workload(args, trb.Api(oracle_library()[0], "orc")).label if args.workload != "c5" else "c5_soup"
        line = {
            "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "triangles/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_label, "cpu_threads": res["cores"], "sample": res["sample"]},
            "cpu_baseline": {"value": res["value"], "unit": "triangles/s", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import tinyrenderder_b200 as trb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = {"torch": torch, "dist": dist, "rank": rank, "world": world, "local_rank": local_rank, "api": trb.load_cuda(),
           "numa": numa}

    line = measure(args, env, args.workload, args.steps, args.warmup, True, args.composite)
    # ---- configs 4 and 5 ride along on the default (config 3) run: short runs, each with its own roofline and
    #      parity_check, so that the driver's record covers them; on N > 1 GPUs config 4 is the sharded one
    if args.workload == "c3" and not args.no_also:
        also = {}
        k, w = max(1, args.also_steps), 2
        try:
            if world > 1:
                also["c4_p2p"] = measure(args, env, "c4", k, w, False, "p2p")
                also["c4_nccl"] = measure(args, env, "c4", k, w, False, "nccl")
            else:
                also["c4"] = measure(args, env, "c4", k, w, False)
                also["c5"] = measure(args, env, "c5", k, w, False)
        except Exception as e:   # the secondary runs must never sink the headline line
            also["error"] = repr(e)
        if line is not None:
            line["also"] = also
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

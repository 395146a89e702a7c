#!/usr/bin/env python3
"""How fast can this host take finished frames from N GPUs at once?  

One process per GPU (torchrun), each copying device buffers of a step's frames (32 x 1920x1080x3 B = 199 MB) into page-locked
host memory, all ranks at the same time.  Variants: one 199 MB copy vs 32 copies of one frame, default vs write-combined
pinned memory, rank bound to the CPU cores next to its GPU or not.  Prints one JSON line (rank 0): aggregate GB/s per variant.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/experiments/d2h_bandwidth.py
"""
import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FRAME = 1920 * 1080 * 3
FRAMES = 32
REPS = 20


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cudart = ctypes.CDLL("libcudart.so.12")
    results = {}
    src = torch.randint(0, 255, (FRAMES * FRAME,), dtype=torch.uint8, device="cuda")

    def host_alloc(flags):
        p = ctypes.c_void_p()
        rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(FRAMES * FRAME), ctypes.c_uint(flags))
        assert rc == 0, rc
        return p

    def run(name, dst_ptr, pieces):
        stream = torch.cuda.Stream()
        n = FRAMES * FRAME // pieces
        def once():
            for k in range(pieces):
                rc = cudart.cudaMemcpyAsync(ctypes.c_void_p(dst_ptr.value + k * n), ctypes.c_void_p(src.data_ptr() + k * n),
                                            ctypes.c_size_t(n), ctypes.c_int(2), ctypes.c_void_p(stream.cuda_stream))
                assert rc == 0, rc
        once()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(REPS):
            once()
        stream.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[name] = {"aggregate_GBps": world * REPS * FRAMES * FRAME / float(t.item()) / 1e9,
                         "per_gpu_GBps": REPS * FRAMES * FRAME / float(t.item()) / 1e9}

    for bound in (False, True):
        if bound:
            import bench
            results["cores_bound_per_rank"] = bench.bind_to_gpu_numa_node(local)
        tag = "numa_bound" if bound else "unbound"
        d = host_alloc(0)
        run("default_32copies_" + tag, d, FRAMES)
        run("default_1copy_" + tag, d, 1)
        cudart.cudaFreeHost(d)
        wc = host_alloc(4)                                   # cudaHostAllocWriteCombined
        run("writecombined_1copy_" + tag, wc, 1)
        cudart.cudaFreeHost(wc)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_step_per_gpu": FRAMES * FRAME, "reps": REPS, "host_cpus": os.cpu_count(),
                          "variants": results}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

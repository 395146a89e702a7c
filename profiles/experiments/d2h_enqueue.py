#!/usr/bin/env python3
"""Does queueing D2H copies behind an event that has not fired yet block the host?  (It decides how
trb_readback_async hands frames to the copy engine.)"""
import time
import torch

dev = torch.device("cuda:0")
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
src = torch.empty(32, 1080, 1920, 3, dtype=torch.uint8, device=dev)


def busy(ms_target=12):
    with torch.cuda.stream(sA):
        for _ in range(int(ms_target / 0.75)):
            torch.matmul(a, a)


for label, ncopies in (("32 copies of 6.2 MB", 32), ("1 copy of 199 MB", 1), ("256 copies of 0.78 MB", 256)):
    if ncopies == 1:
        dst = [torch.empty(32, 1080, 1920, 3, dtype=torch.uint8).pin_memory()]
        srcs = [src]
    else:
        per = 32 * 1080 * 1920 * 3 // ncopies
        flat = src.view(-1)
        dst = [torch.empty(per, dtype=torch.uint8).pin_memory() for _ in range(ncopies)]
        srcs = [flat[i * per:(i + 1) * per] for i in range(ncopies)]
    for rep in range(3):
        torch.cuda.synchronize()
        busy()
        ev = torch.cuda.Event()
        ev.record(sA)
        t0 = time.perf_counter()
        with torch.cuda.stream(sB):
            sB.wait_event(ev)
            for d, s in zip(dst, srcs):
                d.copy_(s, non_blocking=True)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("%-24s enqueue %.2f ms, until done %.2f ms" % (label, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))

#!/usr/bin/env python
"""Micro-probe: does an otherwise idle host slow the per-step device chain (H2D of the observation slab, a small kernel
chain, D2H of the actions)?  Runs the chain spaced by short host sleeps (as in a rollout) with 0 / 2 / 6 / 14 busy helper
processes and prints per-piece CUDA-event times."""
import multiprocessing as mp
import os
import sys
import time

import torch


def spinner(stop):
    x = 0
    while not stop.value:
        for _ in range(10000):
            x += 1


def measure(tag, slab, dev, small_dev, small_pin, reps=300, gap=0.0005):
    s = torch.cuda.current_stream()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
    a = torch.randn(32, 3136, device="cuda")
    w = torch.randn(3136, 256, device="cuda")
    host_wait = []
    for i in range(reps):
        t_end = time.perf_counter() + gap
        while time.perf_counter() < t_end:
            pass
        ev[i][0].record(s)
        dev.copy_(slab, non_blocking=True)
        ev[i][1].record(s)
        for _ in range(6):
            b = a @ w
        ev[i][2].record(s)
        small_pin.copy_(small_dev, non_blocking=True)
        ev[i][3].record(s)
        t0 = time.perf_counter()
        while not ev[i][3].query():
            pass
        host_wait.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    h2d = sorted(e[0].elapsed_time(e[1]) for e in ev)
    ker = sorted(e[1].elapsed_time(e[2]) for e in ev)
    d2h = sorted(e[2].elapsed_time(e[3]) for e in ev)
    hw = sorted(host_wait)
    m = reps // 2
    print("[pcie] %-22s H2D %.0f us (p90 %.0f) = %.1f GB/s | 6 small GEMMs %.0f us | D2H %.0f us | host sees done after %.0f us (p90 %.0f)"
          % (tag, h2d[m] * 1e3, h2d[int(reps * .9)] * 1e3, slab.numel() * 4 / (h2d[m] * 1e-3) / 1e9, ker[m] * 1e3, d2h[m] * 1e3,
             hw[m] * 1e6, hw[int(reps * .9)] * 1e6), flush=True)


def main():
    torch.cuda.init()
    slab = torch.zeros((32, 3, 84, 84), dtype=torch.float32).share_memory_()
    rc = torch.cuda.cudart().cudaHostRegister(slab.data_ptr(), slab.numel() * 4, 0)
    pinned = torch.zeros((32, 3, 84, 84), dtype=torch.float32).pin_memory()
    dev = torch.zeros((32, 3, 84, 84), device="cuda")
    small_dev = torch.zeros(32, dtype=torch.long, device="cuda")
    small_pin = torch.zeros(32, dtype=torch.long).pin_memory()
    print("[pcie] cudaHostRegister rc", rc, "cpus", os.cpu_count(), flush=True)
    measure("idle host, slab", slab, dev, small_dev, small_pin)
    measure("idle host, cudaMallocHost", pinned, dev, small_dev, small_pin)
    measure("idle host, no gaps", slab, dev, small_dev, small_pin, gap=0.0)
    ctx = mp.get_context("fork")
    for n in (2, 6, 14):
        stop = ctx.Value("i", 0)
        procs = [ctx.Process(target=spinner, args=(stop,), daemon=True) for _ in range(n)]
        for p in procs:
            p.start()
        time.sleep(0.3)
        measure("%d busy helpers" % n, slab, dev, small_dev, small_pin)
        stop.value = 1
        for p in procs:
            p.join(2)


if __name__ == "__main__":
    main()

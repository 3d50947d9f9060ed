#!/usr/bin/env python
"""Micro-benchmark of the library GEMMs on the training shapes (CUDA events; run with TRXL_TCGEN05=0/1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "episodic-transformer-memory-ppo_b200"))
import torch  # noqa: E402
import trxl_native as native  # noqa: E402


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    dev = "cuda:0"
    mode = "tcgen05" if os.environ.get("TRXL_TCGEN05", "0") == "1" else "simt"
    for m, n, k in ((2048, 256, 256), (2048, 384, 256), (2048, 256, 3136), (16384, 384, 384), (16384, 512, 512)):
        x, w, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev), torch.randn(n, device=dev)
        y, dy = torch.empty(m, n, device=dev), torch.randn(m, n, device=dev)
        dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
        scratch = torch.empty(64 * n + 64, device=dev)
        t_f = timeit(lambda: native.linear_forward(x, w, b, y, relu=True))
        t_dx = timeit(lambda: native.linear_backward(dy, x, w, dx, None, None, scratch))
        t_dw = timeit(lambda: native.linear_backward(dy, x, w, None, dw, None, scratch))
        fl = 2.0 * m * n * k
        print("%s M=%d N=%d K=%d  fwd %.1f us (%.1f TF)  dgrad %.1f us (%.1f TF)  wgrad %.1f us (%.1f TF)  tc_launches=%d"
              % (mode, m, n, k, t_f, fl / t_f / 1e6, t_dx, fl / t_dx / 1e6, t_dw, fl / t_dw / 1e6, native.tc_gemm_launches()))


if __name__ == "__main__":
    main()

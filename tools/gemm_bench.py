#!/usr/bin/env python
"""Micro-benchmark of the library GEMMs on the training shapes (CUDA events; run with TRXL_TCGEN05=0/1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "episodic-transformer-memory-ppo_b200"))
import torch  # noqa: E402
import trxl_native as native  # noqa: E402


def timeit(fn, iters=20, reps=10):
    """Device time per call in us: `iters` calls are captured into a CUDA graph (no host launch overhead) and replayed."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cs = torch.cuda.Stream()
    with torch.cuda.stream(cs):
        native.graph_begin(cs.cuda_stream)
        for _ in range(iters):
            fn()
        g = native.graph_end(cs.cuda_stream)
        native.graph_launch(g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cs)
        for _ in range(reps):
            native.graph_launch(g)
        b.record(cs)
        torch.cuda.synchronize()
    native.graph_destroy(g)
    return a.elapsed_time(b) / (iters * reps) * 1e3


def main():
    dev = "cuda:0"
    mode = "simt" if os.environ.get("TRXL_TCGEN05", "1") == "0" else "tcgen05+tma"
    for m, n, k in ((2048, 256, 256), (2048, 384, 256), (2048, 256, 3136), (2048, 768, 256), (300, 256, 256), (2048, 64, 256),
                    (16384, 384, 384), (16384, 512, 512)):
        x, w, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev), torch.randn(n, device=dev)
        y, dy = torch.empty(m, n, device=dev), torch.randn(m, n, device=dev)
        dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
        scratch = torch.empty(32 * 74 * 4096, device=dev)
        t_f = timeit(lambda: native.linear_forward(x, w, b, y, relu=True))
        t_dx = timeit(lambda: native.linear_backward(dy, x, w, dx, None, None, scratch))
        t_dw = timeit(lambda: native.linear_backward(dy, x, w, None, dw, None, scratch))
        fl = 2.0 * m * n * k
        native.linear_forward(x, w, b, y, relu=True)
        native.linear_backward(dy, x, w, dx, dw, db, scratch)
        torch.cuda.synchronize()
        xd, wd, dyd = x.double(), w.double(), dy.double()
        ref_y, ref_dx, ref_dw = torch.relu(xd @ wd.t() + b.double()), dyd @ wd, dyd.t() @ xd
        err = [float((a.double() - r).abs().max() / r.abs().max()) for a, r in ((y, ref_y), (dx, ref_dx), (dw, ref_dw))]
        print("   max err / max |ref|: fwd %.1e dgrad %.1e wgrad %.1e" % tuple(err))
        print("%s M=%d N=%d K=%d  fwd %.1f us (%.1f TF)  dgrad %.1f us (%.1f TF)  wgrad %.1f us (%.1f TF)  tc_launches=%d"
              % (mode, m, n, k, t_f, fl / t_f / 1e6, t_dx, fl / t_dx / 1e6, t_dw, fl / t_dw / 1e6, native.tc_gemm_launches()))


if __name__ == "__main__":
    main()

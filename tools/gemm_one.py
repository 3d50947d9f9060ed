#!/usr/bin/env python
"""One GEMM shape through the library, a few launches (for ncu): python tools/gemm_one.py M N K [fwd|dgrad|wgrad]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "episodic-transformer-memory-ppo_b200"))
import torch  # noqa: E402
import trxl_native as native  # noqa: E402
m, n, k = (int(v) for v in sys.argv[1:4])
kind = sys.argv[4] if len(sys.argv) > 4 else "fwd"
dev = "cuda:0"
x, w, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev), torch.randn(n, device=dev)
y, dy = torch.empty(m, n, device=dev), torch.randn(m, n, device=dev)
dx, dw = torch.empty_like(x), torch.empty_like(w)
scratch = torch.empty(1 << 20, device=dev)
for _ in range(6):
    if kind == "fwd":
        native.linear_forward(x, w, b, y, relu=True)
    elif kind == "dgrad":
        native.linear_backward(dy, x, w, dx, None, None, scratch)
    else:
        native.linear_backward(dy, x, w, None, dw, None, scratch)
torch.cuda.synchronize()
print("done")

#!/usr/bin/env python
"""Stage-by-stage check of the tcgen05 encoder against torch (GPU box): dumps the max error of every intermediate
buffer of the encoder workspace (activations, pre-activation gradients) and of the six parameter gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "episodic-transformer-memory-ppo_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from test_gpu_tc_conv import _model
    C, H, W = 4, 84, 84
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    model = _model(C, H, W)
    torch.manual_seed(1)
    obs = torch.rand(n, C, H, W)
    dfeat = torch.randn(n, model._feat_dim)
    convs = [torch.nn.Conv2d(C, 32, 8, 4), torch.nn.Conv2d(32, 64, 4, 2), torch.nn.Conv2d(64, 64, 3, 1)]
    for c, src in zip(convs, (model.conv1, model.conv2, model.conv3)):
        c.weight.data.copy_(src.weight.detach().cpu())
        c.bias.data.copy_(src.bias.detach().cpu())
    zs, ys, x = [], [], obs
    for c in convs:
        z = c(x)
        z.retain_grad()
        x = torch.relu(z)
        zs.append(z)
        ys.append(x)
    feat = x.reshape(n, -1)
    feat.backward(dfeat)
    h1, w1 = ys[0].shape[-2:]
    h2, w2 = ys[1].shape[-2:]
    h3, w3 = ys[2].shape[-2:]

    f = model.encode_train(obs.cuda(), None, n)
    model._grad_arena.fill_(7.0)
    model.encode_backward(n, H, W, dfeat.cuda())
    torch.cuda.synchronize()
    ws = model._enc_ws(n, H, W)[0].cpu().numpy()

    def a64(v):
        return (v + 63) // 64 * 64
    off = [0]

    def take(k):
        o = off[0]
        off[0] += a64(k)
        return ws[o:o + k]
    m1, m2, m3 = n * h1 * w1, n * h2 * w2, n * h3 * w3
    x0 = [take(n * H * W * 4) for _ in range(2)]
    y1 = [take(m1 * 32) for _ in range(2)]
    y2 = [take(m2 * 64) for _ in range(2)]
    y3 = [take(m3 * 64) for _ in range(2)]
    take(m3 * 64)
    d3 = [take(m3 * 64) for _ in range(2)]
    d2 = [take(m2 * 64) for _ in range(2)]
    d1 = [take(m1 * 32) for _ in range(2)]

    def nhwc(t):
        return t.detach().permute(0, 2, 3, 1).contiguous().numpy().reshape(-1)

    def rep(name, got, ref):
        print("%-6s max|ref| %9.4f  max err %.3e  nonzero got %d/%d" % (name, np.abs(ref).max(), np.abs(got - ref).max(),
                                                                         int((got != 0).sum()), got.size), flush=True)
    print("feat   err %.3e" % float((f.cpu() - feat.detach()).abs().max()))
    x0ref = np.zeros((n, H, W, 4), np.float32)
    x0ref[..., :C] = obs.permute(0, 2, 3, 1).numpy()
    rep("x0", x0[0] + x0[1], x0ref.reshape(-1))
    rep("y1", y1[0] + y1[1], nhwc(ys[0]))
    rep("y2", y2[0] + y2[1], nhwc(ys[1]))
    rep("y3", y3[0] + y3[1], nhwc(ys[2]))
    rep("dy3", d3[0] + d3[1], nhwc(zs[2].grad))
    rep("dy2", d2[0] + d2[1], nhwc(zs[1].grad))
    rep("dy1", d1[0] + d1[1], nhwc(zs[0].grad))
    e1 = np.abs((d1[0] + d1[1]) - nhwc(zs[0].grad)).reshape(n, h1, w1, 32).max(axis=3)
    bad = np.argwhere(e1 > 1e-3)
    print("dy1 bad pixels: %d of %d; by class (py,px): %s" % (len(bad), e1.size, {(py, px): int((e1[:, py::2, px::2] > 1e-3).sum())
                                                                                   for py in (0, 1) for px in (0, 1)}))
    print("   bad by iy", (e1 > 1e-3).sum(axis=(0, 2)), "\n   bad by ix", (e1 > 1e-3).sum(axis=(0, 1)))
    print("   bad images (first 20)", np.unique(bad[:, 0])[:20], "count", len(np.unique(bad[:, 0])))
    print("   first bad", bad[:12].tolist())
    got = [t.grad.cpu().numpy() for c in (model.conv1, model.conv2, model.conv3) for t in (c.weight, c.bias)]
    ref = [t.grad.numpy() for c in convs for t in (c.weight, c.bias)]
    for name, a, b in zip("dw1 db1 dw2 db2 dw3 db3".split(), got, ref):
        rep(name, a.reshape(-1), b.reshape(-1))
        if name.startswith("dw"):
            okm = np.abs(a - b) <= 1e-4 * max(1.0, np.abs(b).max())
            print("   fraction within tolerance %.4f; got[:6] %s ref[:6] %s" % (okm.mean(), a.reshape(-1)[:6], b.reshape(-1)[:6]))
            if okm.mean() < 1.0:
                # which input channels / kernel taps / output channels are right
                print("   ok by oc   ", np.round(okm.mean(axis=(1, 2, 3)), 2)[:16])
                print("   ok by c    ", np.round(okm.mean(axis=(0, 2, 3)), 2)[:16])
                print("   ok by ky,kx", np.round(okm.mean(axis=(0, 1)), 2).reshape(-1))


if __name__ == "__main__":
    main()

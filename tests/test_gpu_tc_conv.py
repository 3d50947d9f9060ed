"""tcgen05 (3xTF32) CNN encoder, forward + backward, against torch's fp32 CPU convolutions (the reference's encoder,
model.py:87-94, and its autograd backward).  Tolerance: 1e-4 relative to the tensor's scale (north_star fp32 parity)."""
import numpy as np
import pytest
import torch

import trxl_native as native
from model import ActorCriticModel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(C, H, W):
    cfg = {"hidden_layer_size": 64, "transformer": {"num_blocks": 1, "embed_dim": 64, "num_heads": 1, "memory_length": 4,
                                                      "positional_encoding": "", "layer_norm": "pre", "gtrxl": False}}

    class Space:
        shape = (C, H, W)
    torch.manual_seed(C * 100 + H)
    return ActorCriticModel(cfg, Space(), (3,), 8).to(DEV)


def _convs(model, C):
    convs = [torch.nn.Conv2d(C, 32, 8, 4), torch.nn.Conv2d(32, 64, 4, 2), torch.nn.Conv2d(64, 64, 3, 1)]
    for c, src in zip(convs, (model.conv1, model.conv2, model.conv3)):
        c.weight.data.copy_(src.weight.detach().cpu())
        c.bias.data.copy_(src.bias.detach().cpu())
    return convs


def _torch_reference(model, obs, dfeat):
    convs = _convs(model, obs.shape[1])
    x = obs
    for c in convs:
        x = torch.relu(c(x))
    feat = x.reshape(obs.shape[0], -1)
    feat.backward(dfeat)
    return feat.detach().numpy(), [t.grad.numpy() for c in convs for t in (c.weight, c.bias)]


def _engine_relu_masks(model, n, H, W):
    """(y1 > 0, y2 > 0, y3 > 0) as NCHW bool tensors, read from the encoder workspace (layout: csrc/tc_conv.cu carve())."""
    ws = model._enc_ws(n, H, W)[0].cpu()
    h1, w1 = (H - 8) // 4 + 1, (W - 8) // 4 + 1
    h2, w2 = (h1 - 4) // 2 + 1, (w1 - 4) // 2 + 1
    h3, w3 = h2 - 2, w2 - 2
    a64 = lambda v: (v + 63) // 64 * 64                                          # noqa: E731
    off = 2 * a64(n * H * W * 4)
    masks = []
    for (h, w, c) in ((h1, w1, 32), (h2, w2, 64), (h3, w3, 64)):
        k = n * h * w * c
        masks.append((ws[off:off + k].view(n, h, w, c) > 0).permute(0, 3, 1, 2))
        off += 2 * a64(k)
    return masks


def _reference_backward_with_masks(model, obs, dfeat, masks):
    """The reference's backward in float64, with the ReLU masks of the engine's own forward.  An activation whose
    pre-activation sits at rounding-noise level can land on either side of zero in two correct fp32 implementations, and
    one such pixel moves a whole filter's gradient by |dy| * |x|; fixing the masks makes the comparison well-posed."""
    from torch.nn import grad as G
    convs = [c.double() for c in _convs(model, obs.shape[1])]
    strides = (4, 2, 1)
    xs = [obs.double()]
    for c, m in zip(convs, masks):
        xs.append(torch.where(m, c(xs[-1]), torch.zeros((), dtype=torch.float64)).detach())
    dz = dfeat.double().view_as(xs[3]) * masks[2]
    out = [None] * 6
    for layer in (2, 1, 0):
        c = convs[layer]
        out[2 * layer] = G.conv2d_weight(xs[layer], c.weight.shape, dz, stride=strides[layer]).numpy()
        out[2 * layer + 1] = dz.sum(dim=(0, 2, 3)).numpy()
        if layer > 0:
            dz = G.conv2d_input(xs[layer].shape, c.weight.detach(), dz, stride=strides[layer]) * masks[layer - 1]
    return out


@pytest.mark.parametrize("C,H,W,n,indexed", [(4, 84, 84, 5, False), (4, 84, 84, 300, True), (3, 64, 72, 7, False), (1, 45, 38, 9, True)])
def test_tc_encoder_forward_backward(C, H, W, n, indexed):
    model = _model(C, H, W)
    assert model._tc_encoder
    g = torch.Generator().manual_seed(n)
    pool = torch.rand((n + 11, C, H, W), generator=g)
    sidx = torch.randperm(n + 11, generator=g)[:n] if indexed else None
    obs = pool[sidx] if indexed else pool[:n]
    dfeat = torch.randn((n, model._feat_dim), generator=g)
    ref_feat, _ = _torch_reference(model, obs, dfeat)

    launches0 = native.launch_count()
    feat = model.encode_train(pool.to(DEV), sidx.to(DEV) if indexed else None, n)
    np.testing.assert_allclose(feat.cpu().numpy(), ref_feat, atol=1e-4 * max(1.0, float(np.abs(ref_feat).max())))
    model._grad_arena.fill_(7.0)                      # the conv slices must be overwritten, not accumulated
    model.encode_backward(n, H, W, dfeat.to(DEV))
    torch.cuda.synchronize()
    assert native.launch_count() > launches0
    ref_grads = _reference_backward_with_masks(model, obs, dfeat, _engine_relu_masks(model, n, H, W))
    got = [t.grad.cpu().numpy() for c in (model.conv1, model.conv2, model.conv3) for t in (c.weight, c.bias)]
    for name, a, b in zip("w1 b1 w2 b2 w3 b3".split(), got, ref_grads):
        np.testing.assert_allclose(a, b, atol=1e-4 * max(1.0, float(np.abs(b).max())), err_msg=name)


def test_tc_encoder_is_differentiable_through_forward():
    """The reference API (model.forward under autograd) reaches the same kernels through the autograd bridge."""
    C, H, W, n = 4, 84, 84, 3
    model = _model(C, H, W)
    obs = torch.rand((n, C, H, W))
    feat = model.encode(obs.to(DEV).requires_grad_(False))
    assert feat.requires_grad
    dfeat = torch.randn((n, model._feat_dim))
    model._grad_arena.zero_()
    feat.backward(dfeat.to(DEV))
    _, ref_grads = _torch_reference(model, obs, dfeat)
    got = [t.grad.cpu().numpy() for c in (model.conv1, model.conv2, model.conv3) for t in (c.weight, c.bias)]
    for a, b in zip(got, ref_grads):
        np.testing.assert_allclose(a, b, atol=1e-4 * max(1.0, float(np.abs(b).max())))

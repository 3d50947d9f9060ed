"""CPU pin of the implicit-GEMM geometry the tcgen05 encoder kernels walk (csrc/tc_conv_geom.h): a scalar emulation built
from the very same tables must reproduce torch's conv forward / backward (reference model.py:87-94 + autograd)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT

FP = ctypes.POINTER(ctypes.c_float)


def _p(a):
    return a.ctypes.data_as(FP)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "tc_conv_emu.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(PKG, "csrc"), os.path.join(ROOT, "tests", "emu", "tc_conv_emu.cpp"),
                    "-o", out], check=True)
    return ctypes.CDLL(out)


@pytest.mark.parametrize("C,H,W,n", [(4, 84, 84, 2), (3, 64, 72, 2), (1, 45, 38, 3)])
def test_emulated_implicit_gemm_matches_torch(emu, C, H, W, n):
    torch.manual_seed(C * 1000 + H)
    convs = [torch.nn.Conv2d(C, 32, 8, 4), torch.nn.Conv2d(32, 64, 4, 2), torch.nn.Conv2d(64, 64, 3, 1)]
    obs = torch.rand(n, C, H, W)
    x = obs
    for c in convs:
        x = torch.relu(c(x))
    feat = x.reshape(n, -1)
    dfeat = torch.randn_like(feat)
    feat.backward(dfeat)
    params = [t for c in convs for t in (c.weight, c.bias)]
    ws = [t.detach().numpy().copy() for t in params]
    assert emu.emu_feature_size(C, H, W) == feat.shape[1]
    got = np.zeros(tuple(feat.shape), np.float32)
    o = obs.numpy().copy()
    assert emu.emu_forward(C, H, W, ctypes.c_longlong(n), _p(o), *[_p(w) for w in ws], _p(got)) == 0
    np.testing.assert_allclose(got, feat.detach().numpy(), atol=1e-5)
    grads = [np.zeros_like(w) for w in ws]
    df = dfeat.numpy().copy()
    assert emu.emu_backward(C, H, W, ctypes.c_longlong(n), _p(o), *[_p(w) for w in ws], _p(df), *[_p(g) for g in grads]) == 0
    for g, t in zip(grads, params):
        np.testing.assert_allclose(g, t.grad.numpy(), atol=2e-5 * max(1.0, float(t.grad.abs().max())))


def test_unsupported_shapes_are_reported(emu):
    assert emu.emu_feature_size(5, 84, 84) == -1          # more than 4 channels: the caller keeps the cuDNN encoder
    assert emu.emu_feature_size(4, 30, 84) == -1

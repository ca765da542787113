"""Parity of the backward-pass kernels (through block-level C-ABI entry points) against PyTorch autograd of the
same op in fp32.  fp16 operands / fp32 accumulation => 3e-3 of the output scale against the fp32 reference and
~1e-5 against a reference evaluated on the fp16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def gen(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("B,T,Cin,Cout", [(1, 64, 128, 128), (1, 300, 256, 256), (3, 435, 256, 1024), (2, 77, 80, 256),
                                          (4, 148, 512, 768), (2, 200, 1024, 256), (2, 129, 64, 256),
                                          (2, 130, 256, 160), (16, 435, 256, 128)])
def test_wgrad_dense(B, T, Cin, Cout):
    """dW = X^T dY with both operands fed MN-major straight from the row-major activations"""
    import gpu_util as G
    X, dY = gen(B, T, Cin, seed=1), gen(B, T, Cout, seed=2)
    out = G.wgrad(X, dY)[0]
    ref16 = X.half().double().reshape(-1, Cin).T @ dY.half().double().reshape(-1, Cout)
    ref32 = X.double().reshape(-1, Cin).T @ dY.double().reshape(-1, Cout)
    assert torch.isfinite(out).all()
    assert rel_err(out.double(), ref16) < 2e-5, rel_err(out.double(), ref16)
    assert rel_err(out.double(), ref32) < 3e-3


def test_wgrad_concat():
    """Dense over a concat [x ; ctx] (modules/attention.py:410): rows of dW from two tensors"""
    import gpu_util as G
    B, T = 2, 300
    X, X2, dY = gen(B, T, 256, seed=3), gen(B, T, 256, seed=4), gen(B, T, 256, seed=5)
    out = G.wgrad(X, dY, X2=X2)[0]
    cat = torch.cat([X, X2], -1).half().double().reshape(-1, 512)
    ref = cat.T @ dY.half().double().reshape(-1, 256)
    assert rel_err(out.double(), ref) < 2e-5


@pytest.mark.parametrize("B,T,Cin,Cout", [(2, 100, 80, 256), (3, 67, 256, 256), (2, 148, 512, 512)])
def test_wgrad_conv(B, T, Cin, Cout):
    """Conv1D k=5 'same' weight gradient: row-shifted A tiles, zero fill at the utterance edges"""
    import gpu_util as G
    X, dY = gen(B, T, Cin, seed=6), gen(B, T, Cout, seed=7)
    out = G.wgrad(X, dY, taps=5)
    xh = X.half().double().requires_grad_(False)
    W = torch.zeros(5, Cin, Cout, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv1d(xh.transpose(1, 2), W.permute(2, 1, 0), padding=2).transpose(1, 2)
    (y * dY.half().double()).sum().backward()
    assert rel_err(out.double(), W.grad) < 2e-5, rel_err(out.double(), W.grad)

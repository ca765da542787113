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


def _attn_ref(q, k, v, q_len, k_len, H, causal):
    """MultiHeadScaledProductAttention core (modules/attention.py:217-246) in float64 for autograd."""
    B, Tq, A = q.shape
    Tk = k.shape[1]
    hd = A // H
    qh = q.reshape(B, Tq, H, hd).transpose(1, 2)
    kh = k.reshape(B, Tk, H, hd).transpose(1, 2)
    vh = v.reshape(B, Tk, H, hd).transpose(1, 2)
    logits = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    ar_q, ar_k = torch.arange(Tq), torch.arange(Tk)
    mask = (ar_k[None, None, :] < k_len[:, None, None]) & (ar_q[None, :, None] < q_len[:, None, None])
    if causal:
        mask = mask & (ar_k[None, None, :] <= ar_q[None, :, None])
    logits = torch.where(mask[:, None], logits, torch.full_like(logits, -2.0 ** 32 + 1))
    ali = torch.softmax(logits, dim=3)
    return (ali @ vh).transpose(1, 2).reshape(B, Tq, A)


@pytest.mark.parametrize("B,H,Tq,Tk,causal", [(2, 4, 128, 128, False), (2, 4, 128, 128, True), (3, 4, 435, 148, False),
                                              (3, 4, 435, 435, True), (2, 2, 70, 33, False), (2, 4, 300, 300, True),
                                              (2, 4, 148, 148, False)])
def test_attention_backward(B, H, Tq, Tk, causal):
    """dQ, dK, dV incl. fully masked query rows (uniform attention -> gradient reaches V only, also at padded keys)"""
    import gpu_util as G
    A = H * 64
    q, k, v, do = gen(B, Tq, A, seed=1), gen(B, Tk, A, seed=2), gen(B, Tk, A, seed=3), gen(B, Tq, A, seed=4)
    g = torch.Generator().manual_seed(5)
    q_len = torch.randint(max(1, Tq // 2), Tq + 1, (B,), generator=g)
    q_len[0] = Tq
    k_len = q_len.clone() if causal else torch.randint(max(1, Tk // 2), Tk + 1, (B,), generator=g)
    if not causal:
        k_len[0] = Tk
    dq, dk, dv = G.attention_bwd(q, k, v, do, q_len, k_len, H, causal)
    qd, kd, vd = (t.half().double().requires_grad_(True) for t in (q, k, v))
    ctx = _attn_ref(qd, kd, vd, q_len, k_len, H, causal)
    (ctx * do.half().double()).sum().backward()
    for name, got, ref in (("dq", dq, qd.grad), ("dk", dk, kd.grad), ("dv", dv, vd.grad)):
        assert torch.isfinite(got).all(), name
        err = rel_err(got.double(), ref)
        assert err < 6e-3, (name, err)     # fp16 P / dS / outputs, fp32 accumulation

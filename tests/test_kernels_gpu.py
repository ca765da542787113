"""Parity of the individual sm_100a kernels (through the block-level C-ABI entry points) against plain
fp32 PyTorch references of the same op.  Tolerances: fp16 operands / fp32 accumulate => 2e-3 relative of
the output scale; split-fp16 => 2e-5."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vaenar_oracle as O  # noqa: E402


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def gen(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("M,K,N,block_n", [(128, 64, 128, 128), (300, 256, 256, 128), (1000, 512, 768, 256),
                                           (77, 80, 256, 128), (435, 256, 160, 128), (260, 1024, 256, 256),
                                           (129, 128, 80, 128)])
def test_dense_plain(M, K, N, block_n):
    import gpu_util as G
    A, W, b = gen(M, K, seed=1), gen(K, N, seed=2, scale=1 / math.sqrt(K)), gen(N, seed=3)
    out = G.dense(A, W, b, act=1, block_n=block_n)
    ref16 = torch.relu(A.half().float() @ W.half().float() + b)
    ref32 = torch.relu(A @ W + b)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref16) < 2e-5, rel_err(out, ref16)
    assert rel_err(out, ref32) < 3e-3, rel_err(out, ref32)


@pytest.mark.parametrize("with_res", [False, True])
@pytest.mark.parametrize("M,K,N", [(6000, 256, 512), (13920, 1024, 256), (5001, 320, 384)])
def test_dense_two_ctas_per_sm(M, K, N, with_res):
    """Grids deeper than one wave (> 148 tiles) of the plain epilogues run the GemmCfg<128, 2> instances: 3 pipeline stages,
    12 KB epilogue slabs, two CTAs per SM.  Same numerics as the one-CTA-per-SM instances (ragged last row tile, K not a
    multiple of 64 * stages, N tail)."""
    import gpu_util as G
    A, W = gen(M, K, seed=11), gen(K, N, seed=12, scale=1 / math.sqrt(K))
    b = gen(N, seed=13) if with_res else None
    res = gen(M, N, seed=14) if with_res else None
    out = G.dense(A, W, b, residual=res, block_n=128)
    ref16 = A.half().float() @ W.half().float()
    if with_res:
        ref16 = ref16 + b + res
    assert torch.isfinite(out).all()
    assert rel_err(out, ref16) < 2e-5, rel_err(out, ref16)


@pytest.mark.parametrize("split_cluster", [False, True])
@pytest.mark.parametrize("M,K,N", [(300, 512, 256), (200, 1024, 256), (148, 768, 512), (131, 1024, 512)])
def test_dense_layernorm(M, K, N, split_cluster):
    if N == 512 and not split_cluster:
        pytest.skip("N=512 LayerNorm is only built as the 2-CTA cluster split (2 x 256 columns)")
    """GEMM + bias + residual + LayerNorm epilogue; split_cluster: N split over a 2-CTA cluster with the
    row statistics exchanged through distributed shared memory."""
    import gpu_util as G
    A, W, b = gen(M, K, seed=4), gen(K, N, seed=5, scale=1 / math.sqrt(K)), gen(N, seed=6)
    res, gamma, beta = gen(M, N, seed=7), 1 + 0.1 * gen(N, seed=8), 0.1 * gen(N, seed=9)
    out = G.dense(A, W, b, residual=res, gamma=gamma, beta=beta, ln=True, block_n=N // 2 if split_cluster else N)
    x = A.half().float() @ W.half().float() + b + res
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    ref = (x - mean) / torch.sqrt(var + 1e-3) * gamma + beta
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 5e-5, rel_err(out, ref)


@pytest.mark.parametrize("M,K,N", [(500, 256, 80), (300, 256, 256)])
def test_dense_split_fp16(M, K, N):
    """split-fp16 operands (hi, lo, hi) x (hi, hi, lo): ~fp32 accuracy on the tensor cores"""
    import gpu_util as G
    A, W, b = gen(M, K, seed=10), gen(K, N, seed=11, scale=1 / math.sqrt(K)), gen(N, seed=12)
    out = G.dense(A, W, b, split=True, block_n=128)
    ref = (A.double() @ W.double() + b.double()).float()
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)


@pytest.mark.parametrize("B,T,Cin,Cout,split", [(3, 148, 512, 512, False), (2, 301, 80, 256, True),
                                                (4, 130, 256, 256, True), (1, 7, 256, 256, False)])
def test_conv1d_same(B, T, Cin, Cout, split):
    """Conv1D k=5 'same' as an implicit GEMM with TMA zero-fill at the sequence edges (modules/utils.py:56-85)."""
    import gpu_util as G
    X, W, b = gen(B, T, Cin, seed=13), gen(5, Cin, Cout, seed=14, scale=1 / math.sqrt(5 * Cin)), gen(Cout, seed=15)
    out = G.conv1d(X, W, b, act=2, split=split)
    Xr, Wr = (X, W) if split else (X.half().float(), W.half().float())
    ref = torch.tanh(torch.nn.functional.conv1d(torch.nn.functional.pad(Xr.transpose(1, 2), (2, 2)),
                                                Wr.permute(2, 1, 0), b).transpose(1, 2))
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < 3e-5, rel_err(out, ref)


def ref_attention(q, k, v, q_len, k_len, H, causal):
    B, Tq, A = q.shape
    Tk = k.shape[1]
    qh = q.reshape(B, Tq, H, 64).transpose(1, 2)
    kh = k.reshape(B, Tk, H, 64).transpose(1, 2)
    vh = v.reshape(B, Tk, H, 64).transpose(1, 2)
    logits = qh @ kh.transpose(-1, -2) / 8.0
    mask = O.sequence_mask(k_len, Tk)[:, None, :] & O.sequence_mask(q_len, Tq)[:, :, None]
    if causal:
        mask = mask & torch.ones(Tq, Tk, dtype=torch.bool).tril()[None]
    logits = torch.where(mask[:, None], logits, torch.full_like(logits, O.MASK_FILL))
    ali = torch.softmax(logits, dim=3)
    return (ali @ vh).transpose(1, 2).reshape(B, Tq, A), ali


@pytest.mark.parametrize("B,H,Tq,Tk,causal", [(2, 4, 128, 128, True), (3, 4, 435, 435, True), (3, 4, 435, 148, False),
                                              (2, 2, 50, 300, False), (1, 4, 600, 600, True), (2, 4, 148, 148, False),
                                              (2, 4, 448, 448, True), (3, 4, 100, 100, True), (2, 4, 320, 320, True),
                                              (16, 4, 336, 336, True), (2, 4, 449, 449, True)])
@pytest.mark.parametrize("want_ali", [False, True])
def test_attention(B, H, Tq, Tk, causal, want_ali):
    """MultiHeadScaledProductAttention core incl. ragged lengths and fully masked (uniform) rows."""
    import gpu_util as G
    q, k, v = gen(B, Tq, H * 64, seed=20), gen(B, Tk, H * 64, seed=21), gen(B, Tk, H * 64, seed=22)
    g = torch.Generator().manual_seed(23)
    q_len = torch.randint(max(1, Tq // 2), Tq + 1, (B,), generator=g)
    q_len[0] = Tq
    k_len = q_len.clone() if causal else torch.randint(max(1, Tk // 2), Tk + 1, (B,), generator=g)
    if not causal:
        k_len[0] = Tk
    ctx, ali = G.attention(q, k, v, q_len, k_len, H, causal, want_ali)
    rc, ra = ref_attention(q.half().float(), k.half().float(), v.half().float(), q_len, k_len, H, causal)
    assert torch.isfinite(ctx).all()
    assert rel_err(ctx, rc) < 3e-3, rel_err(ctx, rc)
    if want_ali:
        assert float((ali - ra).abs().max()) < 2e-5
        # fully masked rows are uniform over ALL Tk keys (attention.py:240-242)
        b = int(torch.argmin(q_len))
        if q_len[b] < Tq:
            assert torch.allclose(ali[b, :, int(q_len[b]):, :], torch.full((1,), 1.0 / Tk), rtol=1e-5, atol=0)

"""Helpers for the GPU parity tests: ctypes access to the block-level entry points of the C ABI."""
import ctypes

import torch

from vaenar_tts_b200 import _lib
from vaenar_tts_b200._lib import check


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_WS = {}


def workspace(nbytes=512 << 20):
    dev = torch.cuda.current_device()
    if dev not in _WS or _WS[dev].numel() < nbytes:
        _WS[dev] = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    return _WS[dev]


def dense(A, W, bias=None, residual=None, gamma=None, beta=None, act=0, ln=False, split=False, block_n=128):
    lib = _lib.load()
    A, W = A.cuda().float().contiguous(), W.cuda().float().contiguous()
    M, K = A.shape
    N = W.shape[1]
    cu = lambda t: t.cuda().float().contiguous() if t is not None else None
    bias, residual, gamma, beta = cu(bias), cu(residual), cu(gamma), cu(beta)
    out = torch.full((M, N), float("nan"), device="cuda")
    ws = workspace()
    check(lib.vaenar_test_dense(_p(A), _p(W), _p(bias), _p(residual), _p(gamma), _p(beta), M, K, N, act, int(ln),
                                int(split), block_n, _p(out), _p(ws), ws.numel(), _stream()))
    torch.cuda.synchronize()
    return out.cpu()


def conv1d(X, W, bias, act=0, split=False):
    lib = _lib.load()
    X, W, bias = X.cuda().float().contiguous(), W.cuda().float().contiguous(), bias.cuda().float().contiguous()
    B, T, Cin = X.shape
    taps, _, Cout = W.shape
    out = torch.full((B, T, Cout), float("nan"), device="cuda")
    ws = workspace()
    check(lib.vaenar_test_conv1d(_p(X), _p(W), _p(bias), B, T, Cin, Cout, taps, act, int(split), _p(out), _p(ws),
                                 ws.numel(), _stream()))
    torch.cuda.synchronize()
    return out.cpu()


def attention(q, k, v, q_len, k_len, H, causal, want_ali=True):
    lib = _lib.load()
    q, k, v = (t.cuda().float().contiguous() for t in (q, k, v))
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    ql = q_len.cuda().int().contiguous()
    kl = k_len.cuda().int().contiguous()
    ctx = torch.full((B, Tq, H * 64), float("nan"), device="cuda")
    ali = torch.full((B, H, Tq, Tk), float("nan"), device="cuda") if want_ali else None
    ws = workspace()
    check(lib.vaenar_test_attention(_p(q), _p(k), _p(v), _p(ql), _p(kl), B, H, Tq, Tk, int(causal), _p(ctx), _p(ali),
                                    _p(ws), ws.numel(), _stream()))
    torch.cuda.synchronize()
    return ctx.cpu(), (ali.cpu() if ali is not None else None)


def wgrad(X, dY, X2=None, taps=1):
    """dW[taps, Cin(+Cin2), Cout] of a Dense (taps=1) / Conv1D 'same' layer from X [B,T,Cin], dY [B,T,Cout]."""
    lib = _lib.load()
    X, dY = X.cuda().float().contiguous(), dY.cuda().float().contiguous()
    X2 = X2.cuda().float().contiguous() if X2 is not None else None
    B, T, Cin = X.shape
    Cin2 = X2.shape[2] if X2 is not None else 0
    Cout = dY.shape[2]
    out = torch.full((taps, Cin + Cin2, Cout), float("nan"), device="cuda")
    ws = workspace()
    check(lib.vaenar_test_wgrad(_p(X), _p(X2), _p(dY), B, T, Cin, Cin2, Cout, taps, _p(out), _p(ws), ws.numel(),
                                _stream()))
    torch.cuda.synchronize()
    return out.cpu()


def attention_bwd(q, k, v, dctx, q_len, k_len, H, causal):
    lib = _lib.load()
    q, k, v, dctx = (t.cuda().float().contiguous() for t in (q, k, v, dctx))
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    ql = q_len.cuda().int().contiguous()
    kl = k_len.cuda().int().contiguous()
    dq = torch.full_like(q, float("nan"))
    dk = torch.full_like(k, float("nan"))
    dv = torch.full_like(v, float("nan"))
    ws = workspace()
    check(lib.vaenar_test_attention_bwd(_p(q), _p(k), _p(v), _p(dctx), _p(ql), _p(kl), B, H, Tq, Tk, int(causal), _p(dq),
                                        _p(dk), _p(dv), _p(ws), ws.numel(), _stream()))
    torch.cuda.synchronize()
    return dq.cpu(), dk.cpu(), dv.cpu()

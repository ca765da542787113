"""Parity of the fused per-CrossAttentionBLK row kernel (csrc/xblk_fused.cuh) through the block-level C-ABI entry
point ``vaenar_xblk_stack_fwd``: against the oracle's ``cross_attention_blk`` chain (modules/attention.py:436-452)
and against the per-op launch chain it replaces.  Tolerances: fp16 operands / fp32 accumulate, LayerNorm outputs of
unit scale => 4e-3 absolute vs the fp32 oracle, 2e-3 between the two CUDA paths."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vaenar_oracle as O  # noqa: E402
from oracle.hparams import LJHPS as OHPS  # noqa: E402


@pytest.fixture(scope="module")
def model_and_params():
    from vaenar_tts_b200 import VAENAR, LJHPS
    P = O.init_params(OHPS, seed=11, zero_init_std=0.02)
    model = VAENAR(LJHPS, device="cuda:0")
    model.load_state_dict(P)
    return model, P


def _oracle_stack(P, prefix, nblk, x, mem, q_len, t_len):
    alis = []
    for i in range(nblk):
        x, a = O.cross_attention_blk(P, f"{prefix}.{i}", x, mem, q_len, t_len, 4)
        alis.append(a)
    return x, torch.stack(alis)


CASES = [
    # module, oracle prefix, B, T, Tt, q_len, t_len
    ("decoder", "decoder.attentions", 2, 200, 148, [200, 131], [148, 77]),
    ("posterior", "posterior.attentions", 3, 435, 148, [435, 300, 17], [148, 100, 9]),
    (("prior", 3), "prior.glow.3.affine_coupling.net.attentions", 2, 128, 64, [128, 64], [64, 33]),
    (("prior", 0), "prior.glow.0.affine_coupling.net.attentions", 1, 77, 152, [50], [152]),
    ("decoder", "decoder.attentions", 2, 130, 19, [130, 1], [19, 3]),
    ("decoder", "decoder.attentions", 1, 256, 192, [256], [192]),
    ("decoder", "decoder.attentions", 2, 140, 130, [140, 90], [130, 129]),
]


@pytest.mark.parametrize("module,prefix,B,T,Tt,q_len,t_len", CASES)
def test_fused_row_kernel_vs_oracle_and_unfused(model_and_params, module, prefix, B, T, Tt, q_len, t_len):
    from vaenar_tts_b200 import _lib
    model, P = model_and_params
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, 256, generator=g)
    mem = torch.randn(B, Tt, 512, generator=g)
    ql, tl = torch.tensor(q_len, dtype=torch.int32), torch.tensor(t_len, dtype=torch.int32)
    nblk = 2
    with torch.no_grad():
        ref, ref_ali = _oracle_stack(P, prefix, nblk, x, mem, ql, tl)
    want_ali = module == "decoder"
    outs = {}
    for fused in (1, 0):
        lib.vaenar_set_fused(fused)
        try:
            y, ali = model.cross_attention_blocks(module, x, mem, ql, tl, return_alignments=want_ali)
            torch.cuda.synchronize()
        finally:
            lib.vaenar_set_fused(1)
        outs[fused] = (y.cpu(), ali.cpu() if ali is not None else None)
    y1, a1 = outs[1]
    y0, a0 = outs[0]
    assert torch.isfinite(y1).all()
    err_ref = float((y1 - ref).abs().max())
    err_old = float((y0 - ref).abs().max())
    err_paths = float((y1 - y0).abs().max())
    print(f"{module} B{B} T{T} Tt{Tt}: fused-oracle {err_ref:.2e}  unfused-oracle {err_old:.2e}  fused-unfused {err_paths:.2e}")
    assert err_ref < 8e-3, err_ref
    assert err_ref < 2.0 * err_old + 1e-3, (err_ref, err_old)
    assert float((y1 - ref).abs().mean()) < 6e-4
    if want_ali:
        # row-stochastic, exact zeros on masked keys of live rows, uniform 1/Tt on fully masked rows (attention.py:234-246)
        assert torch.isfinite(a1).all()
        assert float((a1.sum(-1) - 1).abs().max()) < 1e-3
        assert float((a1 - ref_ali).abs().max()) < 3e-3
        for b in range(B):
            if t_len[b] < Tt:
                assert float(a1[:, b, :, :q_len[b], t_len[b]:].abs().max()) == 0.0
            if q_len[b] < T:
                dead = a1[:, b, :, q_len[b]:, :]
                assert float((dead - 1.0 / Tt).abs().max()) < 1e-7

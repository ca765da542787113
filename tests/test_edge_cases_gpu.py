"""Edge cases of the CUDA path against the oracle: ragged / minimal / long inputs, every reduction factor,
single-utterance batches, sequences longer than one attention key block, DataBaker hparams."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vaenar_oracle as O  # noqa: E402
from oracle.hparams import LJHPS as OLJ, DataBakerHPS as ODB  # noqa: E402
from test_model_gpu import make_model, masked_mae, rel, MEL_MAE_TOL, REL_TOL  # noqa: E402


def _params(ohps, seed):
    P = O.init_params(ohps, seed=seed, zero_init_std=0.02)
    O.randomize_bn_stats(P, seed=seed + 1)
    return P


@pytest.mark.parametrize("B,Tt,Tm,rf", [(1, 5, 9, 2), (2, 1, 2, 2), (3, 33, 131, 1), (2, 40, 203, 3), (2, 50, 257, 4),
                                        (5, 64, 301, 5), (1, 200, 1200, 2), (7, 17, 77, 2)])
def test_inference_shapes(B, Tt, Tm, rf):
    P = _params(OLJ, 21)
    m = make_model(OLJ, P)
    texts, mels, t_len, m_len = O.synthetic_batch(OLJ, B, Tt, Tm, rf=rf, seed=B * 1000 + Tt)
    Tz = int(((m_len + rf - 1) // rf).max())
    eps = torch.randn(B, Tz, 128, generator=torch.Generator().manual_seed(Tm))
    mel, ali = m.inference(texts, m_len, t_len, reduction_factor=rf, epsilon=eps)
    with torch.no_grad():
        ref, ref_ali, aux = O.vaenar_inference(P, OLJ, texts, m_len, t_len, rf, eps)
    assert mel.shape == ref.shape
    assert torch.isfinite(mel).all()
    assert masked_mae(mel, ref, torch.minimum(m_len, torch.tensor(ref.shape[1]))) <= MEL_MAE_TOL
    for k in ref_ali:
        assert float((ali[k].cpu() - ref_ali[k]).abs().max()) < 5e-3


def test_minimal_lengths_and_equal_lengths():
    """utterances of length 1 next to full-length ones; and a batch without any padding"""
    P = _params(OLJ, 23)
    m = make_model(OLJ, P)
    B, Tt, Tm = 4, 12, 40
    texts, mels, t_len, m_len = O.synthetic_batch(OLJ, B, Tt, Tm, seed=3)
    t_len = torch.tensor([Tt, 1, Tt, 2], dtype=torch.int32)
    m_len = torch.tensor([Tm, 2, Tm, 3], dtype=torch.int32)
    for tl, ml in ((t_len, m_len), (torch.full((B,), Tt, dtype=torch.int32), torch.full((B,), Tm, dtype=torch.int32))):
        eps = torch.randn(B, 20, 128, generator=torch.Generator().manual_seed(5))
        mel, _ = m.inference(texts, ml, tl, reduction_factor=2, epsilon=eps)
        with torch.no_grad():
            ref, _, _ = O.vaenar_inference(P, OLJ, texts, ml, tl, 2, eps)
        assert masked_mae(mel, ref, ml) <= MEL_MAE_TOL
        # padded frames are live in the reference (no masking on conv/BN/LN/FFN): compare them too
        assert float((mel.cpu() - ref).abs().max()) < 2e-2


def test_call_long_and_databaker():
    """VAENAR.call on a DataBaker-shaped batch (C4 per-GPU shape: B16, T_text 152, T_mel 640)"""
    P = _params(ODB, 25)
    m = make_model(ODB, P)
    B, Tt, Tm = 16, 152, 640
    texts, mels, t_len, m_len = O.synthetic_batch(ODB, B, Tt, Tm)
    eps = torch.randn(B, 1, Tm // 2, 128, generator=torch.Generator().manual_seed(6))
    mel, l2, kl, ll, _ = m(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len, reduction_factor=2,
                           training=False, reduce_loss=False, eps=eps)
    with torch.no_grad():
        rm, rl2, rkl, rll, _, _ = O.vaenar_call(P, ODB, texts, mels, m_len, t_len, 2, eps, reduce_loss=False)
    assert masked_mae(mel, rm, m_len) <= MEL_MAE_TOL
    assert rel(l2, rl2) < REL_TOL and rel(kl, rkl) < REL_TOL


def test_noise_generator_statistics_and_determinism():
    """The Philox N(0,1) generator that replaces tf.random.normal: mean/std/kurtosis and (seed, stream) determinism."""
    from vaenar_tts_b200 import VAENAR, LJHPS
    m = VAENAR(LJHPS, device="cuda", seed=7)
    a = m._noise((1 << 20,))
    m2 = VAENAR(LJHPS, device="cuda", seed=7)
    b = m2._noise((1 << 20,))
    c = m2._noise((1 << 20,))
    assert torch.equal(a, b) and not torch.equal(b, c)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05
    half = m._noise((4096,), stddev=0.5)
    assert abs(float(half.std()) - 0.5) < 0.03


def test_temperature_zero_is_deterministic():
    """inference.py default --temperature 0.0: epsilon = 0 (inference.py:95, prior.py:35-36)"""
    P = _params(OLJ, 27)
    m = make_model(OLJ, P)
    texts, mels, t_len, m_len = O.synthetic_batch(OLJ, 2, 10, 30)
    emb = m.text_encoder(texts, t_len, pos_step=OLJ.Common.mel_text_len_ratio / 2.0, training=False)
    z_len = (m_len + 1) // 2
    z1, lp1 = m.prior.sample(z_len, emb, t_len, training=False, temperature=0.0)
    z2, lp2 = m.prior.sample(z_len, emb, t_len, training=False, temperature=0.0)
    assert torch.equal(z1, z2)
    with torch.no_grad():
        zr, lpr = O.prior_sample(P, OLJ, torch.zeros(2, int(z_len.max()), 128), z_len, emb.cpu(), t_len)
    zmask = O.sequence_mask(z_len, z1.shape[1], torch.float32)[:, :, None]
    assert float(((z1.cpu() - zr).abs() * zmask).max()) < 2e-2
    assert rel(lp1, lpr) < REL_TOL


def test_adam_step_matches_keras_formula():
    """vaenar_adam_step vs the oracle's Keras-Adam restatement (train.py:116-117) over three steps; BatchNorm moving
    statistics must not move."""
    from vaenar_tts_b200 import VAENAR, LJHPS
    m = VAENAR(LJHPS, device="cuda", seed=3)
    p0 = m.flat_parameters().clone()
    ref_p, ref_m, ref_v = p0.cpu().double(), torch.zeros_like(p0).cpu().double(), torch.zeros_like(p0).cpu().double()
    g = torch.Generator().manual_seed(0)
    for step in (1, 2, 3):
        grads = torch.randn(p0.numel(), generator=g) * 1e-3
        m.apply_gradients(grads, step)
        ref_p, ref_m, ref_v = O.adam_update(ref_p, grads.double(), ref_m, ref_v, step)
    mask = m._trainable_mask.cpu().bool()
    got = m.flat_parameters().cpu().double()
    # fp32 kernel vs float64 restatement: one ulp of the largest parameters (gamma = 1) is 1.2e-7
    assert float((got[mask] - ref_p[mask]).abs().max()) < 5e-7
    assert float(((got[mask] - p0.cpu().double()[mask]) - (ref_p[mask] - p0.cpu().double()[mask])).abs().max()) < 5e-7
    assert torch.equal(got[~mask].float(), p0.cpu()[~mask])
    sd = m.state_dict()
    assert float(sd["decoder.postnet.conv_stack.0.bn.moving_variance"].min()) == 1.0

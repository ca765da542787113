"""The oracle restatement versus golden vectors produced by the reference's own Python sources
executed over oracle/tf_shim.py (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import vaenar_oracle as O
from golden_util import CASES, load_case, t, train_masks

TOL = dict(rtol=2e-4, atol=2e-5)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, **kw), float((a - b).abs().max())


@pytest.mark.parametrize("case", list(CASES))
def test_call_eval(case):
    hps, g, P = load_case(case)
    rf = int(g["rf"])
    with torch.no_grad():
        mel, l2, kl, ll, ali, aux = O.vaenar_call(P, hps, t(g, "texts"), t(g, "mels"), t(g, "m_len"), t(g, "t_len"),
                                                  rf, t(g, "eval_eps"), training=False, reduce_loss=True)
        close(mel, g["eval_mel"])
        close(l2, g["eval_l2"])
        close(kl, g["eval_kl"], rtol=1e-4)
        close(ll, g["eval_len"])
        for k, v in ali.items():
            close(v, g["eval_ali_" + k], atol=1e-5)
        _, l2u, klu, llu, _, _ = O.vaenar_call(P, hps, t(g, "texts"), t(g, "mels"), t(g, "m_len"), t(g, "t_len"),
                                               rf, t(g, "eval_eps"), training=False, reduce_loss=False)
        close(l2u, g["eval_l2_per"])
        close(klu, g["eval_kl_per"], rtol=1e-4)
        close(llu, g["eval_len_per"])


@pytest.mark.parametrize("case", list(CASES))
def test_inference(case):
    hps, g, P = load_case(case)
    with torch.no_grad():
        mel, ali, aux = O.vaenar_inference(P, hps, t(g, "texts"), t(g, "m_len"), t(g, "t_len"), int(g["rf"]),
                                           t(g, "inf_epsilon"))
    close(mel, g["inf_mel"])
    for k, v in ali.items():
        close(v, g["inf_ali_" + k], atol=1e-5)


@pytest.mark.parametrize("case", list(CASES))
def test_submodule_sequence(case):
    """The call sequence of inference.py:125-143 (encoder -> length predictor -> prior.sample -> ...)
    and the flow round trip prior.log_probability(prior.sample(eps)) (SURVEY.md §4 KAT 1)."""
    hps, g, P = load_case(case)
    t_len, m_len = t(g, "t_len"), t(g, "m_len")
    with torch.no_grad():
        emb = O.text_encoder(P, hps, t(g, "texts"), t_len, hps.Common.mel_text_len_ratio / 2.0)
        close(emb, g["sub_text_embd"])
        close(O.length_predictor(P, emb, t_len), g["sub_pred_len"])
        z_len = (m_len + 1) // 2
        z, logp = O.prior_sample(P, hps, t(g, "sub_epsilon"), z_len, emb, t_len)
        close(z, g["sub_z"])
        close(logp, g["sub_logp"], rtol=1e-4)
        back = O.prior_log_probability(P, hps, z, emb, z_len, t_len)
        close(back, g["sub_logp_roundtrip"], rtol=1e-4)
        close(back, logp, rtol=2e-3)


@pytest.mark.parametrize("case", list(CASES))
def test_call_train(case):
    hps, g, P = load_case(case)
    new_stats = {}
    with torch.no_grad():
        mel, l2, kl, ll, _, _ = O.vaenar_call(P, hps, t(g, "texts"), t(g, "mels"), t(g, "m_len"), t(g, "t_len"),
                                              int(g["rf"]), t(g, "train_eps"), training=True, reduce_loss=True,
                                              masks=train_masks(hps, g, "train"), new_stats=new_stats)
    close(mel, g["train_mel"])
    close(l2, g["train_l2"])
    close(kl, g["train_kl"], rtol=1e-4)
    close(ll, g["train_len"])
    assert len(new_stats) == 2 * (hps.Encoder.n_conv + hps.Decoder.post_n_conv)
    for k, v in new_stats.items():
        close(v, g["train_bnstat/" + k])


@pytest.mark.parametrize("case", list(CASES))
def test_init(case):
    hps, g, P = load_case(case)
    new_stats = {}
    with torch.no_grad():
        mel = O.vaenar_init(P, hps, t(g, "texts"), t(g, "m_len"), t(g, "t_len"), t(g, "init_epsilon"),
                            masks=train_masks(hps, g, "init"), new_stats=new_stats)
    close(mel, g["init_mel"], rtol=1e-3, atol=1e-4)
    for k in P:
        if ".actnorm." in k:
            close(P[k], g["init_actnorm/" + k], rtol=1e-3, atol=1e-4)

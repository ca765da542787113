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


@pytest.mark.parametrize("case", list(CASES))
def test_train_step_gradients_and_adam(case):
    """The oracle's train_step (loss of train.py:135, autograd, Keras Adam) against the golden vectors produced by the
    reference's OWN train_step closure (train.py:127-138, compiled unmodified) over the shim
    (tests/golden/make_golden_grads.py): losses, gradient norm + a random projection of every trainable tensor, the Adam
    update of every tensor, full gradients of a few small tensors."""
    import os
    import zlib
    from golden_util import GOLDEN_DIR
    hps, g, P = load_case(case)
    G = dict(np.load(os.path.join(GOLDEN_DIR, case.replace(".npz", "_train_step.npz")), allow_pickle=False))
    Pg = {k: v.clone().requires_grad_(O.is_trainable(k)) for k, v in P.items()}
    loss, l2, kl, ll = O.train_step_loss(Pg, hps, t(g, "texts"), t(g, "mels"), t(g, "t_len"), t(g, "m_len"), float(G["kl_weight"]),
                                         int(g["rf"]), t(g, "train_eps"), masks=train_masks(hps, g), new_stats={})
    loss.backward()
    close(loss.detach(), G["loss"])
    close(l2.detach(), G["mel_l2"])
    close(kl.detach(), G["kl"], rtol=1e-4)
    close(ll.detach(), G["length_l2"])
    names = [str(n) for n in G["names"]]
    assert sorted(names) == sorted(k for k in P if O.is_trainable(k)) and len(names) == 485
    worst = 0.0
    for i, k in enumerate(names):
        gr = Pg[k].grad if Pg[k].grad is not None else torch.zeros_like(Pg[k])
        d = torch.randn(gr.numel(), generator=torch.Generator().manual_seed(zlib.crc32(k.encode())), dtype=torch.float64)
        flat = gr.double().reshape(-1)
        ref_n, ref_p = float(G["grad_norm"][i]), float(G["grad_proj"][i])
        scale = max(ref_n, 1e-12)
        if ref_n < 1e-7:                      # mathematically zero gradients (conv bias straight into BatchNorm, unused
            assert float(flat.norm()) < 1e-6, k   # out_projection columns): fp32 noise only
            continue
        e = max(abs(float(flat.norm()) - ref_n), abs(float(flat @ d) - ref_p)) / scale
        worst = max(worst, e)
        assert e < 2e-3, (k, e, ref_n)
        # Keras Adam step (first step: m = (1-b1) g, v = (1-b2) g^2): compare the update itself
        new, _, _ = O.adam_update(Pg[k].detach(), gr, torch.zeros_like(gr), torch.zeros_like(gr), 1, lr=hps.Train.learning_rate)
        upd = (new - P[k]).double().reshape(-1)
        un, up = float(G["upd_norm"][i]), float(G["upd_proj"][i])
        assert abs(float(upd.norm()) - un) <= 2e-3 * un + 1e-9, (k, float(upd.norm()), un)
        assert abs(float(upd @ d) - up) <= 1e-2 * un + 1e-9, (k, float(upd @ d), up)
    for key in G:
        if key.startswith("grad/"):
            k = key[5:]
            close(Pg[k].grad, G[key], rtol=2e-3, atol=2e-3 * float(np.abs(G[key]).max()) + 1e-12)

"""Generate golden input/output vectors by executing the REFERENCE'S OWN Python sources
(/root/reference/models/models.py + modules/*.py, unmodified) over oracle/tf_shim.py.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Weights are NOT stored: they are regenerated from ``oracle.vaenar_oracle.init_params(hps, seed,
zero_init_std=0.02)`` + ``randomize_bn_stats`` (deterministic CPU RNG) and loaded into the reference
model by attribute path; a per-case checksum of the weights guards against RNG drift.  Each .npz
holds the inputs, the injected noise / dropout masks and the reference outputs.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle.tf_shim as shim  # noqa: E402
from oracle import vaenar_oracle as O  # noqa: E402
from oracle.hparams import LJHPS, DataBakerHPS  # noqa: E402

REF = "/root/reference"


def weights_checksum(P):
    h = hashlib.sha256()
    for k in sorted(P):
        h.update(k.encode())
        h.update(P[k].detach().cpu().numpy().astype(np.float32).tobytes())
    return h.hexdigest()


def load_reference(hps_name):
    shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from configs.hparams import LJHPS as RLJ, DataBakerHPS as RDB  # the reference's own hparams
    from models.models import VAENAR
    return VAENAR, (RLJ if hps_name == "ljspeech" else RDB)


def build_reference_model(hps_name, P, texts, mels, t_len, m_len):
    """Instantiate the reference VAENAR over the shim, run once to build variables, then overwrite
    every variable from ``P`` by attribute path."""
    VAENAR, RH = load_reference(hps_name)
    shim.reset(seed=7)
    model = VAENAR(RH)
    with torch.no_grad():
        model.init(text_inputs=texts, mel_lengths=m_len, text_lengths=t_len)
        model(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len,
              reduction_factor=2, training=False, reduce_loss=True)
        ref_vars = shim.extract_variables(model)
        assert set(ref_vars) == set(P), (sorted(set(ref_vars) ^ set(P))[:20])
        for k, v in ref_vars.items():
            assert tuple(v.shape) == tuple(P[k].shape), (k, v.shape, P[k].shape)
            v.copy_(P[k])
    return model, ref_vars


def np_(x):
    return x.detach().cpu().numpy().copy()


def make_case(hps, seed, B, T_t, T_m, rf, out_name):
    texts, mels, t_len, m_len = O.synthetic_batch(hps, B, T_t, T_m, rf=rf, seed=seed)
    P = O.init_params(hps, seed=seed, zero_init_std=0.02)
    O.randomize_bn_stats(P, seed=seed + 1)
    # keep a pristine copy: init() and training-mode BN mutate variables
    P0 = {k: v.clone() for k, v in P.items()}
    model, ref_vars = build_reference_model(hps.name, P, texts, mels, t_len, m_len)
    st = shim.state()
    out = dict(texts=np_(texts), mels=np_(mels), t_len=np_(t_len), m_len=np_(m_len), rf=np.int32(rf),
               seed=np.int64(seed), weights_sha256=np.array(weights_checksum(P0)))

    def restore():
        with torch.no_grad():
            for k, v in ref_vars.items():
                v.copy_(P0[k])

    with torch.no_grad():
        # ---- 1. VAENAR.call, training=False -------------------------------------------------
        shim.reset(seed=11)
        dec, l2, kl, ll, ali = model(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len,
                                     reduction_factor=rf, training=False, reduce_loss=True)
        assert len(st.normal_log) == 1
        out.update(eval_eps=np_(st.normal_log[0]), eval_mel=np_(dec), eval_l2=np_(l2), eval_kl=np_(kl),
                   eval_len=np_(ll))
        for k, v in ali.items():
            out["eval_ali_" + k] = np_(v)
        # unreduced losses too
        shim.reset(seed=11)
        _, l2u, klu, llu, _ = model(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len,
                                    reduction_factor=rf, training=False, reduce_loss=False)
        out.update(eval_l2_per=np_(l2u), eval_kl_per=np_(klu), eval_len_per=np_(llu))

        # ---- 2. VAENAR.inference (temperature 1.0) ------------------------------------------
        shim.reset(seed=12)
        mel, ali = model.inference(inputs=texts, mel_lengths=m_len, text_lengths=t_len, reduction_factor=rf)
        assert len(st.normal_log) == 1
        out.update(inf_epsilon=np_(st.normal_log[0]), inf_mel=np_(mel))
        for k, v in ali.items():
            out["inf_ali_" + k] = np_(v)

        # ---- 3. inference.py test_step sub-module sequence (inference.py:125-143) -----------
        shim.reset(seed=13)
        pos_step = model.mel_text_len_ratio / 2.0
        text_embd = model.text_encoder(texts, t_len, pos_step=pos_step, training=False)
        pred = model.length_predictor(text_embd.detach(), t_len, training=False)
        out.update(sub_text_embd=np_(text_embd), sub_pred_len=np_(pred))
        z, logp = model.prior.sample((m_len + 1) // 2, text_embd, t_len, training=False, temperature=1.0)
        out.update(sub_epsilon=np_(st.normal_log[0]), sub_z=np_(z), sub_logp=np_(logp))
        lp_back = model.prior.log_probability(z=z, condition_inputs=text_embd, z_lengths=(m_len + 1) // 2,
                                              condition_lengths=t_len, training=False)
        out.update(sub_logp_roundtrip=np_(lp_back))

        # ---- 4. VAENAR.call, training=True (dropout masks + BN batch stats recorded) --------
        shim.reset(seed=14)
        dec, l2, kl, ll, _ = model(inputs=texts, mel_targets=mels, mel_lengths=m_len, text_lengths=t_len,
                                   reduction_factor=rf, training=True, reduce_loss=True)
        out.update(train_eps=np_(st.normal_log[0]), train_mel=np_(dec), train_l2=np_(l2), train_kl=np_(kl),
                   train_len=np_(ll))
        for i, m in enumerate(st.dropout_log):
            out[f"train_dropout_{i:02d}"] = np_(m).astype(np.float32)
        out["train_n_dropout"] = np.int32(len(st.dropout_log))
        # moving statistics after the step (BN side effect)
        for k, v in ref_vars.items():
            if k.endswith("moving_mean") or k.endswith("moving_variance"):
                out["train_bnstat/" + k] = np_(v)
        restore()

        # ---- 5. VAENAR.init (data-dependent ActNorm init, rf = 5, training=True) ------------
        shim.reset(seed=15)
        mel5 = model.init(text_inputs=texts, mel_lengths=m_len, text_lengths=t_len)
        out.update(init_epsilon=np_(st.normal_log[0]), init_mel=np_(mel5))
        for i, m in enumerate(st.dropout_log):
            out[f"init_dropout_{i:02d}"] = np_(m).astype(np.float32)
        out["init_n_dropout"] = np.int32(len(st.dropout_log))
        for k, v in ref_vars.items():
            if ".actnorm." in k:
                out["init_actnorm/" + k] = np_(v)
        restore()

    path = os.path.join(HERE, out_name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("reference sources not present; goldens can only be regenerated in the build container")
    make_case(LJHPS, seed=123456, B=3, T_t=21, T_m=67, rf=2, out_name="lj_b3_t21_m67_rf2.npz")
    make_case(LJHPS, seed=4242, B=2, T_t=9, T_m=43, rf=3, out_name="lj_b2_t9_m43_rf3.npz")
    make_case(DataBakerHPS, seed=12, B=2, T_t=17, T_m=50, rf=2, out_name="db_b2_t17_m50_rf2.npz")

"""Golden vectors for the TRAINING STEP: executes the reference's OWN ``train_step`` closure -- the source text of
/root/reference/train.py:127-138, extracted with ``ast`` and compiled unmodified -- over oracle/tf_shim.py
(GradientTape -> torch autograd, tf.keras.optimizers.Adam -> the TF 2.2 ResourceApplyAdam formula) on the inputs, posterior
noise and dropout masks of the existing golden cases, and records

  * the four returned scalars (loss, mel_l2, kl, length_l2),
  * for every trainable variable the L2 norm of its gradient and its projection on a fixed pseudo-random direction
    (seeded by the CRC32 of the variable name), the same two numbers for the Adam update (after - before),
  * the full gradient of a few small variables.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden_grads.py
"""
import ast
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle.tf_shim as shim  # noqa: E402
from golden_util import CASES, load_case, t  # noqa: E402
from make_golden import REF, build_reference_model, load_reference  # noqa: E402

FULL = ("text_encoder.pos_weight", "posterior.pos_weight", "length_predictor.projection.kernel",
        "length_predictor.projection.bias", "prior.glow.0.actnorm.log_scale", "prior.glow.0.actnorm.bias",
        "prior.glow.5.affine_coupling.net.pos_weight", "posterior.mu_projection.bias", "decoder.residual_projection.bias")


def direction(name, numel):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    return torch.randn(numel, generator=g, dtype=torch.float64)


def reference_train_step(tf, model, hparams, optimizer):
    """Compile the reference's train_step closure (train.py:127-138) as it stands."""
    tree = ast.parse(open(os.path.join(REF, "train.py")).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "train_step")
    code = compile(ast.Module(body=[fn], type_ignores=[]), os.path.join(REF, "train.py"), "exec")
    ns = dict(tf=tf, model=model, hparams=hparams, optimizer=optimizer, print=lambda *a, **k: None)
    exec(code, ns)
    return ns["train_step"]


def make_case(name):
    hps, g, P = load_case(name)
    texts, mels, t_len, m_len = t(g, "texts"), t(g, "mels"), t(g, "t_len"), t(g, "m_len")
    rf = int(g["rf"])
    model, ref_vars = build_reference_model(hps.name, P, texts, mels, t_len, m_len)
    _, RH = load_reference(hps.name)
    tf = sys.modules["tensorflow"]
    optimizer = tf.keras.optimizers.Adam(RH.Train.learning_rate, beta_1=0.9, beta_2=0.999, epsilon=1e-07)   # train.py:116-117
    train_step = reference_train_step(tf, model, RH, optimizer)
    shim.enable_grad(model)
    before = {k: v.detach().clone() for k, v in ref_vars.items()}
    # the generator state of make_golden.py step 4: same posterior noise and dropout masks as the stored golden case
    shim.reset(seed=14)
    grads_box = {}
    real_gradient = shim.GradientTape.gradient

    def spy(self, target, sources):
        out = real_gradient(self, target, sources)
        grads_box["g"] = out
        return out
    shim.GradientTape.gradient = spy
    try:
        loss, l2, kl, ll = train_step(texts, mels, t_len, m_len, torch.tensor(float(RH.Train.kl_weight_init)), rf)
    finally:
        shim.GradientTape.gradient = real_gradient
    st = shim.state()
    assert np.allclose(st.normal_log[0].detach().numpy(), g["train_eps"]), "posterior noise differs from the golden case"
    assert abs(float(l2) - float(g["train_l2"])) <= 1e-5 * abs(float(g["train_l2"])), (float(l2), float(g["train_l2"]))
    assert abs(float(kl) - float(g["train_kl"])) <= 1e-5 * abs(float(g["train_kl"]))
    tv = model.trainable_variables
    names = [k for k in shim.extract_variables(model) if not (k.endswith("moving_mean") or k.endswith("moving_variance"))]
    assert len(tv) == len(names) == len(grads_box["g"]) == 485
    out = dict(loss=np.float64(float(loss)), mel_l2=np.float64(float(l2)), kl=np.float64(float(kl)), length_l2=np.float64(float(ll)),
               kl_weight=np.float64(float(RH.Train.kl_weight_init)), names=np.array(names))
    gn, gp, un, up = [], [], [], []
    for k, v, gr in zip(names, tv, grads_box["g"]):
        gr = torch.zeros_like(v) if gr is None else gr
        d = direction(k, v.numel())
        gflat = gr.detach().double().reshape(-1)
        uflat = (v.detach() - before[k]).double().reshape(-1)
        gn.append(float(gflat.norm())); gp.append(float(gflat @ d))
        un.append(float(uflat.norm())); up.append(float(uflat @ d))
        if k in FULL:
            out["grad/" + k] = gr.detach().numpy().copy()
    out.update(grad_norm=np.array(gn), grad_proj=np.array(gp), upd_norm=np.array(un), upd_proj=np.array(up))
    path = os.path.join(HERE, name.replace(".npz", "_train_step.npz"))
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", "loss", float(loss))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("reference sources not present; goldens can only be regenerated in the build container")
    for case in CASES:
        make_case(case)

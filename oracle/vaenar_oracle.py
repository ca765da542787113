"""TEST INFRASTRUCTURE — op-by-op CPU restatement (PyTorch, fp32 or fp64) of the VAENAR-TTS
non-autoregressive mel-synthesis hot path.  NOT part of the product; see oracle/__init__.py.

PARITY PIN STATUS: TensorFlow 2.2 cannot be installed in this image and the reference ships
no golden vectors, so the arithmetic of the *leaf* TF/Keras ops is restated from their
published semantics ("parity unpinned" at the leaf level).  The *wiring* (every reference
function below) is pinned by executing the reference's own Python sources over a minimal TF
API shim (oracle/tf_shim.py, tests/golden/make_golden.py) and comparing with this file.

All randomness is an explicit input (posterior ``eps``, prior ``epsilon``, dropout masks) and
all weights live in a flat ``name -> tensor`` dict with Keras layouts (Dense kernel [in,out],
Conv1D kernel [k,in,out]).  Every function cites the reference file:line it follows
(paths relative to /root/reference).
"""
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

LN_EPS = 1e-3          # Keras LayerNormalization default epsilon
BN_EPS = 1e-3          # Keras BatchNormalization default epsilon
BN_MOMENTUM = 0.99     # Keras BatchNormalization default momentum
MASK_FILL = float(-2.0 ** 32 + 1)   # modules/attention.py:240


# ----------------------------------------------------------------------------------------
# parameter construction (Keras default initialisers; SURVEY.md Appendix B naming)
# ----------------------------------------------------------------------------------------
def _glorot(gen, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(lim).float()


def _dense(P, gen, name, din, dout, bias=True, zero_kernel=False, zero_std=0.0):
    if zero_kernel:
        k = torch.zeros(din, dout)
        if zero_std > 0:
            k = torch.randn(din, dout, generator=gen, dtype=torch.float64).mul_(zero_std).float()
    else:
        k = _glorot(gen, (din, dout), din, dout)
    P[name + ".kernel"] = k
    if bias:
        P[name + ".bias"] = torch.zeros(dout)


def _ln(P, name, d):
    P[name + ".gamma"] = torch.ones(d)
    P[name + ".beta"] = torch.zeros(d)


def _conv_bn(P, gen, name, k, cin, cout):
    P[name + ".conv1d.kernel"] = _glorot(gen, (k, cin, cout), k * cin, k * cout)
    P[name + ".conv1d.bias"] = torch.zeros(cout)
    P[name + ".bn.gamma"] = torch.ones(cout)
    P[name + ".bn.beta"] = torch.zeros(cout)
    P[name + ".bn.moving_mean"] = torch.zeros(cout)
    P[name + ".bn.moving_variance"] = torch.ones(cout)


def _ffn(P, gen, name, din, hidden):
    _dense(P, gen, name + ".dense1", din, hidden)
    _dense(P, gen, name + ".dense2", hidden, din)
    _ln(P, name + ".layer_norm", din)


def _xblk(P, gen, name, d, mem_dim, ffn_hidden):
    """CrossAttentionBLK variables, modules/attention.py:418-434 (input_dim == attention_dim == d)."""
    for qkv, din in (("query", d), ("key", d), ("value", d)):
        _dense(P, gen, f"{name}.self_attention.{qkv}_layer", din, d, bias=False)
    _dense(P, gen, name + ".att_proj1", 2 * d, d)
    _ln(P, name + ".layer_norm1", d)
    _dense(P, gen, name + ".cross_attention.query_layer", d, d, bias=False)
    _dense(P, gen, name + ".cross_attention.key_layer", mem_dim, d, bias=False)
    _dense(P, gen, name + ".cross_attention.value_layer", mem_dim, d, bias=False)
    _dense(P, gen, name + ".att_proj2", 2 * d, d)
    _ln(P, name + ".layer_norm2", d)
    _ffn(P, gen, name + ".ffn", d, ffn_hidden)


NON_TRAINABLE_SUFFIXES = (".bn.moving_mean", ".bn.moving_variance")


def is_trainable(name: str) -> bool:
    return not name.endswith(NON_TRAINABLE_SUFFIXES)


def init_params(hps, seed: int = 0, zero_init_std: float = 0.0) -> Params:
    """Random-init weights with the Keras default distributions (models/models.py:16-65).

    ``zero_init_std > 0`` replaces the zero-initialised projection kernels
    (modules/posterior.py:108-113, modules/transform.py:12-17) by N(0, std) so that parity
    runs exercise the blocks feeding them.
    """
    gen = torch.Generator().manual_seed(seed)
    P: Params = {}
    E, D, C, O = hps.Encoder, hps.Decoder, hps.Common, hps.Common.output_dim
    # --- text encoder (modules/encoder.py:58-77)
    P["text_encoder.emb_layer.embeddings"] = (
        (torch.rand(E.vocab_size, E.embd_dim, generator=gen, dtype=torch.float64) * 0.1 - 0.05).float())
    P["text_encoder.pos_weight"] = torch.tensor(1.0)
    cin = E.embd_dim
    for i in range(E.n_conv):
        _conv_bn(P, gen, f"text_encoder.prenet.conv_stack.{i}", E.conv_kernel, cin, E.pre_hidden)
        cin = E.pre_hidden
    _dense(P, gen, "text_encoder.prenet.projection", E.pre_hidden, E.pre_hidden)
    for i in range(E.n_blk):
        n = f"text_encoder.self_attentions.{i}"
        for qkv in ("query", "key", "value"):
            _dense(P, gen, f"{n}.attention.{qkv}_layer", E.pre_hidden, E.attention_dim, bias=False)
        _dense(P, gen, n + ".att_proj", E.pre_hidden + E.attention_dim, E.pre_hidden)
        _ln(P, n + ".layer_norm", E.pre_hidden)
        _ffn(P, gen, n + ".ffn", E.pre_hidden, E.ffn_hidden)
    mem = E.pre_hidden
    # --- length predictor (modules/length_predictor.py:33)
    _dense(P, gen, "length_predictor.projection", mem, 1)
    # --- posterior (modules/posterior.py:90-113)
    Q = hps.Posterior
    P["posterior.pos_weight"] = torch.tensor(1.0)
    _dense(P, gen, "posterior.prenet.dense1", O, Q.pre_hidden)
    _dense(P, gen, "posterior.prenet.dense2", Q.pre_hidden, Q.pre_hidden)
    for i in range(Q.nblk):
        _xblk(P, gen, f"posterior.attentions.{i}", Q.attention_dim, mem, Q.ffn_hidden)
    _dense(P, gen, "posterior.mu_projection", Q.attention_dim, C.latent_dim, zero_kernel=True, zero_std=zero_init_std)
    _dense(P, gen, "posterior.logvar_projection", Q.attention_dim, C.latent_dim, zero_kernel=True, zero_std=zero_init_std)
    # --- prior (modules/prior.py:79-99, modules/flow.py:116-211, modules/transform.py:8-43)
    R = hps.Prior
    half = C.latent_dim // 2
    for i in range(R.n_blk):
        g = f"prior.glow.{i}"
        P[g + ".actnorm.log_scale"] = torch.randn(C.latent_dim, generator=gen, dtype=torch.float64).mul_(0.05).float()
        P[g + ".actnorm.bias"] = torch.zeros(C.latent_dim)
        w = torch.randn(C.latent_dim, C.latent_dim, generator=gen, dtype=torch.float64)
        P[g + ".linear.weight"] = torch.linalg.qr(w)[0].float()
        n = g + ".affine_coupling.net"
        P[n + ".pos_weight"] = torch.tensor(1.0)
        _dense(P, gen, n + ".pre_projection", half, R.attention_dim)
        for j in range(R.n_transformer_blk):
            _xblk(P, gen, f"{n}.attentions.{j}", R.attention_dim, mem, R.ffn_hidden)
        _dense(P, gen, n + ".log_scale_proj", R.attention_dim, half, zero_kernel=True, zero_std=zero_init_std)
        _dense(P, gen, n + ".shift_proj", R.attention_dim, half, zero_kernel=True, zero_std=zero_init_std)
    # --- decoder (modules/decoder.py:156-179)
    _dense(P, gen, "decoder.pre_projection", C.latent_dim, D.attention_dim)
    for i in range(D.nblk):
        _xblk(P, gen, f"decoder.attentions.{i}", D.attention_dim, mem, D.ffn_hidden)
    _dense(P, gen, "decoder.out_projection", D.attention_dim, O * C.max_reduction_factor)
    cin = O
    for i in range(D.post_n_conv):
        _conv_bn(P, gen, f"decoder.postnet.conv_stack.{i}", D.post_conv_kernel, cin, D.post_conv_filters)
        cin = D.post_conv_filters
    _dense(P, gen, "decoder.residual_projection", D.post_conv_filters, O)
    return P


def randomize_bn_stats(P: Params, seed: int = 1) -> None:
    """Give BN moving stats / affine non-trivial values so inference-mode BN is exercised."""
    gen = torch.Generator().manual_seed(seed)
    for k in list(P.keys()):
        if k.endswith(".bn.moving_mean"):
            P[k] = torch.randn(P[k].shape, generator=gen) * 0.1
        elif k.endswith(".bn.moving_variance"):
            P[k] = torch.rand(P[k].shape, generator=gen) * 0.5 + 0.5
        elif k.endswith(".bn.gamma"):
            P[k] = torch.rand(P[k].shape, generator=gen) * 0.4 + 0.8
        elif k.endswith(".bn.beta"):
            P[k] = torch.randn(P[k].shape, generator=gen) * 0.1


def cast_params(P: Params, dtype) -> Params:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in P.items()}


# ----------------------------------------------------------------------------------------
# leaf ops
# ----------------------------------------------------------------------------------------
_OPERAND_ROUND = None   # None = exact fp32/fp64.  See emulate_operand_dtype().


class emulate_operand_dtype:
    """Context manager: round the operands of every tensor-core contraction (Dense, Conv1D, QK^T, PV)
    to ``dtype`` (torch.bfloat16 / torch.float16) while accumulating in the working precision.  Used to
    predict the error budget of the mixed-precision CUDA path; the flow arithmetic stays exact."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _OPERAND_ROUND
        self._prev = _OPERAND_ROUND
        _OPERAND_ROUND = self.dtype
        return self

    def __exit__(self, *a):
        global _OPERAND_ROUND
        _OPERAND_ROUND = self._prev


def _r(x):
    return x if _OPERAND_ROUND is None else x.to(_OPERAND_ROUND).to(x.dtype)


def sequence_mask(lengths: Tensor, maxlen: int, dtype=torch.bool) -> Tensor:
    """tf.sequence_mask: mask[b, t] = t < lengths[b]."""
    return (torch.arange(maxlen)[None, :] < lengths[:, None].long()).to(dtype)


def dense(P, name, x, activation=None):
    y = _r(x) @ _r(P[name + ".kernel"])
    if (name + ".bias") in P:
        y = y + P[name + ".bias"]
    if activation == "relu":
        y = torch.relu(y)
    return y


def layer_norm(P, name, x):
    """Keras LayerNormalization over the last axis, epsilon 1e-3 (modules/attention.py:402,428,433;
    modules/utils.py:46)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + LN_EPS) * P[name + ".gamma"] + P[name + ".beta"]


def positional_encoding(length: int, dim: int, step=1.0, dtype=torch.float32) -> Tensor:
    """modules/utils.py:332-355: even d -> sin(t*step / 10000^(d/D)), odd d -> cos(t*step / 10000^((d-1)/D))."""
    pos = (torch.arange(length, dtype=dtype) * step)[:, None].expand(length, dim)
    d = torch.arange(dim, dtype=dtype)[None, :].expand(length, dim)
    even = (torch.arange(dim) % 2 == 0)[None, :]
    base = torch.tensor(10000.0, dtype=dtype)
    s = torch.sin(pos / torch.pow(base, d / dim))
    c = torch.cos(pos / torch.pow(base, (d - 1) / dim))
    return torch.where(even, s, c)


def conv1d_bn(P, name, x, activation, training, mask=None, bn_before_act=False, new_stats=None):
    """Conv1D wrapper: conv k 'same' -> act -> BatchNorm -> dropout (modules/utils.py:56-85).

    BN in training mode uses the batch mean / population variance over (batch, time) including
    padded frames, and updates the moving stats with momentum 0.99 (returned via ``new_stats``).
    """
    w = P[name + ".conv1d.kernel"]                      # [k, cin, cout]
    k = w.shape[0]
    pad_l = (k - 1) // 2
    xt = torch.nn.functional.pad(x.transpose(1, 2), (pad_l, k - 1 - pad_l))
    y = torch.nn.functional.conv1d(_r(xt), _r(w.permute(2, 1, 0)), P[name + ".conv1d.bias"]).transpose(1, 2)

    def act(v):
        if activation == "relu":
            return torch.relu(v)
        if activation == "tanh":
            return torch.tanh(v)
        return v

    def bn(v):
        if training:
            mean = v.mean(dim=(0, 1))
            var = ((v - mean) ** 2).mean(dim=(0, 1))
            if new_stats is not None:
                new_stats[name + ".bn.moving_mean"] = (
                    P[name + ".bn.moving_mean"] * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM))
                new_stats[name + ".bn.moving_variance"] = (
                    P[name + ".bn.moving_variance"] * BN_MOMENTUM + var.detach() * (1 - BN_MOMENTUM))
        else:
            mean, var = P[name + ".bn.moving_mean"], P[name + ".bn.moving_variance"]
        return (v - mean) * torch.rsqrt(var + BN_EPS) * P[name + ".bn.gamma"] + P[name + ".bn.beta"]

    y = act(bn(y)) if bn_before_act else bn(act(y))
    if mask is not None:
        y = y * mask
    return y


def mha(P, name, x, memory, heads, memory_lengths, query_lengths, causality, temperature=1.0):
    """MultiHeadScaledProductAttention.call (modules/attention.py:217-246)."""
    q = _r(x) @ _r(P[name + ".query_layer.kernel"])
    k = _r(memory) @ _r(P[name + ".key_layer.kernel"])
    v = _r(memory) @ _r(P[name + ".value_layer.kernel"])
    B, Tq, A = q.shape
    Tk = k.shape[1]
    hd = A // heads
    qh = q.reshape(B, Tq, heads, hd).transpose(1, 2)
    kh = k.reshape(B, Tk, heads, hd).transpose(1, 2)
    vh = v.reshape(B, Tk, heads, hd).transpose(1, 2)
    logits = _r(qh) @ _r(kh).transpose(-1, -2)
    logits = logits / math.sqrt(float(hd))
    logits = logits / temperature
    mmask = sequence_mask(memory_lengths, Tk)[:, None, :].expand(B, Tq, Tk)
    qmask = sequence_mask(query_lengths, Tq)[:, :, None].expand(B, Tq, Tk)
    mask = mmask & qmask
    if causality:
        mask = mask & torch.ones(Tq, Tk, dtype=torch.bool).tril()[None]
    mask = mask[:, None].expand(B, heads, Tq, Tk)
    logits = torch.where(mask, logits, torch.full_like(logits, MASK_FILL))
    ali = torch.softmax(logits, dim=3)
    ctx = (_r(ali) @ _r(vh)).transpose(1, 2).reshape(B, Tq, A)
    return ctx, ali


def ffn(P, name, x):
    """FFN.call (modules/utils.py:48-53): LN(dense2(relu(dense1 x)) + x)."""
    h = dense(P, name + ".dense1", x, "relu")
    return layer_norm(P, name + ".layer_norm", dense(P, name + ".dense2", h) + x)


def self_attention_blk(P, name, x, lengths, heads):
    """SelfAttentionBLK.call (modules/attention.py:405-415); encoder use: memory = inputs, no causality."""
    ctx, ali = mha(P, name + ".attention", x, x, heads, lengths, lengths, causality=False)
    proj = dense(P, name + ".att_proj", torch.cat([x, ctx], dim=-1))
    h = layer_norm(P, name + ".layer_norm", x + proj)
    return ffn(P, name + ".ffn", h), ali


def cross_attention_blk(P, name, x, memory, query_lengths, memory_lengths, heads):
    """CrossAttentionBLK.call (modules/attention.py:436-452)."""
    a1, _ = mha(P, name + ".self_attention", x, x, heads, query_lengths, query_lengths, causality=True)
    s = layer_norm(P, name + ".layer_norm1", dense(P, name + ".att_proj1", torch.cat([x, a1], dim=-1)) + x)
    a2, cross_ali = mha(P, name + ".cross_attention", s, memory, heads, memory_lengths, query_lengths,
                        causality=False)
    c = layer_norm(P, name + ".layer_norm2", dense(P, name + ".att_proj2", torch.cat([s, a2], dim=-1)) + s)
    return ffn(P, name + ".ffn", c), cross_ali


# ----------------------------------------------------------------------------------------
# modules
# ----------------------------------------------------------------------------------------
def text_encoder(P, hps, texts, text_lengths, pos_step, training=False, masks=None, new_stats=None):
    """TransformerEncoder.call (modules/encoder.py:79-93)."""
    masks = masks or {}
    E = hps.Encoder
    x = P["text_encoder.emb_layer.embeddings"][texts.long()]
    for i in range(E.n_conv):                                              # ConvPreNet, utils.py:33-38
        x = conv1d_bn(P, f"text_encoder.prenet.conv_stack.{i}", x, "relu", training,
                      masks.get(f"enc.prenet.{i}"), E.bn_before_act, new_stats)
    x = dense(P, "text_encoder.prenet.projection", x)
    pe = positional_encoding(x.shape[1], x.shape[2], pos_step, x.dtype)
    x = x + P["text_encoder.pos_weight"] * pe
    if "enc.pos" in masks:
        x = x * masks["enc.pos"]
    for i in range(E.n_blk):
        x, _ = self_attention_blk(P, f"text_encoder.self_attentions.{i}", x, text_lengths, E.attention_heads)
    return x


def length_predictor(P, text_embd, text_lengths):
    """DenseLengthPredictor.call (modules/length_predictor.py:35-42), identity activation."""
    proj = dense(P, "length_predictor.projection", text_embd)
    mask = sequence_mask(text_lengths, text_embd.shape[1], text_embd.dtype)[:, :, None]
    return (torch.exp(proj) * mask).sum(dim=(1, 2))


def posterior(P, hps, reduced_mels, text_embd, text_lengths, z_lengths, masks=None):
    """TransformerPosterior.call (modules/posterior.py:115-130).  Returns (mu_projection out,
    logvar_projection out) in the reference's own return order."""
    masks = masks or {}
    Q = hps.Posterior
    h = dense(P, "posterior.prenet.dense1", reduced_mels, "relu")           # PreNet, utils.py:13-18
    if "post.prenet.1" in masks:
        h = h * masks["post.prenet.1"]
    h = dense(P, "posterior.prenet.dense2", h, "relu")
    if "post.prenet.2" in masks:
        h = h * masks["post.prenet.2"]
    pe = positional_encoding(h.shape[1], h.shape[2], 1.0, h.dtype)
    h = h + P["posterior.pos_weight"] * pe
    if "post.pos" in masks:
        h = h * masks["post.pos"]
    for i in range(Q.nblk):
        h, _ = cross_attention_blk(P, f"posterior.attentions.{i}", h, text_embd, z_lengths, text_lengths,
                                   Q.attention_heads)
    return dense(P, "posterior.mu_projection", h), dense(P, "posterior.logvar_projection", h)


def reparameterize(mu, logvar, eps):
    """BasePosterior.reparameterize (modules/posterior.py:20-39); eps: [B, n, T, D]."""
    std = torch.exp(0.5 * logvar)
    return eps * std[:, None] + mu[:, None]


def posterior_log_probability(mu, logvar, eps, seq_lengths):
    """BasePosterior.log_probability with eps given (modules/posterior.py:41-72)."""
    dim = mu.shape[2]
    t = -0.5 * (dim * math.log(2 * math.pi) + (logvar[:, None] + eps ** 2.0).sum(dim=3))
    mask = sequence_mask(seq_lengths, mu.shape[1], mu.dtype)[:, None, :]
    return (mask * t).sum(dim=2)


def actnorm_forward(P, name, z, lengths):
    """ActNormFlow._forward (modules/flow.py:166-175)."""
    s = P[name + ".log_scale"]
    return z * torch.exp(s) + P[name + ".bias"], lengths.to(z.dtype) * s.sum()


def actnorm_backward(P, name, z, lengths, epsilon=1e-8):
    """ActNormFlow._backward (modules/flow.py:177-187)."""
    s = P[name + ".log_scale"]
    return (z - P[name + ".bias"]) / (torch.exp(s) + epsilon), lengths.to(z.dtype) * (-s.sum())


def actnorm_init(P, name, z, lengths, init_scale=1.0, epsilon=1e-8):
    """ActNormFlow.init (modules/flow.py:189-196): data-dependent, over ALL B*T positions."""
    flat = z.reshape(-1, z.shape[-1])
    mean = flat.mean(dim=0)
    std = torch.sqrt(((flat - mean) ** 2).mean(dim=0))          # tf.math.reduce_std = population std
    P[name + ".log_scale"] = torch.log(init_scale / (std + epsilon)).detach()
    P[name + ".bias"] = (-mean / (std + epsilon)).detach()
    return actnorm_forward(P, name, z, lengths)


def invlinear_forward(P, name, z, lengths):
    """InvertibleLinearFlow._forward (modules/flow.py:123-135); logdet via float64 slogdet."""
    w = P[name + ".weight"]
    logdet = torch.linalg.slogdet(w.double())[1].to(z.dtype)
    return z @ w, lengths.to(z.dtype) * logdet


def invlinear_backward(P, name, z, lengths):
    """InvertibleLinearFlow._backward (modules/flow.py:137-150): fp32 inverse for the matmul,
    float64 inverse + slogdet for the log-determinant."""
    w = P[name + ".weight"]
    logdet = torch.linalg.slogdet(torch.linalg.inv(w.double()))[1].to(z.dtype)
    return z @ torch.linalg.inv(w), lengths.to(z.dtype) * logdet


def transformer_transform(P, hps, name, z, text_embd, text_lengths, z_lengths):
    """TransformerTransform.call (modules/transform.py:45-59)."""
    R = hps.Prior
    h = dense(P, name + ".pre_projection", z)
    pe = positional_encoding(h.shape[1], h.shape[2], 1.0, h.dtype)
    h = h + P[name + ".pos_weight"] * pe
    for j in range(R.n_transformer_blk):
        h, _ = cross_attention_blk(P, f"{name}.attentions.{j}", h, text_embd, z_lengths, text_lengths,
                                   R.attention_heads)
    return dense(P, name + ".log_scale_proj", h), dense(P, name + ".shift_proj", h)


def coupling(P, hps, name, x, text_embd, z_lengths, text_lengths, upper: bool, backward: bool):
    """TransformerCoupling._forward / _backward (modules/flow.py:223-257)."""
    half = x.shape[-1] // 2
    lower_pt, upper_pt = x[..., :half], x[..., half:]
    z, zp = (lower_pt, upper_pt) if upper else (upper_pt, lower_pt)
    log_scale, shift = transformer_transform(P, hps, name + ".net", z, text_embd, text_lengths, z_lengths)
    scale = torch.sigmoid(log_scale + 2.0)
    mask = sequence_mask(z_lengths, x.shape[1], x.dtype)[:, :, None]
    if backward:
        zp = (zp - shift) / (scale + 1e-12)
        logdet = -(torch.log(scale) * mask).sum(dim=(1, 2))
    else:
        zp = scale * zp + shift
        logdet = (torch.log(scale) * mask).sum(dim=(1, 2))
    out = torch.cat([z, zp], dim=-1) if upper else torch.cat([zp, z], dim=-1)
    return out, logdet


def _initial_logprob(epsilon, lengths):
    """BasePrior._initial_sample log-density part (modules/prior.py:37-41)."""
    lp = -0.5 * (math.log(2.0 * math.pi) + epsilon ** 2)
    mask = sequence_mask(lengths, epsilon.shape[1], epsilon.dtype)[:, :, None]
    return (mask * lp).sum(dim=(1, 2))


def prior_sample(P, hps, epsilon, z_lengths, text_embd, text_lengths, init=False):
    """TransformerPrior.sample (modules/prior.py:154-169) / .init (:171-186).

    ``epsilon`` is the already-drawn N(0, temperature) noise [B, max(z_lengths), latent]."""
    logprobs = _initial_logprob(epsilon, z_lengths)
    z = epsilon
    for i in range(hps.Prior.n_blk):
        g = f"prior.glow.{i}"
        if init:
            z, ld = actnorm_init(P, g + ".actnorm", z, z_lengths)
        else:
            z, ld = actnorm_forward(P, g + ".actnorm", z, z_lengths)
        logprobs = logprobs - ld
        z, ld = invlinear_forward(P, g + ".linear", z, z_lengths)
        logprobs = logprobs - ld
        z, ld = coupling(P, hps, g + ".affine_coupling", z, text_embd, z_lengths, text_lengths,
                         upper=(i % 2 == 0), backward=False)
        logprobs = logprobs - ld
    return z, logprobs


def prior_log_probability(P, hps, z, text_embd, z_lengths, text_lengths):
    """TransformerPrior.log_probability (modules/prior.py:119-152)."""
    eps = z
    accum = torch.zeros(z.shape[0], dtype=z.dtype)
    for i in reversed(range(hps.Prior.n_blk)):
        g = f"prior.glow.{i}"
        eps, ld = coupling(P, hps, g + ".affine_coupling", eps, text_embd, z_lengths, text_lengths,
                           upper=(i % 2 == 0), backward=True)
        accum = accum + ld
        eps, ld = invlinear_backward(P, g + ".linear", eps, z_lengths)
        accum = accum + ld
        eps, ld = actnorm_backward(P, g + ".actnorm", eps, z_lengths)
        accum = accum + ld
    return _initial_logprob(eps, z_lengths) + accum


def decoder(P, hps, z, text_embd, z_lengths, text_lengths, reduction_factor, training=False, masks=None,
            new_stats=None):
    """TransformerDecoder.call (modules/decoder.py:181-199)."""
    masks = masks or {}
    D = hps.Decoder
    O = hps.Common.output_dim
    B, T = z.shape[0], z.shape[1]
    h = dense(P, "decoder.pre_projection", z)
    alignments = {}
    for i in range(D.nblk):
        h, ali = cross_attention_blk(P, f"decoder.attentions.{i}", h, text_embd, z_lengths, text_lengths,
                                     D.attention_heads)
        alignments[f"decoder-attention-{i}"] = ali
    initial = dense(P, "decoder.out_projection", h)[:, :, : reduction_factor * O]
    initial = initial.reshape(B, T * reduction_factor, O)
    r = initial
    for i in range(D.post_n_conv):                                         # PostNet, utils.py:98-115
        act = "tanh" if i < D.post_n_conv - 1 else None
        r = conv1d_bn(P, f"decoder.postnet.conv_stack.{i}", r, act, training, masks.get(f"dec.postnet.{i}"),
                      False, new_stats)
    r = dense(P, "decoder.residual_projection", r)
    return initial, r + initial, alignments


# ----------------------------------------------------------------------------------------
# model API (models/models.py)
# ----------------------------------------------------------------------------------------
def compute_l2_loss(rec, tgt, lengths, reduce):
    """VAENAR._compute_l2_loss (models/models.py:67-86), n_sample = 1."""
    mask = sequence_mask(lengths, rec.shape[1], rec.dtype)
    l2 = (((rec - tgt) ** 2).mean(dim=-1) * mask).sum(dim=-1) / lengths.to(rec.dtype)
    return l2.mean() if reduce else l2


def length_l2_loss(pred, target_lengths, reduce):
    """VAENAR._length_l2_loss (models/models.py:96-103)."""
    d = (torch.log(pred) - torch.log(target_lengths.to(pred.dtype))) ** 2
    return d.mean() if reduce else d


def vaenar_call(P, hps, texts, mels, mel_lengths, text_lengths, reduction_factor, eps, training=False,
                reduce_loss=True, masks=None, new_stats=None):
    """VAENAR.call (models/models.py:105-197) with n_sample = 1.

    eps: posterior noise [B, 1, T_z, latent].  Returns (decoded_outs, l2, kl, length_loss,
    dec_alignments) plus an ``aux`` dict of intermediates for parity tests.
    """
    rf = int(reduction_factor)
    mel_max_len = mels.shape[1]
    reduced_mels = mels[:, ::rf, :]
    reduced_lens = (mel_lengths + rf - 1) // rf
    pos_step = hps.Common.mel_text_len_ratio / float(rf)
    text_embd = text_encoder(P, hps, texts, text_lengths, pos_step, training, masks, new_stats)
    pred_len = length_predictor(P, text_embd.detach(), text_lengths)
    length_loss = length_l2_loss(pred_len, mel_lengths, reduce_loss)
    # models.py:136 unpacks (mu, logvar, None) as (logvar, mu, _): the *name swap* is reproduced.
    logvar, mu = posterior(P, hps, reduced_mels, text_embd, text_lengths, reduced_lens,
                           masks if training else None)
    samples = reparameterize(mu, logvar, eps)
    post_logp = posterior_log_probability(mu, logvar, eps, reduced_lens)          # [B, 1]
    z = samples.reshape(samples.shape[0], samples.shape[2], samples.shape[3])
    initial, outs, ali = decoder(P, hps, z, text_embd, reduced_lens, text_lengths, rf, training, masks, new_stats)
    initial = initial[:, :mel_max_len]
    outs = outs[:, :mel_max_len]
    l2 = compute_l2_loss(outs, mels, mel_lengths, reduce_loss) + compute_l2_loss(initial, mels, mel_lengths,
                                                                                 reduce_loss)
    prior_logp = prior_log_probability(P, hps, z, text_embd, reduced_lens, text_lengths)[:, None]
    kl = (post_logp - prior_logp).mean(dim=1)
    if reduce_loss:
        kl = kl.mean()
    aux = dict(text_embd=text_embd, mu=mu, logvar=logvar, z=z, post_logp=post_logp, prior_logp=prior_logp,
               initial=initial, pred_len=pred_len)
    return outs, l2, kl, length_loss, ali, aux


def vaenar_inference(P, hps, texts, mel_lengths, text_lengths, reduction_factor, epsilon):
    """VAENAR.inference (models/models.py:199-210). epsilon: [B, max(reduced lens), latent] ~ N(0,1)."""
    rf = int(reduction_factor)
    reduced_lens = (mel_lengths + rf - 1) // rf
    pos_step = hps.Common.mel_text_len_ratio / float(rf)
    text_embd = text_encoder(P, hps, texts, text_lengths, pos_step, training=False)
    z, logp = prior_sample(P, hps, epsilon, reduced_lens, text_embd, text_lengths)
    initial, mel, ali = decoder(P, hps, z, text_embd, reduced_lens, text_lengths, rf, training=False)
    return mel, ali, dict(text_embd=text_embd, z=z, logp=logp, initial=initial)


def vaenar_init(P, hps, texts, mel_lengths, text_lengths, epsilon, masks=None, new_stats=None):
    """VAENAR.init (models/models.py:212-226): data-dependent ActNorm init at rf = max_reduction_factor,
    training=True.  Mutates the actnorm entries of ``P``."""
    rf = hps.Common.max_reduction_factor
    reduced_lens = (mel_lengths + rf - 1) // rf
    pos_step = hps.Common.mel_text_len_ratio / float(rf)
    text_embd = text_encoder(P, hps, texts, text_lengths, pos_step, True, masks, new_stats)
    z, logp = prior_sample(P, hps, epsilon, reduced_lens, text_embd, text_lengths, init=True)
    _, mel, _ = decoder(P, hps, z, text_embd, reduced_lens, text_lengths, rf, True, masks, new_stats)
    return mel


def train_step_loss(P, hps, texts, mels, text_lengths, mel_lengths, kl_weight, reduction_factor, eps,
                    masks=None, new_stats=None):
    """Loss of the train_step closure (train.py:127-135)."""
    _, l2, kl, length_l2, _, _ = vaenar_call(P, hps, texts, mels, mel_lengths, text_lengths, reduction_factor,
                                             eps, training=True, reduce_loss=True, masks=masks,
                                             new_stats=new_stats)
    loss = l2 + kl_weight * torch.clamp(kl, min=0.0) + hps.Train.length_weight * length_l2
    return loss, l2, kl, length_l2


def adam_update(param, grad, m, v, step, lr=1.25e-4, b1=0.9, b2=0.999, eps=1e-7):
    """Keras Adam (TF 2.2 ``ResourceApplyAdam`` form; train.py:116-117): step counts from 1."""
    lr_t = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    return param - lr_t * m / (torch.sqrt(v) + eps), m, v


# ----------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------
def synthetic_batch(hps, B, T_text, T_mel, rf=2, seed=None):
    """LJSpeech-shaped synthetic batch: row 0 has the maximum lengths, the rest are ragged."""
    seed = hps.Train.random_seed if seed is None else seed
    rng = np.random.default_rng(seed)
    V = hps.Encoder.vocab_size
    ratio = hps.Common.mel_text_len_ratio
    t_len = rng.integers((T_text + 1) // 2, T_text + 1, size=B)
    t_len[0] = T_text
    m_len = np.minimum(T_mel, np.maximum(rf, np.round(ratio * t_len).astype(np.int64)))
    m_len[0] = T_mel
    texts = np.zeros((B, T_text), dtype=np.int32)
    mels = np.zeros((B, T_mel, hps.num_mels), dtype=np.float32)
    for b in range(B):
        n = int(t_len[b])
        body = rng.integers(3, V, size=max(n - 2, 0))
        seq = np.concatenate([[1], body, [2]])[:n]
        texts[b, :n] = seq
        mels[b, : m_len[b]] = rng.random((int(m_len[b]), hps.num_mels), dtype=np.float32)
    return (torch.from_numpy(texts), torch.from_numpy(mels), torch.from_numpy(t_len.astype(np.int32)),
            torch.from_numpy(m_len.astype(np.int32)))

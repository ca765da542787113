"""Host-side mirror of the reference's ``models.models.VAENAR`` (models/models.py:9-226) on top of the
C ABI of libvaenar_sm100.so.  Same constructor, method and attribute names as the Keras model so that
train.py:114-179 / inference.py:38-72,121-143 can use it unchanged:

    model = VAENAR(hps)
    model(inputs=, mel_targets=, mel_lengths=, text_lengths=, reduction_factor=, training=, reduce_loss=)
    model.inference(inputs=, mel_lengths=, text_lengths=, reduction_factor=)
    model.text_encoder(t, t_l, pos_step=, training=) ; model.length_predictor(x, t_l, training=)
    model.prior.sample(lens, text_embd, t_l, training=, temperature=) ; model.prior.log_probability(...)
    model.decoder(z, text_embd, z_lens, t_l, training=, reduction_factor=) ; model.mel_text_len_ratio

PyTorch tensors are only the device-memory containers; all arithmetic happens in the sm_100a kernels.
There is no CPU / eager fallback: a compute call without a CUDA device (or without the built library) raises.
"""
import ctypes
import math
from collections import OrderedDict

import torch

from . import _lib
from ._lib import HParamsStruct, TrainOpts, VaenarError, check


def _tf(hps_section):
    """Accept both the reference nesting (hps.Encoder.Transformer.x) and a flattened one (hps.Encoder.x)."""
    return getattr(hps_section, "Transformer", hps_section)


def hparams_struct(hps) -> HParamsStruct:
    E, D, Q, R, C = _tf(hps.Encoder), _tf(hps.Decoder), _tf(hps.Posterior), _tf(hps.Prior), hps.Common
    s = HParamsStruct()
    s.vocab_size, s.embd_dim, s.enc_n_conv, s.enc_hidden = E.vocab_size, E.embd_dim, E.n_conv, E.pre_hidden
    s.enc_conv_kernel, s.enc_n_blk, s.enc_att_dim, s.enc_heads, s.enc_ffn = (
        E.conv_kernel, E.n_blk, E.attention_dim, E.attention_heads, E.ffn_hidden)
    s.dec_nblk, s.dec_att_dim, s.dec_heads, s.dec_ffn = D.nblk, D.attention_dim, D.attention_heads, D.ffn_hidden
    s.post_n_conv, s.post_filters, s.post_kernel = D.post_n_conv, D.post_conv_filters, D.post_conv_kernel
    s.posterior_pre_hidden, s.posterior_nblk, s.posterior_att_dim = Q.pre_hidden, Q.nblk, Q.attention_dim
    s.posterior_heads, s.posterior_ffn = Q.attention_heads, Q.ffn_hidden
    s.prior_n_blk, s.prior_n_tblk, s.prior_att_dim = R.n_blk, R.n_transformer_blk, R.attention_dim
    s.prior_heads, s.prior_ffn = R.attention_heads, R.ffn_hidden
    s.latent_dim, s.out_dim = C.latent_dim, C.output_dim
    s.max_reduction_factor, s.final_reduction_factor = C.max_reduction_factor, C.final_reduction_factor
    s.mel_text_len_ratio = float(C.mel_text_len_ratio)
    s.enc_pre_drop_rate = float(getattr(E, "pre_drop_rate", 0.1))
    s.enc_pos_drop_rate = float(getattr(E, "pos_drop_rate", 0.1))
    s.posterior_pre_drop_rate = float(getattr(Q, "pre_drop_rate", 0.5))
    s.posterior_pos_drop_rate = float(getattr(Q, "pos_drop_rate", 0.2))
    s.post_drop_rate = float(getattr(D, "post_drop_rate", 0.2))
    for name, temp in (("Encoder", getattr(E, "attention_temperature", 1.0)),
                       ("Decoder", getattr(D, "attention_temperature", 1.0)),
                       ("Posterior", getattr(Q, "temperature", 1.0)), ("Prior", getattr(R, "temperature", 1.0))):
        if float(temp) != 1.0:
            raise VaenarError(f"{name} attention temperature {temp} != 1.0 is not supported by the fused kernel")
    if getattr(R, "inverse", False):
        raise VaenarError("Prior.inverse=True is not supported (reference configs use inverse=False)")
    for name, sec in (("Encoder", E), ("Decoder", D)):
        if getattr(sec, "bn_before_act", False):
            raise VaenarError(f"{name}.bn_before_act=True is not supported (modules/utils.py:58 default False in all reference configs)")
    return s


class _Sub:
    """Callable sub-module facade (Keras ``Layer.__call__`` look-alike)."""

    def __init__(self, fn, **extra):
        self._fn = fn
        for k, v in extra.items():
            setattr(self, k, v)

    def __call__(self, *a, **kw):
        return self._fn(*a, **kw)


class VAENAR:
    def __init__(self, hps, name="VAENAR", device=None, seed=None, noise_seed_offset=0, **kwargs):
        """``seed`` initialises the parameters (shared by all data-parallel replicas); ``noise_seed_offset`` (e.g. the rank) is
        mixed into the posterior-noise / dropout streams so that replicas draw different noise for their different data."""
        self.hps = hps
        self.name = name
        self.n_sample = hps.Train.num_samples
        if self.n_sample != 1:
            raise VaenarError("num_samples != 1 is not implemented (reference configs use 1)")
        self.mel_text_len_ratio = hps.Common.mel_text_len_ratio
        self.max_reduction_factor = hps.Common.max_reduction_factor
        self._lib = _lib.load()
        self._hp = hparams_struct(hps)
        h = ctypes.c_void_p()
        check(self._lib.vaenar_create(ctypes.byref(self._hp), ctypes.byref(h)))
        self._h = h
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # ---- parameter manifest: one flat fp32 buffer, named views into it
        L = self._lib
        self._manifest = []
        for i in range(L.vaenar_num_params(h)):
            nd = L.vaenar_param_ndim(h, i)
            shape = tuple(int(L.vaenar_param_dim(h, i, d)) for d in range(nd))
            self._manifest.append((L.vaenar_param_name(h, i).decode(), shape, int(L.vaenar_param_offset(h, i)),
                                   bool(L.vaenar_param_trainable(h, i))))
        self._flat = torch.zeros(int(L.vaenar_param_floats(h)), dtype=torch.float32, device=self.device)
        self._views = OrderedDict()
        for n, shape, off, _ in self._manifest:
            numel = int(math.prod(shape))
            self._views[n] = self._flat[off:off + numel].view(shape)
        self._packed = None
        self._dirty = True
        self._ws = None
        self._noise_calls = 0
        self._seed = int(hps.Train.random_seed if seed is None else seed)
        self._noise_seed = self._seed + 7919 * int(noise_seed_offset)
        self.reset_parameters(self._seed)
        # ---- sub-modules with the reference's attribute names (inference.py:59-72,129-143)
        self.text_encoder = _Sub(self._text_encoder)
        self.length_predictor = _Sub(self._length_predictor)
        self.decoder = _Sub(self._decoder)
        self.posterior = _Sub(self._posterior, reparameterize=self._posterior_reparameterize,
                              log_probability=self._posterior_log_probability, sample_fused=self._posterior_sample)
        self.prior = _Sub(self._prior_sample, sample=self._prior_sample, log_probability=self._prior_log_probability,
                          init=self._prior_init)

    def __del__(self):
        try:
            for base in getattr(self, "_peer_bases", []):
                self._lib.vaenar_ipc_close(ctypes.c_void_p(base))
            self._peer_bases = []
        except Exception:
            pass
        try:
            if getattr(self, "_h", None):
                self._lib.vaenar_destroy(self._h)
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def reset_parameters(self, seed):
        """Keras default initialisers (glorot_uniform Dense/Conv1D, U(-.05,.05) Embedding, zero-init
        projections of modules/posterior.py:108-113 & modules/transform.py:12-17, QR-orthogonal
        InvertibleLinear of modules/flow.py:120, N(0,.05) ActNorm log-scale of :160-162)."""
        g = torch.Generator().manual_seed(int(seed))
        sd = {}
        for n, shape, _, _ in self._manifest:
            if n.endswith((".bias", ".beta", ".moving_mean", ".actnorm.bias")):
                v = torch.zeros(shape)
            elif n.endswith((".gamma", ".moving_variance", "pos_weight")):
                v = torch.ones(shape)
            elif n.endswith("embeddings"):
                v = torch.rand(shape, generator=g) * 0.1 - 0.05
            elif n.endswith("actnorm.log_scale"):
                v = torch.randn(shape, generator=g) * 0.05
            elif n.endswith("linear.weight"):
                v = torch.linalg.qr(torch.randn(shape, generator=g, dtype=torch.float64))[0].float()
            elif any(n.endswith(s + ".kernel") for s in ("mu_projection", "logvar_projection", "log_scale_proj",
                                                          "shift_proj")):
                v = torch.zeros(shape)
            elif n.endswith(".kernel"):
                if len(shape) == 3:
                    fan_in, fan_out = shape[0] * shape[1], shape[0] * shape[2]
                else:
                    fan_in, fan_out = shape
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                v = (torch.rand(shape, generator=g) * 2 - 1) * lim
            else:
                raise VaenarError(f"no initialiser for {n}")
            sd[n] = v
        self.load_state_dict(sd)

    def state_dict(self):
        return OrderedDict((n, v.detach().clone()) for n, v in self._views.items())

    def load_state_dict(self, sd, strict=True):
        missing = [n for n in self._views if n not in sd]
        extra = [n for n in sd if n not in self._views]
        if strict and (missing or extra):
            raise VaenarError(f"state dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        with torch.no_grad():
            for n, v in self._views.items():
                if n in sd:
                    src = torch.as_tensor(sd[n], dtype=torch.float32)
                    if src.numel() != v.numel():
                        raise VaenarError(f"{n}: expected {tuple(v.shape)}, got {tuple(src.shape)}")
                    v.copy_(src.reshape(v.shape))
        self._dirty = True

    def load_tf_checkpoint(self, prefix, strict=True, restore_optimizer=True):
        """Restore the model weights from a TF2 object-graph checkpoint written by the reference
        (``tf.train.Checkpoint(step=, optimizer=, model=)``, train.py:246-248 / inference.py:39-41): ``prefix`` =
        ``.../ckpt-N``; when the checkpoint carries them, also the Adam moments and the optimizer's iteration count
        (a resumed run continues the bias correction instead of restarting at t = 1).  TensorFlow is not needed
        (vaenar_tts_b200/tf_checkpoint.py)."""
        from . import tf_checkpoint
        sd = tf_checkpoint.load_tf_checkpoint(prefix)
        if any(v.ndim == 0 for v in sd.values()):
            sd = {k: (v.reshape(1) if v.ndim == 0 else v) for k, v in sd.items()}     # TF scalars have shape []
        self.load_state_dict(sd, strict=strict)
        if restore_optimizer:
            m, v, it = tf_checkpoint.load_tf_optimizer_state(prefix)
            if m and v:
                self._ensure_adam_state()
                for n, shape, off, tr in self._manifest:
                    if tr and n in m and n in v:
                        numel = int(math.prod(shape))
                        self._adam_m[off:off + numel].copy_(torch.as_tensor(m[n], dtype=torch.float32).reshape(-1))
                        self._adam_v[off:off + numel].copy_(torch.as_tensor(v[n], dtype=torch.float32).reshape(-1))
            if it is not None:
                self._opt_step = int(it)

    def save_tf_checkpoint(self, prefix, step=0):
        """Write what ``tf.train.Checkpoint(step=, optimizer=, model=)`` persists (train.py:246): the weights under the
        reference's checkpoint keys, the Adam moments as optimizer slots and the optimizer iteration count."""
        from . import tf_checkpoint
        m = v = None
        if hasattr(self, "_adam_m"):
            m, v = {}, {}
            for n, shape, off, tr in self._manifest:
                if tr:
                    numel = int(math.prod(shape))
                    m[n] = self._adam_m[off:off + numel].view(shape)
                    v[n] = self._adam_v[off:off + numel].view(shape)
        elif getattr(self, "_peer", None) is not None:
            raise VaenarError("save_tf_checkpoint with the sharded peer optimizer: gather the moment shards first "
                              "(each rank holds 1/world of Adam's m and v)")
        tf_checkpoint.save_tf_checkpoint(prefix, self.state_dict(), step=step, adam_m=m, adam_v=v,
                                         opt_step=getattr(self, "_opt_step", None))

    def _ensure_adam_state(self):
        if not hasattr(self, "_adam_m"):
            self._adam_m = torch.zeros_like(self._flat)
            self._adam_v = torch.zeros_like(self._flat)
            host = torch.zeros(self._flat.numel(), dtype=torch.uint8)
            check(self._lib.vaenar_trainable_mask(self._h, ctypes.c_void_p(host.data_ptr())))
            self._trainable_mask = host.to(self.device)

    @property
    def trainable_variables(self):
        return [self._views[n] for n, _, _, tr in self._manifest if tr]

    @property
    def variables(self):
        return list(self._views.values())

    def parameter_names(self):
        return [n for n, _, _, _ in self._manifest]

    def mark_weights_changed(self):
        self._dirty = True

    def apply_gradients(self, flat_grads, step, lr=None, beta_1=0.9, beta_2=0.999, epsilon=1e-7, grad_scale=1.0,
                        skip_flag=None):
        """optimizer.apply_gradients of train.py:137 (Keras Adam, lr 1.25e-4) over the flat gradient buffer, which has
        the layout of the flat parameter buffer.  One fused kernel; Adam moments are created on first use."""
        self._require_cuda()
        g = self._f32(flat_grads)
        if g.numel() != self._flat.numel():
            raise VaenarError("gradient buffer must have the flat parameter layout")
        self._ensure_adam_state()
        lr = float(self.hps.Train.learning_rate if lr is None else lr)
        check(self._lib.vaenar_adam_step(self._p(self._flat), self._p(g), self._p(self._adam_m), self._p(self._adam_v),
                                         self._p(self._trainable_mask), self._flat.numel(), int(step), lr, float(beta_1),
                                         float(beta_2), float(epsilon), float(grad_scale), self._p(skip_flag),
                                         self._stream()))
        self._dirty = True

    def flat_parameters(self):
        return self._flat

    # ------------------------------------------------------------------ plumbing
    def _require_cuda(self):
        if self.device.type != "cuda":
            raise VaenarError("VAENAR compute needs a CUDA device (sm_100a); there is no CPU fallback")

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)

    def _i32(self, t):
        return torch.as_tensor(t).to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()

    def _f32(self, t):
        return torch.as_tensor(t).to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()

    def _prepare(self, B, Tt, Tz, rf, defer_flow_join=False):
        self._require_cuda()
        if self._packed is None:
            self._packed = torch.empty(int(self._lib.vaenar_packed_bytes(self._h)), dtype=torch.uint8,
                                       device=self.device)
        if self._dirty:
            pack = self._lib.vaenar_pack_weights_async if defer_flow_join else self._lib.vaenar_pack_weights
            check(pack(self._h, self._p(self._flat), self._p(self._packed), self._stream()))
            self._dirty = False
        need = int(self._lib.vaenar_workspace_bytes(self._h, B, Tt, Tz, rf))
        if need < 0:
            check(-1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)

    def _noise(self, shape, stddev=1.0):
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        self._noise_calls += 1
        check(self._lib.vaenar_randn(self._p(out), out.numel(), self._noise_seed, self._noise_calls, float(stddev),
                                     self._stream()))
        return out

    def _train_opts(self, dropout_masks, update_bn_stats=True):
        """vaenar_train_opts_t: injected dropout keep-masks (parity tests) or on-device generation from the seed."""
        o = TrainOpts()
        self._noise_calls += 1
        o.seed = (self._noise_seed << 20) + self._noise_calls
        o.update_bn_stats = 1 if update_bn_stats else 0
        keep = None
        if dropout_masks is not None:
            keep = [self._f32(m) for m in dropout_masks]
            arr = (ctypes.c_void_p * len(keep))(*[m.data_ptr() for m in keep])
            o.n_masks = len(keep)
            o.masks = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
            keep.append(arr)
        return o, keep

    @staticmethod
    def _no_training(training, what):
        if training:
            raise NotImplementedError(
                f"{what}(training=True) as a stand-alone sub-module call is not exposed: training-mode arithmetic (dropout, "
                "batch-statistics BatchNorm) runs inside VAENAR.__call__(training=True) / train_step / init; refusing to "
                "silently run inference-mode arithmetic here")

    @staticmethod
    def _max_len(lengths):
        return int(torch.as_tensor(lengths).max().item())

    # ------------------------------------------------------------------ sub-modules
    def _text_encoder(self, inputs, input_lengths=None, pos_step=1.0, training=None):
        """TransformerEncoder.call (modules/encoder.py:79-93)."""
        self._no_training(training, "text_encoder")
        texts = self._i32(inputs)
        B, Tt = texts.shape
        t_len = self._i32(input_lengths) if input_lengths is not None else torch.full((B,), Tt, dtype=torch.int32,
                                                                                     device=self.device)
        self._prepare(B, Tt, 1, 1)
        out = torch.empty(B, Tt, self._hp.enc_hidden, dtype=torch.float32, device=self.device)
        check(self._lib.vaenar_text_encoder_fwd(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                                self._ws.numel(), self._p(texts), self._p(t_len), B, Tt,
                                                float(pos_step), self._p(out), self._stream()))
        return out

    def _length_predictor(self, inputs, input_lengths, training=None):
        """DenseLengthPredictor.call (modules/length_predictor.py:35-42)."""
        self._require_cuda()
        x = self._f32(inputs)
        B, Tt, _ = x.shape
        t_len = self._i32(input_lengths)
        out = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self._lib.vaenar_length_predictor_fwd(self._h, self._p(self._flat), self._p(x), self._p(t_len), B, Tt,
                                                    self._p(out), self._stream()))
        return out

    def _prior_sample(self, targets_lengths, condition_inputs, condition_lengths=None, training=None,
                      temperature=1.0, epsilon=None):
        """TransformerPrior.sample (modules/prior.py:154-169).  ``epsilon`` (optional) injects the N(0,1)
        draw of _initial_sample (prior.py:35) for parity tests; it is scaled by ``temperature``."""
        emb = self._f32(condition_inputs)
        B, Tt, _ = emb.shape
        Tz = self._max_len(targets_lengths)
        z_len = self._i32(targets_lengths)
        t_len = self._i32(condition_lengths)
        self._prepare(B, Tt, Tz, 1)
        L = self._hp.latent_dim
        if epsilon is None:
            z = self._noise((B, Tz, L), temperature)
        else:
            z = (self._f32(epsilon) * float(temperature)).contiguous().clone()
            if tuple(z.shape) != (B, Tz, L):
                raise VaenarError(f"epsilon shape {tuple(z.shape)} != {(B, Tz, L)}")
        logp = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self._lib.vaenar_prior_sample(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                            self._ws.numel(), self._p(emb), self._p(t_len), self._p(z_len), B, Tt, Tz,
                                            self._p(z), self._p(logp), self._stream()))
        return z, logp

    def _prior_log_probability(self, z, condition_inputs, z_lengths=None, condition_lengths=None, training=None):
        """TransformerPrior.log_probability (modules/prior.py:119-152)."""
        z = self._f32(z)
        emb = self._f32(condition_inputs)
        B, Tz, _ = z.shape
        Tt = emb.shape[1]
        z_len = self._i32(z_lengths)
        t_len = self._i32(condition_lengths)
        self._prepare(B, Tt, Tz, 1)
        logp = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self._lib.vaenar_prior_log_probability(self._h, self._p(self._flat), self._p(self._packed),
                                                     self._p(self._ws), self._ws.numel(), self._p(z), self._p(emb),
                                                     self._p(t_len), self._p(z_len), B, Tt, Tz, self._p(logp),
                                                     self._stream()))
        return logp

    def _prior_init(self, targets_lengths, condition_inputs, condition_lengths=None, training=None, epsilon=None):
        """TransformerPrior.init (modules/prior.py:171-186): data-dependent ActNorm initialisation of every flow step from
        an N(0,1) sample pushed through the flow; writes actnorm.log_scale / bias, returns (z, logp) like ``sample``.
        The reference only reaches it through VAENAR.init; as a stand-alone call it needs the text encoding."""
        emb = self._f32(condition_inputs)
        B, Tt, _ = emb.shape
        Tz = self._max_len(targets_lengths)
        z_len = self._i32(targets_lengths)
        t_len = self._i32(condition_lengths)
        self._prepare(B, Tt, Tz, 1)
        L = self._hp.latent_dim
        z = self._noise((B, Tz, L)) if epsilon is None else self._f32(epsilon).contiguous().clone()
        check(self._lib.vaenar_prior_init(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                          self._ws.numel(), self._p(emb), self._p(t_len), self._p(z_len), B, Tt, Tz,
                                          self._p(z), self._stream()))
        self._dirty = True                  # ActNorm parameters changed: operands are re-packed on the next call
        return z, self._prior_log_probability(z, emb, z_len, t_len)

    def _posterior(self, inputs, src_enc, src_lengths=None, target_lengths=None, training=None):
        """TransformerPosterior.call (modules/posterior.py:115-130) -> (mu, logvar, None): the outputs of mu_projection
        and logvar_projection, NOT swapped (VAENAR.call swaps them when unpacking, models/models.py:136).  ``inputs`` are
        the reduced mels [B, T_z, 80]."""
        self._no_training(training, "posterior")
        rm = self._f32(inputs)
        emb = self._f32(src_enc)
        B, Tz, _ = rm.shape
        Tt = emb.shape[1]
        z_len = self._i32(target_lengths) if target_lengths is not None else torch.full((B,), Tz, dtype=torch.int32, device=self.device)
        t_len = self._i32(src_lengths) if src_lengths is not None else torch.full((B,), Tt, dtype=torch.int32, device=self.device)
        L = self._hp.latent_dim
        self._prepare(B, Tt, Tz, 1)
        mu = torch.empty(B, Tz, L, dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mu)
        check(self._lib.vaenar_posterior_params(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                                self._ws.numel(), self._p(rm), self._p(emb), self._p(t_len), self._p(z_len),
                                                B, Tt, Tz, self._p(mu), self._p(logvar), self._stream()))
        return mu, logvar, None

    def _posterior_reparameterize(self, mu, logvar, nsamples=1, random=True, eps=None):
        """BasePosterior.reparameterize (modules/posterior.py:20-39) -> (samples, eps) [B, nsamples, T, dim].  Facade helper on
        device tensors (elementwise); VAENAR.__call__ / train_step use the fused GEMM epilogue instead."""
        mu, logvar = self._f32(mu), self._f32(logvar)
        B, T, D = mu.shape
        n = int(nsamples)
        if eps is None:
            eps = self._noise((B, n, T, D)) if random else torch.zeros(B, n, T, D, dtype=torch.float32, device=self.device)
        else:
            eps = self._f32(eps).reshape(B, n, T, D)
        return eps * torch.exp(0.5 * logvar)[:, None] + mu[:, None], eps

    def _posterior_log_probability(self, mu, logvar, z=None, eps=None, seq_lengths=None, epsilon=1e-8):
        """BasePosterior.log_probability (modules/posterior.py:41-72) -> [B, nsamples]."""
        mu, logvar = self._f32(mu), self._f32(logvar)
        B, T, D = mu.shape
        if eps is not None:
            ns = self._f32(eps)
        else:
            ns = (self._f32(z) - mu[:, None]) / (torch.exp(0.5 * logvar)[:, None] + float(epsilon))
        tl = -0.5 * (D * math.log(2 * math.pi) + (logvar[:, None] + ns ** 2).sum(-1))
        if seq_lengths is not None:
            mask = (torch.arange(T, device=self.device)[None, :] < self._i32(seq_lengths)[:, None]).float()
            tl = tl * mask[:, None]
        return tl.sum(-1)

    def _posterior_sample(self, reduced_mels, src_enc, src_lengths=None, target_lengths=None, eps=None):
        """The fused path VAENAR.call uses: posterior + reparameterize + log_probability (one GEMM epilogue), with the
        (logvar, mu) swap of models/models.py:136 applied -> (z [B,T_z,latent], logq [B])."""
        rm = self._f32(reduced_mels)
        emb = self._f32(src_enc)
        B, Tz, _ = rm.shape
        Tt = emb.shape[1]
        z_len = self._i32(target_lengths)
        t_len = self._i32(src_lengths)
        L = self._hp.latent_dim
        e = self._noise((B, Tz, L)) if eps is None else self._f32(eps).reshape(B, Tz, L).contiguous()
        self._prepare(B, Tt, Tz, 1)
        z = torch.empty(B, Tz, L, dtype=torch.float32, device=self.device)
        logq = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self._lib.vaenar_posterior_fwd(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                             self._ws.numel(), self._p(rm), self._p(emb), self._p(t_len),
                                             self._p(z_len), self._p(e), B, Tt, Tz, Tz, 1, self._p(z), self._p(logq),
                                             self._stream()))
        return z, logq

    def _decoder(self, inputs, text_embd, z_lengths=None, text_lengths=None, reduction_factor=2, training=None,
                 return_alignments=True):
        """TransformerDecoder.call (modules/decoder.py:181-199) -> (initial_outs, outputs, alignments)."""
        self._no_training(training, "decoder")
        z = self._f32(inputs)
        emb = self._f32(text_embd)
        B, Tz, _ = z.shape
        Tt = emb.shape[1]
        rf = int(reduction_factor)
        z_len = self._i32(z_lengths)
        t_len = self._i32(text_lengths)
        self._prepare(B, Tt, Tz, rf)
        O, H, nb = self._hp.out_dim, self._hp.dec_heads, self._hp.dec_nblk
        initial = torch.empty(B, Tz * rf, O, dtype=torch.float32, device=self.device)
        mel = torch.empty_like(initial)
        ali = torch.empty(nb, B, H, Tz, Tt, dtype=torch.float32, device=self.device) if return_alignments else None
        check(self._lib.vaenar_decoder_fwd(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                           self._ws.numel(), self._p(z), self._p(emb), self._p(z_len), self._p(t_len),
                                           B, Tt, Tz, rf, self._p(initial), self._p(mel), self._p(ali),
                                           self._stream()))
        return initial, mel, self._ali_dict(ali)

    def cross_attention_blocks(self, module, x, memory, query_lengths, memory_lengths, return_alignments=False):
        """The CrossAttentionBLK stack (modules/attention.py:418-452) of one sub-module on caller-supplied activations:
        ``module`` = "decoder" (modules/decoder.py:170-174), "posterior" (modules/posterior.py:100-106) or ("prior", s)
        for the coupling net of flow step s (modules/transform.py:37-43).  Block-level parity hook of the C ABI."""
        if module == "decoder":
            mid, nb = 0, self._hp.dec_nblk
        elif module == "posterior":
            mid, nb = 1, self._hp.posterior_nblk
        else:
            mid, nb = 2 + int(module[1]), self._hp.prior_n_tblk
        x = self._f32(x).clone()
        mem = self._f32(memory)
        B, T, _ = x.shape
        Tt = mem.shape[1]
        q_len, t_len = self._i32(query_lengths), self._i32(memory_lengths)
        self._prepare(B, Tt, T, 1)
        ali = torch.empty(nb, B, 4, T, Tt, dtype=torch.float32, device=self.device) if return_alignments else None
        check(self._lib.vaenar_xblk_stack_fwd(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                              self._ws.numel(), mid, self._p(x), self._p(mem), self._p(q_len),
                                              self._p(t_len), B, T, Tt, self._p(ali), self._stream()))
        return x, ali

    def _ali_dict(self, ali):
        if ali is None:
            return {}
        return {f"decoder-attention-{i}": ali[i] for i in range(ali.shape[0])}

    # ------------------------------------------------------------------ model API (models/models.py)
    def inference(self, inputs, mel_lengths, text_lengths=None, reduction_factor=2, epsilon=None,
                  return_alignments=True, check_finite=False):
        """VAENAR.inference (models/models.py:199-210) -> (predicted_mel, dec_alignments).  ``check_finite`` (one host
        sync): raise instead of returning non-finite mels -- the fp16 tensor-core operands saturate at 65504, which a
        pathological checkpoint can exceed where the fp32 reference would not."""
        rf = int(reduction_factor)
        texts = self._i32(inputs)
        B, Tt = texts.shape
        m_len = torch.as_tensor(mel_lengths)
        z_len_host = (m_len.to(torch.int64) + rf - 1) // rf
        Tz = self._max_len(z_len_host)
        z_len = self._i32(z_len_host)
        t_len = self._i32(text_lengths)
        self._prepare(B, Tt, Tz, rf)
        L, O, H, nb, E = (self._hp.latent_dim, self._hp.out_dim, self._hp.dec_heads, self._hp.dec_nblk,
                          self._hp.enc_hidden)
        z = self._noise((B, Tz, L)) if epsilon is None else self._f32(epsilon).contiguous().clone()
        if tuple(z.shape) != (B, Tz, L):
            raise VaenarError(f"epsilon shape {tuple(z.shape)} != {(B, Tz, L)}")
        emb = torch.empty(B, Tt, E, dtype=torch.float32, device=self.device)
        mel = torch.empty(B, Tz * rf, O, dtype=torch.float32, device=self.device)
        ali = torch.empty(nb, B, H, Tz, Tt, dtype=torch.float32, device=self.device) if return_alignments else None
        logp = torch.empty(B, dtype=torch.float32, device=self.device)
        check(self._lib.vaenar_inference(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                         self._ws.numel(), self._p(texts), self._p(t_len), self._p(z_len), B, Tt, Tz,
                                         rf, self._p(z), self._p(emb), self._p(mel), self._p(ali), self._p(logp),
                                         self._stream()))
        self._last = dict(z=z, text_embd=emb, logp=logp)
        if check_finite and not bool(torch.isfinite(mel).all()):
            raise VaenarError("inference produced non-finite mel values: an activation exceeded the fp16 operand range "
                              "(|x| > 65504) of the tensor-core path")
        return mel, self._ali_dict(ali)

    def test_step(self, t, t_l, temperature=0.0, return_alignments=True):
        """The ``test_step`` closure of inference.py:125-143 (length-predictor-driven synthesis): encoder -> predicted
        lengths (truncated to int32, +80 frames of safety margin) -> prior.sample(temperature; the reference default 0.0
        makes the synthesis deterministic) -> decoder at the final reduction factor.
        Returns (mel [B, T_z*rf, 80], predicted_mel_lengths + 80 [B] int32, alignments dict)."""
        rf = int(self.hps.Common.final_reduction_factor)
        texts = self._i32(t)
        t_len = self._i32(t_l)
        emb = self._text_encoder(texts, t_len, pos_step=self.mel_text_len_ratio / float(rf), training=False)
        pred = self._length_predictor(emb, t_len, training=False)
        pred_m_l = pred.to(torch.int32)                                     # tf.cast(float -> int32) truncates
        reduced = (pred_m_l + 80 + rf - 1) // rf
        z, _ = self._prior_sample(reduced, emb, t_len, training=False, temperature=float(temperature))
        _, mel, ali = self._decoder(z, emb, reduced, t_len, reduction_factor=rf, training=False,
                                    return_alignments=return_alignments)
        return mel, pred_m_l + 80, ali

    synthesize = test_step

    def call(self, inputs, mel_targets, mel_lengths, text_lengths=None, reduction_factor=2, training=None,
             reduce_loss=None, eps=None, return_alignments=True, dropout_masks=None, update_bn_stats=True):
        """VAENAR.call (models/models.py:105-197) -> (decoded_outs, l2_loss, kl_divergence, length_loss,
        dec_alignments).  ``eps`` (optional) injects the posterior noise [B,1,T_z,latent] for parity tests."""
        rf = int(reduction_factor)
        texts = self._i32(inputs)
        mels = self._f32(mel_targets)
        B, Tt = texts.shape
        Tm = mels.shape[1]
        Tz = (Tm + rf - 1) // rf
        m_len = self._i32(mel_lengths)
        t_len = self._i32(text_lengths)
        z_len = (m_len + (rf - 1)) // rf
        self._prepare(B, Tt, Tz, rf)
        L, O, H, nb = self._hp.latent_dim, self._hp.out_dim, self._hp.dec_heads, self._hp.dec_nblk
        e = self._noise((B, Tz, L)) if eps is None else self._f32(eps).reshape(B, Tz, L).contiguous()
        mel = torch.empty(B, Tm, O, dtype=torch.float32, device=self.device)
        l2 = torch.empty(B, dtype=torch.float32, device=self.device)
        kl = torch.empty_like(l2)
        ll = torch.empty_like(l2)
        ali = torch.empty(nb, B, H, Tz, Tt, dtype=torch.float32, device=self.device) if return_alignments else None
        if training:
            # forward with training=True semantics (BN batch statistics + moving-average update, dropout);
            # gradients: train_step_grads / train_step.
            opts, keep = self._train_opts(dropout_masks, update_bn_stats)
            check(self._lib.vaenar_elbo_fwd_train(
                self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws), self._ws.numel(), self._p(texts),
                self._p(mels), self._p(m_len), self._p(t_len), self._p(z_len), self._p(e), B, Tt, Tm, Tz, rf,
                ctypes.byref(opts), self._p(mel), self._p(l2), self._p(kl), self._p(ll), self._p(ali), self._stream()))
            if update_bn_stats:
                self._dirty = True          # folded inference BatchNorm constants are stale now
        else:
            check(self._lib.vaenar_elbo_fwd(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                            self._ws.numel(), self._p(texts), self._p(mels), self._p(m_len),
                                            self._p(t_len), self._p(z_len), self._p(e), B, Tt, Tm, Tz, rf, self._p(mel),
                                            self._p(l2), self._p(kl), self._p(ll), self._p(ali), self._stream()))
        if reduce_loss:
            l2, kl, ll = l2.mean(), kl.mean(), ll.mean()
        return mel, l2, kl, ll, self._ali_dict(ali)

    __call__ = call

    # ------------------------------------------------------------------ training (train.py:120-138)
    DEFAULT_LOSS_SCALE = 65536.0

    def train_step_grads(self, texts, mels, t_lengths, m_lengths, kl_weight, reduction_factor, eps=None,
                         dropout_masks=None, update_bn_stats=True, loss_scale=None, return_mel=False):
        """Forward (training=True) + hand-written backward of the train_step closure (train.py:127-136).
        Returns (losses, grads[, mel]): ``losses`` = device tensor [total, mel_l2, kl, length_l2]; ``grads`` = flat
        gradient buffer in the layout of ``flat_parameters()``, multiplied by ``loss_scale``."""
        rf = int(reduction_factor)
        texts = self._i32(texts)
        mels = self._f32(mels)
        B, Tt = texts.shape
        Tm = mels.shape[1]
        Tz = (Tm + rf - 1) // rf
        m_len = self._i32(m_lengths)
        t_len = self._i32(t_lengths)
        z_len = (m_len + (rf - 1)) // rf
        self._prepare(B, Tt, Tz, rf, defer_flow_join=True)      # joined inside vaenar_train_step_grads, before the prior
        need = int(self._lib.vaenar_train_workspace_bytes(self._h, B, Tt, Tz, rf))
        if need < 0:
            check(-1)
        if getattr(self, "_train_ws", None) is None or self._train_ws.numel() < need:
            self._train_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if getattr(self, "_grads", None) is None:
            self._grads = torch.zeros_like(self._flat)
            self._losses = torch.zeros(4, dtype=torch.float32, device=self.device)
        L, O = self._hp.latent_dim, self._hp.out_dim
        e = self._noise((B, Tz, L)) if eps is None else self._f32(eps).reshape(B, Tz, L).contiguous()
        mel = torch.empty(B, Tm, O, dtype=torch.float32, device=self.device) if return_mel else None
        S = float(self.DEFAULT_LOSS_SCALE if loss_scale is None else loss_scale)
        opts, keep = self._train_opts(dropout_masks, update_bn_stats)
        check(self._lib.vaenar_train_step_grads(
            self._h, self._p(self._flat), self._p(self._packed), self._p(self._train_ws), self._train_ws.numel(),
            self._p(texts), self._p(mels), self._p(m_len), self._p(t_len), self._p(z_len), self._p(e), B, Tt, Tm, Tz, rf,
            ctypes.byref(opts), float(kl_weight), float(self.hps.Train.length_weight), S, self._p(self._grads),
            self._p(self._losses), self._p(mel), self._stream()))
        if update_bn_stats:
            self._dirty = True
        self._last_loss_scale = S
        return (self._losses, self._grads, mel) if return_mel else (self._losses, self._grads)

    def broadcast_parameters(self, src=0, group=None):
        """Make every replica start from rank ``src``'s parameters (after ``init``: the data-dependent ActNorm
        initialisation ran on that rank's batch) -- one broadcast of the flat parameter buffer."""
        import torch.distributed as dist
        dist.broadcast(self._flat, src, group=group)
        self._dirty = True

    def enable_peer_optimizer(self, group=None):
        """Data-parallel training over NVLink peer memory: share the flat parameter / gradient buffers of all ranks of
        ``group`` through CUDA IPC so that train_step can run the gradient exchange and Adam as ONE kernel
        (vaenar_adam_step_sharded) instead of NCCL all-reduce + Adam.  One process per GPU on one node, world <= 8."""
        import torch.distributed as dist
        self._require_cuda()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > 8:
            raise VaenarError("peer optimizer needs 2..8 ranks on one node")
        if getattr(self, "_grads", None) is None:
            self._grads = torch.zeros_like(self._flat)
            self._losses = torch.zeros(4, dtype=torch.float32, device=self.device)
        def export(t):
            # cudaIpcMemHandle_t of the cudaMalloc segment that holds t + byte offset of t inside it
            buf = ctypes.create_string_buffer(64)
            off = ctypes.c_int64()
            with torch.cuda.device(self.device):
                check(self._lib.vaenar_ipc_export(self._p(t), ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(off)))
            return bytes(buf.raw), int(off.value)
        mine = [export(self._flat), export(self._grads)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        self._peer_bases = []                        # IPC mappings opened on OUR device (closed in __del__)
        pp, pg = [], []
        with torch.cuda.device(self.device):
            for r in range(world):
                if r == rank:
                    pp.append(self._flat.data_ptr())
                    pg.append(self._grads.data_ptr())
                    continue
                ptrs = []
                opened = {}                                  # both buffers may live in the same cudaMalloc segment:
                for handle, off in gathered[r]:              # a handle can be opened only once per process
                    if handle not in opened:
                        base = ctypes.c_void_p()
                        buf = ctypes.create_string_buffer(handle, 64)
                        check(self._lib.vaenar_ipc_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(base)))
                        self._peer_bases.append(base.value)
                        opened[handle] = base.value
                    ptrs.append(opened[handle] + off)
                pp.append(ptrs[0])
                pg.append(ptrs[1])
        self._peer_params = (ctypes.c_void_p * world)(*pp)
        self._peer_grads = (ctypes.c_void_p * world)(*pg)
        S = int(self._lib.vaenar_adam_shard_floats(self._flat.numel(), world))
        self._shard_m = torch.zeros(S, dtype=torch.float32, device=self.device)
        self._shard_v = torch.zeros(S, dtype=torch.float32, device=self.device)
        host = torch.zeros(self._flat.numel(), dtype=torch.uint8)
        check(self._lib.vaenar_trainable_mask(self._h, ctypes.c_void_p(host.data_ptr())))
        self._trainable_mask = host.to(self.device)
        self._peer = (rank, world, group)
        self._barrier_flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        dist.barrier(group=group)

    def _peer_adam(self, step, grad_scale, lr=None, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        import torch.distributed as dist
        rank, world, group = self._peer
        lr = float(self.hps.Train.learning_rate if lr is None else lr)
        # stream-ordered barrier: every rank's gradients are final -- the all-reduced value is the job-wide count of
        # non-finite gradient entries, so every replica takes the same skip decision
        dist.all_reduce(self._ovf, group=group)
        check(self._lib.vaenar_adam_step_sharded(
            ctypes.cast(self._peer_params, ctypes.c_void_p), ctypes.cast(self._peer_grads, ctypes.c_void_p),
            self._p(self._shard_m), self._p(self._shard_v), self._p(self._trainable_mask), self._flat.numel(), rank, world,
            int(step), lr, float(beta_1), float(beta_2), float(epsilon), float(grad_scale), self._p(self._ovf),
            self._stream()))
        dist.all_reduce(self._barrier_flag, group=group)          # every peer has written its shard into our parameters
        self._dirty = True

    # ---- dynamic loss scaling (no counterpart in the fp32 reference: the fp16 gradient operands can overflow)
    MAX_LOSS_SCALE = 65536.0
    SCALE_GROWTH_INTERVAL = 1000

    @property
    def loss_scale(self):
        return float(getattr(self, "_loss_scale", self.DEFAULT_LOSS_SCALE))

    @property
    def skipped_steps(self):
        self._poll_overflow(block=True)
        return int(getattr(self, "_skipped", 0))

    def _poll_overflow(self, block=False):
        """Consume the overflow counts of finished steps (pinned host copies, no device sync unless ``block``): a step with
        non-finite gradients was skipped on the device; halve the scale.  After SCALE_GROWTH_INTERVAL clean steps double it."""
        q = getattr(self, "_ovf_queue", None)
        if not q:
            return
        while q and (block or q[0][0].query()):
            ev, host = q.pop(0)
            ev.synchronize()
            if float(host[0]) != 0.0:
                self._skipped = getattr(self, "_skipped", 0) + 1
                self._loss_scale = max(self.loss_scale / 2.0, 1.0)
                self._good_steps = 0
            else:
                self._good_steps = getattr(self, "_good_steps", 0) + 1
                if self._good_steps >= self.SCALE_GROWTH_INTERVAL and self.loss_scale < self.MAX_LOSS_SCALE:
                    self._loss_scale = min(self.loss_scale * 2.0, self.MAX_LOSS_SCALE)
                    self._good_steps = 0

    def exchange_and_apply(self, group=None):
        """Second half of train_step (train.py:136-137) on the gradients of the last train_step_grads: overflow count,
        gradient exchange (data-parallel), Keras Adam.  A step whose gradients are not finite anywhere in the job is
        skipped on every replica (parameters and moments untouched)."""
        import torch.distributed as dist
        grads = self._grads
        if getattr(self, "_ovf", None) is None:
            self._ovf = torch.zeros(1, dtype=torch.float32, device=self.device)
        if getattr(self, "_ovf_queue", None) is None:
            self._ovf_queue = []
        self._ovf.zero_()
        self._opt_step = getattr(self, "_opt_step", 0) + 1
        S = self._last_loss_scale
        if getattr(self, "_peer", None) is not None:
            # gradient exchange + Adam fused over NVLink peer memory (reduce-scatter -> Adam -> all-gather in one kernel)
            check(self._lib.vaenar_grad_nonfinite(self._p(grads), grads.numel(), self._p(self._ovf), self._stream()))
            self._peer_adam(self._opt_step, 1.0 / (S * self._peer[1]))
        else:
            world = 1
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                world = dist.get_world_size(group)
                dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)      # the single collective of the training path
            # after the all-reduce: a non-finite entry of any replica is non-finite in the sum on every replica
            check(self._lib.vaenar_grad_nonfinite(self._p(grads), grads.numel(), self._p(self._ovf), self._stream()))
            self.apply_gradients(grads, self._opt_step, grad_scale=1.0 / (S * world), skip_flag=self._ovf)
        host = torch.empty(1, dtype=torch.float32).pin_memory()
        host.copy_(self._ovf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._ovf_queue.append((ev, host))
        if len(self._ovf_queue) > 64:
            self._poll_overflow(block=True)

    def train_step(self, texts, mels, t_lengths, m_lengths, kl_weight, reduction_factor, eps=None, dropout_masks=None,
                   loss_scale=None, group=None):
        """The train_step closure of train.py:120-138: loss, gradients, (data-parallel: ONE exchange of the flat gradient
        buffer), Keras Adam.  Returns (loss, mel_l2, kl, length_l2) as device tensor views.  The activation gradients are
        carried multiplied by a loss scale (default: dynamic, starting at 2^16); a step whose gradients overflow is skipped
        and the scale halved -- check ``skipped_steps`` / ``loss_scale``."""
        self._poll_overflow()
        S = self.loss_scale if loss_scale is None else float(loss_scale)
        losses, grads = self.train_step_grads(texts, mels, t_lengths, m_lengths, kl_weight, reduction_factor, eps=eps,
                                              dropout_masks=dropout_masks, loss_scale=S)
        self.exchange_and_apply(group=group)
        return losses[0], losses[1], losses[2], losses[3]

    def init(self, text_inputs, mel_lengths, text_lengths=None, epsilon=None, dropout_masks=None):
        """VAENAR.init (models/models.py:212-226): data-dependent ActNorm initialisation at
        rf = max_reduction_factor with training=True (dropout + BN batch statistics); returns the predicted mel.
        ``epsilon`` / ``dropout_masks`` optionally inject the random draws (parity tests)."""
        rf = self.max_reduction_factor
        texts = self._i32(text_inputs)
        B, Tt = texts.shape
        z_len_host = (torch.as_tensor(mel_lengths).to(torch.int64) + rf - 1) // rf
        Tz = self._max_len(z_len_host)
        z_len = self._i32(z_len_host)
        t_len = self._i32(text_lengths)
        self._prepare(B, Tt, Tz, rf)
        L, O = self._hp.latent_dim, self._hp.out_dim
        z = self._noise((B, Tz, L)) if epsilon is None else self._f32(epsilon).contiguous().clone()
        if tuple(z.shape) != (B, Tz, L):
            raise VaenarError(f"epsilon shape {tuple(z.shape)} != {(B, Tz, L)}")
        mel = torch.empty(B, Tz * rf, O, dtype=torch.float32, device=self.device)
        opts, keep = self._train_opts(dropout_masks, True)
        check(self._lib.vaenar_init(self._h, self._p(self._flat), self._p(self._packed), self._p(self._ws),
                                    self._ws.numel(), self._p(texts), self._p(t_len), self._p(z_len), B, Tt, Tz,
                                    ctypes.byref(opts), self._p(z), self._p(mel), self._stream()))
        self._dirty = True                  # ActNorm parameters and BN moving averages changed
        return mel


class InferenceSession:
    """VAENAR.inference for a fixed (B, T_text, T_z, rf) captured once into a CUDA graph: the steady-state
    serving call is  pinned-host -> device copies, one graph launch, device -> pinned-host copy of the mel."""

    def __init__(self, model: VAENAR, B, T_text, T_z, rf=2, return_alignments=False, seed=0):
        model._require_cuda()
        self.m, self.B, self.Tt, self.Tz, self.rf = model, B, T_text, T_z, rf
        dev = model.device
        hp = model._hp
        self.texts = torch.zeros(B, T_text, dtype=torch.int32, device=dev)
        self.t_len = torch.ones(B, dtype=torch.int32, device=dev)
        self.z_len = torch.ones(B, dtype=torch.int32, device=dev)
        self.eps = torch.zeros(B, T_z, hp.latent_dim, dtype=torch.float32, device=dev)
        self.z = torch.empty_like(self.eps)
        self.emb = torch.empty(B, T_text, hp.enc_hidden, dtype=torch.float32, device=dev)
        self.mel = torch.empty(B, T_z * rf, hp.out_dim, dtype=torch.float32, device=dev)
        self.logp = torch.empty(B, dtype=torch.float32, device=dev)
        self.ali = (torch.empty(hp.dec_nblk, B, hp.dec_heads, T_z, T_text, dtype=torch.float32, device=dev)
                    if return_alignments else None)
        self.seed = seed
        self.calls = 0
        model._prepare(B, T_text, T_z, rf)
        # The captured graph bakes in raw pointers: the session owns its workspace (the model's shared one may be
        # re-allocated by a later call with a larger shape) and keeps the packed-operand arena alive.
        need = int(model._lib.vaenar_workspace_bytes(model._h, B, T_text, T_z, rf))
        if need < 0:
            check(-1)
        self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self._packed_ref = model._packed
        self.h_texts = torch.zeros(B, T_text, dtype=torch.int32).pin_memory()
        self.h_t_len = torch.ones(B, dtype=torch.int32).pin_memory()
        self.h_z_len = torch.ones(B, dtype=torch.int32).pin_memory()
        self.h_mel = torch.empty(B, T_z * rf, hp.out_dim, dtype=torch.float32).pin_memory()
        self.h_ali = torch.empty(self.ali.shape, dtype=torch.float32).pin_memory() if return_alignments else None
        self.graph = None
        self.launches_per_call = None

    def _launch(self):
        m = self.m
        self.z.copy_(self.eps)
        if m._packed is not self._packed_ref:
            raise VaenarError("the model's packed-operand arena was re-allocated after this session captured its pointer")
        check(m._lib.vaenar_inference(m._h, m._p(m._flat), m._p(m._packed), m._p(self.ws), self.ws.numel(),
                                      m._p(self.texts), m._p(self.t_len), m._p(self.z_len), self.B, self.Tt, self.Tz,
                                      self.rf, m._p(self.z), m._p(self.emb), m._p(self.mel), m._p(self.ali),
                                      m._p(self.logp), m._stream()))

    def capture(self):
        s = torch.cuda.Stream(self.m.device)
        s.wait_stream(torch.cuda.current_stream(self.m.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._launch()
        torch.cuda.current_stream(self.m.device).wait_stream(s)
        torch.cuda.synchronize(self.m.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._launch()
        return self

    def set_inputs(self, texts, text_lengths, mel_lengths):
        """Host-side staging into pinned buffers (outside the device timeline)."""
        self.h_texts.copy_(torch.as_tensor(texts, dtype=torch.int32))
        self.h_t_len.copy_(torch.as_tensor(text_lengths, dtype=torch.int32))
        self.h_z_len.copy_(((torch.as_tensor(mel_lengths).to(torch.int64) + self.rf - 1) // self.rf).to(torch.int32))

    def run_device(self, new_noise=True):
        """Device-resident step: fresh noise + one graph launch (inputs already in HBM)."""
        if new_noise:
            self.calls += 1
            check(self.m._lib.vaenar_randn(self.m._p(self.eps), self.eps.numel(), self.seed, self.calls, 1.0,
                                           self.m._stream()))
        if self.m._dirty:       # weights changed since the last pack (training, load_state_dict): the graph reads the same
            self.m._prepare(self.B, self.Tt, self.Tz, self.rf)          # arena, so re-packing in place is enough
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        return self.mel

    def run_e2e(self, new_noise=True):
        """Serving step: H2D of the pinned inputs, the graph, D2H of the mel into pinned host memory."""
        self.texts.copy_(self.h_texts, non_blocking=True)
        self.t_len.copy_(self.h_t_len, non_blocking=True)
        self.z_len.copy_(self.h_z_len, non_blocking=True)
        self.run_device(new_noise)
        self.h_mel.copy_(self.mel, non_blocking=True)
        if self.h_ali is not None:      # VAENAR.inference returns (mel, alignments) (models/models.py:199-210)
            self.h_ali.copy_(self.ali, non_blocking=True)
        return self.h_mel

    @property
    def h2d_bytes(self):
        return self.h_texts.numel() * 4 + self.h_t_len.numel() * 4 + self.h_z_len.numel() * 4

    @property
    def d2h_bytes(self):
        return self.h_mel.numel() * 4 + (self.h_ali.numel() * 4 if self.h_ali is not None else 0)

"""TEST INFRASTRUCTURE — a minimal ``tensorflow`` API shim (backed by torch, CPU) that is just
large enough to *execute the reference's own Python sources* (/root/reference/models/models.py and
modules/*.py) without TensorFlow.  It exists so that tests/golden/make_golden.py can run the
unmodified reference wiring (call order, masks, residual order, the mu/logvar swap, flow order ...)
and pin oracle/vaenar_oracle.py against it.  Leaf ops follow the published TF 2.2 / Keras
semantics; the shim contains no model logic.

Usage:  ``import oracle.tf_shim as shim; shim.install()`` then ``sys.path.insert(0, '/root/reference')``.
Never imported by the product.
"""
import inspect
import math
import sys
import types

import numpy as np
import torch

_DT = torch.float32
_state = types.SimpleNamespace(
    gen=torch.Generator().manual_seed(0),
    normal_log=[],        # every tf.random.normal draw, in call order
    normal_queue=[],      # if non-empty, tf.random.normal pops from here instead of drawing
    dropout_log=[],       # every Dropout mask (already scaled by 1/(1-rate)), in call order
    training_stack=[None],
    variables={},         # id(tensor) -> True
)


def reset(seed=0, dtype=torch.float32):
    global _DT
    _DT = dtype
    _state.gen = torch.Generator().manual_seed(seed)
    _state.normal_log.clear()
    _state.normal_queue.clear()
    _state.dropout_log.clear()
    _state.training_stack[:] = [None]


def state():
    return _state


# torch.Tensor conveniences the reference relies on (tf.Tensor.set_shape, Variable.assign)
torch.Tensor.set_shape = lambda self, shape: None
torch.Tensor.assign = lambda self, v: self.copy_(torch.as_tensor(v, dtype=self.dtype))
if not hasattr(torch.Tensor, "numpy_"):
    pass


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(x, dtype=dtype if dtype is not None else (_DT if isinstance(x, float) else None))


def _ints(x):
    if isinstance(x, torch.Tensor):
        return [int(v) for v in x.reshape(-1).tolist()]
    return [int(v) if not isinstance(v, int) else v for v in x]


# ---- dtypes ------------------------------------------------------------------------------
class _Dtypes:
    float32 = "float32"
    float64 = "float64"
    int32 = "int32"
    int64 = "int64"
    bool = "bool"


def _dtype(d):
    if isinstance(d, torch.dtype):
        return d
    return {"float32": _DT, "float64": torch.float64, "int32": torch.int32, "int64": torch.int64,
            "bool": torch.bool, "float": _DT}[str(d)]


# ---- basic ops ---------------------------------------------------------------------------
def shape(x):
    return tuple(int(s) for s in x.shape)


def cast(x, dtype=None, **kw):
    return _t(x).to(_dtype(dtype))


def constant(v, dtype=None, **kw):
    return _t(v, _dtype(dtype) if dtype is not None else None)


def reshape(x, shp, **kw):
    return x.reshape(_ints(shp))


def transpose(x, perm=None, **kw):
    return x.permute(*perm)


def expand_dims(x, axis, **kw):
    return _t(x).unsqueeze(axis)


def tile(x, multiples, **kw):
    return x.repeat(*_ints(multiples))


def concat(values, axis, **kw):
    return torch.cat(list(values), dim=axis)


def stack(values, axis=0, **kw):
    return torch.stack(list(values), dim=axis)


def split(x, num_or_size_splits, axis=0, **kw):
    n = num_or_size_splits
    return list(torch.chunk(x, n, dim=axis)) if isinstance(n, int) else list(torch.split(x, list(n), dim=axis))


def ones(shp, dtype=None, **kw):
    return torch.ones(_ints(shp), dtype=_dtype(dtype) if dtype else _DT)


def zeros(shp, dtype=None, **kw):
    return torch.zeros(_ints(shp), dtype=_dtype(dtype) if dtype else _DT)


def ones_like(x, dtype=None, **kw):
    return torch.ones_like(x, dtype=_dtype(dtype) if dtype else None)


def range_(start, limit=None, delta=1, dtype=None, **kw):
    if limit is None:
        start, limit = 0, start
    f = lambda v: v.item() if isinstance(v, torch.Tensor) else v
    return torch.arange(f(start), f(limit), f(delta), dtype=_dtype(dtype) if dtype else None)


def where(condition, x=None, y=None, **kw):
    return torch.where(condition, x, y)


def sequence_mask(lengths, maxlen=None, dtype="bool", name=None):
    lengths = _t(lengths).long()
    if maxlen is None:
        maxlen = int(lengths.max().item())
    maxlen = int(maxlen.item()) if isinstance(maxlen, torch.Tensor) else int(maxlen)
    m = torch.arange(maxlen)[None, :] < lengths[..., None]
    return m.to(_dtype(dtype))


def reduce_sum(x, axis=None, keepdims=False, **kw):
    if axis is None:
        return x.sum()
    return x.sum(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False, **kw):
    if axis is None:
        return x.mean()
    return x.mean(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)


def reduce_max(x, axis=None, **kw):
    return x.max() if axis is None else x.amax(dim=axis)


def reduce_std(x, axis=None, **kw):
    mean = x.mean(dim=axis, keepdim=True)
    return torch.sqrt(((x - mean) ** 2).mean(dim=axis))


def matmul(a, b, transpose_a=False, transpose_b=False, **kw):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def softmax(x, axis=-1, **kw):
    return torch.softmax(x, dim=axis)


def band_part(x, num_lower, num_upper, **kw):
    assert num_lower == -1 and num_upper == 0
    tri = torch.ones(x.shape[-2], x.shape[-1], dtype=torch.bool).tril()
    return x & tri if x.dtype == torch.bool else x * tri


def slogdet(x):
    s, l = torch.linalg.slogdet(x)
    return s, l


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, **kw):
    shp = shape
    shp = _ints(shp)
    if _state.normal_queue:
        std_n = _state.normal_queue.pop(0)
        assert list(std_n.shape) == shp, (std_n.shape, shp)
    else:
        std_n = torch.randn(shp, generator=_state.gen, dtype=torch.float64).to(_DT)
    _state.normal_log.append(std_n)
    sd = stddev.item() if isinstance(stddev, torch.Tensor) else stddev
    return std_n * sd + mean


def _unary(fn):
    return lambda x, *a, **kw: fn(_t(x, _DT) if not isinstance(x, torch.Tensor) else x)


def Variable(initial_value, trainable=True, dtype=None, name=None, **kw):
    v = _t(initial_value, _dtype(dtype) if dtype is not None else None).clone()
    if v.is_floating_point():
        v = v.to(_DT)
    _state.variables[id(v)] = True
    return v


# ---- Keras layers ------------------------------------------------------------------------
def _activation(a):
    if a is None:
        return None
    if isinstance(a, str):
        return {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "linear": None}[a]
    return a


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name

    def __call__(self, *args, **kwargs):
        # Keras call-context semantics (TF 2.2 base_layer.__call__): an explicit non-None
        # ``training`` wins; otherwise the value of the enclosing layer call is inherited.
        sig = inspect.signature(self.call)
        value = None
        if "training" in sig.parameters:
            bound = sig.bind_partial(*args, **kwargs)
            value = bound.arguments.get("training")
            if value is None:
                value = _state.training_stack[-1]
                if value is not None:
                    bound.arguments["training"] = value
            args, kwargs = bound.args, bound.kwargs
        else:
            value = _state.training_stack[-1]
        _state.training_stack.append(value)
        try:
            return self.call(*args, **kwargs)
        finally:
            _state.training_stack.pop()


def _trainable_variables(self):
    """tf.keras.Model.trainable_variables: every variable except the BatchNorm moving statistics, in attribute order."""
    return [v for k, v in extract_variables(self).items() if not (k.endswith("moving_mean") or k.endswith("moving_variance"))]


Layer.trainable_variables = property(_trainable_variables)
Model = Layer


def enable_grad(model):
    """Mark the trainable variables of a shim-backed model as autograd leaves (needed by GradientTape)."""
    for v in model.trainable_variables:
        v.requires_grad_(True)


class GradientTape:
    """tf.GradientTape over torch autograd: the forward pass inside the ``with`` block builds the graph as usual."""

    def __init__(self, *a, **kw):
        pass

    def __enter__(self):
        self._prev = torch.is_grad_enabled()
        torch.set_grad_enabled(True)
        return self

    def __exit__(self, *exc):
        torch.set_grad_enabled(self._prev)
        return False

    def gradient(self, target, sources):
        sources = list(sources)
        with torch.enable_grad():
            grads = torch.autograd.grad(target, sources, allow_unused=True)
        return [g for g in grads]


class Adam:
    """tf.keras.optimizers.Adam of TF 2.2 (ResourceApplyAdam): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t),
    m <- b1 m + (1-b1) g, v <- b2 v + (1-b2) g^2, var <- var - lr_t m / (sqrt(v) + eps); variables with a None gradient
    are skipped."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kw):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta_1, beta_2, epsilon
        self.iterations = 0
        self.slots = {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                m, vv = self.slots.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                vv.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                v.sub_(lr_t * m / (vv.sqrt() + self.eps))


def _glorot(shape_, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape_, generator=_state.gen, dtype=torch.float64) * 2 - 1) * lim).to(_DT)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", name=None, **kw):
        super().__init__(name=name)
        self.units, self.activation, self.use_bias = units, _activation(activation), use_bias
        self.kernel_initializer = kernel_initializer
        self.kernel = None
        self.bias = None

    def call(self, inputs):
        if self.kernel is None:
            din = inputs.shape[-1]
            k = torch.zeros(din, self.units, dtype=_DT) if self.kernel_initializer == "zeros" else _glorot(
                (din, self.units), din, self.units)
            self.kernel = Variable(k)
            if self.use_bias:
                self.bias = Variable(torch.zeros(self.units, dtype=_DT))
        y = inputs @ self.kernel
        if self.use_bias:
            y = y + self.bias
        return self.activation(y) if self.activation is not None else y


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, name=None, **kw):
        super().__init__(name=name)
        assert strides == 1 and padding.lower() == "same"
        self.filters, self.kernel_size, self.activation = filters, kernel_size, _activation(activation)
        self.kernel = None
        self.bias = None

    def call(self, inputs):
        k = self.kernel_size
        if self.kernel is None:
            cin = inputs.shape[-1]
            self.kernel = Variable(_glorot((k, cin, self.filters), k * cin, k * self.filters))
            self.bias = Variable(torch.zeros(self.filters, dtype=_DT))
        pad_l = (k - 1) // 2
        xt = torch.nn.functional.pad(inputs.transpose(1, 2), (pad_l, k - 1 - pad_l))
        y = torch.nn.functional.conv1d(xt, self.kernel.permute(2, 1, 0), self.bias).transpose(1, 2)
        return self.activation(y) if self.activation is not None else y


class LayerNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, name=None, **kw):
        super().__init__(name=name)
        self.epsilon = epsilon
        self.gamma = None
        self.beta = None

    def call(self, inputs, training=None):
        if self.gamma is None:
            d = inputs.shape[-1]
            self.gamma = Variable(torch.ones(d, dtype=_DT))
            self.beta = Variable(torch.zeros(d, dtype=_DT))
        mean = inputs.mean(dim=-1, keepdim=True)
        var = ((inputs - mean) ** 2).mean(dim=-1, keepdim=True)
        return (inputs - mean) / torch.sqrt(var + self.epsilon) * self.gamma + self.beta


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **kw):
        super().__init__(name=name)
        self.momentum, self.epsilon = momentum, epsilon
        self.gamma = None

    def call(self, inputs, training=None):
        if self.gamma is None:
            d = inputs.shape[-1]
            self.gamma = Variable(torch.ones(d, dtype=_DT))
            self.beta = Variable(torch.zeros(d, dtype=_DT))
            self.moving_mean = Variable(torch.zeros(d, dtype=_DT))
            self.moving_variance = Variable(torch.ones(d, dtype=_DT))
        if training:
            dims = tuple(range(inputs.dim() - 1))
            mean = inputs.mean(dim=dims)
            var = ((inputs - mean) ** 2).mean(dim=dims)
            with torch.no_grad():   # the moving averages are not differentiated (non-trainable variables)
                self.moving_mean.copy_(self.moving_mean * self.momentum + mean.detach() * (1 - self.momentum))
                self.moving_variance.copy_(self.moving_variance * self.momentum + var.detach() * (1 - self.momentum))
        else:
            mean, var = self.moving_mean, self.moving_variance
        return (inputs - mean) * torch.rsqrt(var + self.epsilon) * self.gamma + self.beta


class Dropout(Layer):
    def __init__(self, rate, name=None, **kw):
        super().__init__(name=name)
        self.rate = rate

    def call(self, inputs, training=None):
        if not training or self.rate == 0:
            return inputs
        keep = (torch.rand(inputs.shape, generator=_state.gen, dtype=torch.float64) >= self.rate).to(inputs.dtype)
        mask = keep / (1.0 - self.rate)
        _state.dropout_log.append(mask)
        return inputs * mask


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, name=None, **kw):
        super().__init__(name=name)
        self.embeddings = Variable(
            (torch.rand(input_dim, output_dim, generator=_state.gen, dtype=torch.float64) * 0.1 - 0.05).to(_DT))

    def call(self, inputs):
        return self.embeddings[inputs.long()]


class _Dummy:
    """Placeholder for TF symbols that only dead code references."""

    def __init__(self, *a, **kw):
        pass

    def __call__(self, *a, **kw):
        raise NotImplementedError("tf_shim: symbol not implemented (dead code in the reference)")


class _NS(types.ModuleType):
    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Dummy


def _build_module():
    tf = _NS("tensorflow")
    for k in ("float32", "float64", "int32", "int64", "bool"):
        setattr(tf, k, getattr(_Dtypes, k))
    tf.Tensor = torch.Tensor
    tf.shape, tf.cast, tf.constant, tf.reshape, tf.transpose = shape, cast, constant, reshape, transpose
    tf.expand_dims, tf.tile, tf.concat, tf.stack, tf.split = expand_dims, tile, concat, stack, split
    tf.ones, tf.zeros, tf.ones_like, tf.range, tf.where = ones, zeros, ones_like, range_, where
    tf.sequence_mask, tf.reduce_sum, tf.reduce_mean, tf.matmul = sequence_mask, reduce_sum, reduce_mean, matmul
    tf.square = _unary(torch.square)
    tf.exp = _unary(torch.exp)
    tf.sqrt = _unary(torch.sqrt)
    tf.abs = _unary(torch.abs)
    tf.pow = lambda x, y, **kw: torch.pow(_t(x, _DT) if not isinstance(x, torch.Tensor) else x, y)
    tf.identity = lambda x, **kw: x
    tf.stop_gradient = lambda x, **kw: x.detach()
    tf.logical_and = lambda a, b, **kw: a & b
    tf.Variable = Variable
    tf.function = lambda *a, **kw: (a[0] if a and callable(a[0]) else (lambda f: f))
    tf.GradientTape = GradientTape
    tf.TensorSpec = lambda *a, **kw: None

    m = _NS("tensorflow.math")
    m.log, m.exp, m.sqrt, m.sin, m.cos, m.tanh, m.sigmoid = (_unary(torch.log), _unary(torch.exp),
                                                              _unary(torch.sqrt), _unary(torch.sin),
                                                              _unary(torch.cos), _unary(torch.tanh),
                                                              _unary(torch.sigmoid))
    m.reduce_mean, m.reduce_sum, m.reduce_max, m.reduce_std = reduce_mean, reduce_sum, reduce_max, reduce_std
    m.softmax = softmax
    m.mod = lambda x, y, **kw: torch.remainder(x, y)
    m.equal = lambda x, y, **kw: x == y
    m.logical_and = lambda a, b, **kw: a & b
    m.maximum = lambda a, b, **kw: torch.maximum(_t(a, _DT) if not isinstance(a, torch.Tensor) else a,
                                                 torch.as_tensor(b, dtype=a.dtype))
    tf.math = m

    la = _NS("tensorflow.linalg")
    la.matmul, la.band_part, la.slogdet = matmul, band_part, slogdet
    la.inv = lambda x, **kw: torch.linalg.inv(x)
    tf.linalg = la

    nn = _NS("tensorflow.nn")
    nn.relu, nn.tanh, nn.sigmoid = torch.relu, torch.tanh, torch.sigmoid
    tf.nn = nn

    rnd = _NS("tensorflow.random")
    rnd.normal = random_normal
    tf.random = rnd

    keras = _NS("tensorflow.keras")
    layers = _NS("tensorflow.keras.layers")
    layers.Layer, layers.Dense, layers.Conv1D = Layer, Dense, Conv1D
    layers.LayerNormalization, layers.BatchNormalization = LayerNormalization, BatchNormalization
    layers.Dropout, layers.Embedding = Dropout, Embedding
    keras.layers = layers
    keras.Model = Model
    keras.initializers = _NS("tensorflow.keras.initializers")
    keras.optimizers = _NS("tensorflow.keras.optimizers")
    keras.optimizers.Adam = Adam
    tf.keras = keras
    tf.losses = _NS("tensorflow.losses")
    tf.nest = _NS("tensorflow.nest")
    return tf


def install():
    tf = _build_module()
    sys.modules["tensorflow"] = tf
    return tf


# ---- variable extraction -------------------------------------------------------------------
_GLOW_NAMES = ("actnorm", "linear", "affine_coupling")


def extract_variables(obj, prefix=""):
    """Walk a shim-backed reference model and return {attribute.path: tensor} with the naming of
    SURVEY.md Appendix B (lists -> index, the glow tuples -> actnorm/linear/affine_coupling)."""
    out = {}
    for attr, val in vars(obj).items():
        path = f"{prefix}{attr}"
        if isinstance(val, torch.Tensor):
            if id(val) in _state.variables:
                out[path] = val
        elif isinstance(val, Layer):
            out.update(extract_variables(val, path + "."))
        elif isinstance(val, (list, tuple)):
            for i, item in enumerate(val):
                if isinstance(item, Layer):
                    out.update(extract_variables(item, f"{path}.{i}."))
                elif isinstance(item, tuple) and len(item) == 3 and all(isinstance(x, Layer) for x in item):
                    for nm, sub in zip(_GLOW_NAMES, item):
                        out.update(extract_variables(sub, f"{path}.{i}.{nm}."))
    return out

"""A minimal stand-in for the ``tensorflow`` package, backed by torch-CPU, that is just big enough to
IMPORT AND EXECUTE the reference's unmodified layer code (/root/reference/deepctr/layers/*.py).

TEST INFRASTRUCTURE ONLY (used by tests/golden/generate.py in the build container, where
/root/reference exists but TensorFlow cannot be installed).  It lets the golden vectors under
tests/golden/ be produced by the reference's own ``call`` bodies - its own op sequence, reshapes,
transposes, splits, masks and paddings - instead of by a restatement.  What it cannot pin is the
arithmetic INSIDE TensorFlow's kernels (Eigen reduction order, FarmHash): those come from torch here
and are stated as unpinned in DESIGN.md.

``install()`` registers the fake modules in sys.modules and loads ``deepctr.layers`` from the
reference tree without running ``deepctr/__init__.py`` (which would start a thread that calls PyPI).

Round 2: the Keras functional surface the reference's COMPOSITION code needs (``Input``, ``Embedding``
with ``mask_zero``, ``Dense``, ``Lambda``, ``Concatenate``, ``Flatten``, ``Add``, ``Model``) is real
now, executed EAGERLY: ``Input(name=...)`` returns the batch that ``CTX.feed`` holds for that name, so
calling an unmodified builder (``deepctr.models.DeepFM(...)``) runs the reference's own
``feature_column.py`` / ``inputs.py`` / builder body on that batch and ``Model.outputs`` is the result.
``Layer.__call__`` follows tf.keras' implicit-mask protocol (``_keras_mask`` on tensors, ``mask=``
injected when ``call`` accepts it, ``compute_mask`` after the call, identity copy when a layer
returns its input) - that is what DIN's attention and the mask-based pooling layers consume.
"""
import importlib
import sys
import types

import numpy as np
import torch


class _DType(object):
    def __init__(self, name, torch_dtype):
        self.name, self.torch = name, torch_dtype

    def __eq__(self, other):
        if isinstance(other, _DType):
            return self.name == other.name
        if isinstance(other, torch.dtype):
            return self.torch == other
        return self.name == other

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "tf." + self.name


float32, int32, int64, bool_, string = (_DType("float32", torch.float32), _DType("int32", torch.int32),
                                        _DType("int64", torch.int64), _DType("bool", torch.bool),
                                        _DType("string", None))
_BY_NAME = {"float32": float32, "int32": int32, "int64": int64, "bool": bool_, "string": string}


def as_dtype(d):
    if isinstance(d, _DType):
        return d
    if isinstance(d, str) and d in _BY_NAME:
        return _BY_NAME[d]
    raise TypeError("unknown dtype %r" % (d,))


def _td(d):
    return as_dtype(d).torch if not isinstance(d, torch.dtype) else d


class StrTensor(object):
    """numpy array of bytes standing in for a tf.string tensor."""

    def __init__(self, arr):
        self.arr = np.asarray(arr, dtype=object)
        self.dtype = string
        self.shape = self.arr.shape

    def get_shape(self):
        return self.shape


def _t(x):
    if isinstance(x, (list, tuple)):
        if len(x) and isinstance(x[0], torch.Tensor):
            return torch.stack(list(x))
        return torch.as_tensor(np.asarray(x))
    return x


def constant(value, dtype=None, shape=None, name=None):
    a = np.asarray(value)
    if a.dtype.kind in ("U", "S", "O"):
        return StrTensor(np.vectorize(lambda s: s.encode() if isinstance(s, str) else s, otypes=[object])(a))
    t = torch.as_tensor(a)
    if dtype is not None:
        t = t.to(_td(dtype))
    elif t.dtype == torch.float64:
        t = t.float()
    return t


def zeros(shape, dtype=float32):
    return torch.zeros(tuple(shape), dtype=_td(dtype))


def as_string(x, **kw):
    if isinstance(x, StrTensor):
        return x
    a = x.numpy()
    return StrTensor(np.vectorize(lambda v: str(int(v)).encode("ascii"), otypes=[object])(a))


def split(value, num_or_size_splits, axis=0, **kw):
    if isinstance(num_or_size_splits, int):
        return list(torch.chunk(value, num_or_size_splits, dim=axis))
    return list(torch.split(value, list(num_or_size_splits), dim=axis))


def matmul(a, b, transpose_a=False, transpose_b=False, **kw):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def reshape(tensor, shape=None, **kw):
    return tensor.reshape(tuple(int(s) for s in shape))


def transpose(a, perm=None, **kw):
    return a.permute(*perm) if perm is not None else a.t()


def concat(values, axis=0, **kw):
    return torch.cat(list(values), dim=axis)


def expand_dims(x, axis=-1, **kw):
    return x.unsqueeze(axis)


def squeeze(x, axis=None, **kw):
    return x.squeeze(axis) if axis is not None else x.squeeze()


def stack(values, axis=0, **kw):
    return torch.stack(list(values), dim=axis)


def tile(x, multiples, **kw):
    return x.repeat(*[int(m) for m in multiples])


def cast(x, dtype, **kw):
    return x.to(_td(dtype))


def where(cond, x=None, y=None, **kw):
    return torch.where(cond, x, y)


def ones_like(x, dtype=None, **kw):
    return torch.ones_like(x, dtype=_td(dtype) if dtype is not None else None)


def zeros_like(x, dtype=None, **kw):
    return torch.zeros_like(x, dtype=_td(dtype) if dtype is not None else None)


def sequence_mask(lengths, maxlen=None, dtype=bool_, **kw):
    lengths = lengths.to(torch.int64)
    m = torch.arange(int(maxlen)).reshape((1,) * lengths.dim() + (-1,)) < lengths.unsqueeze(-1)
    return m.to(_td(dtype))


def tensordot(a, b, axes, **kw):
    if isinstance(axes, int):
        return torch.tensordot(a, b, dims=axes)
    a_ax, b_ax = axes
    a_ax = [a_ax] if isinstance(a_ax, int) else list(a_ax)
    b_ax = [b_ax] if isinstance(b_ax, int) else list(b_ax)
    return torch.tensordot(a, b, dims=(a_ax, b_ax))


def einsum(eq, *ops, **kw):
    return torch.einsum(eq, *ops)


def _reduce(fn):
    def f(input_tensor, axis=None, keepdims=False, name=None):     # NB: TF2 signature - `keep_dims` raises
        if axis is None:
            return fn(input_tensor)
        return fn(input_tensor, dim=axis, keepdim=keepdims)
    return f


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)


def reduce_max(input_tensor, axis=None, keepdims=False, name=None):
    if axis is None:
        return input_tensor.max()
    return torch.amax(input_tensor, dim=axis, keepdim=keepdims)


def divide(x, y, name=None):
    return x / y


def multiply(x, y, name=None):
    return x * y


def not_equal(x, y, name=None):
    if isinstance(x, StrTensor):
        yv = y.arr if isinstance(y, StrTensor) else np.asarray(y)
        return torch.as_tensor((x.arr != yv.reshape(-1)[0]).astype(np.bool_))
    return x != y


def sigmoid(x, name=None):
    return torch.sigmoid(x)


def square(x, name=None):
    return x * x


def squeeze_(x):
    return x


class _NN(types.ModuleType):
    pass


def _conv1d(x, filters=None, stride=1, padding="VALID", **kw):
    # x [B, W, Cin] (NWC), filters [kw, Cin, Cout]
    w = filters.permute(2, 1, 0)           # [Cout, Cin, kw]
    return torch.nn.functional.conv1d(x.permute(0, 2, 1), w, stride=stride).permute(0, 2, 1)


def _bias_add(value, bias, data_format=None, name=None):
    return value + bias


def _softmax(logits, axis=-1, name=None):       # NB: `dim=` raises TypeError as in TF2
    return torch.softmax(logits, dim=axis)


# ---- initializers / regularizers ------------------------------------------------------------------
class _Init(object):
    def __init__(self, seed=None, **kw):
        self.seed = seed

    def _rng(self):
        return np.random.RandomState(self.seed)


class Zeros(_Init):
    def __call__(self, shape, dtype=None):
        return torch.zeros(tuple(shape))


class Ones(_Init):
    def __call__(self, shape, dtype=None):
        return torch.ones(tuple(shape))


class Constant(_Init):
    def __init__(self, value=0.0, **kw):
        self.value = value

    def __call__(self, shape, dtype=None):
        return torch.full(tuple(shape), float(self.value))


class TruncatedNormal(_Init):
    def __init__(self, mean=0.0, stddev=0.05, seed=None, **kw):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        return torch.as_tensor(self._rng().normal(self.mean, self.stddev, size=tuple(shape)).astype(np.float32))


RandomNormal = TruncatedNormal


class glorot_normal(_Init):
    def __call__(self, shape, dtype=None):
        return torch.as_tensor(self._rng().normal(0, 0.1, size=tuple(shape)).astype(np.float32))


glorot_uniform = glorot_normal


class l2(object):
    def __init__(self, l2=0.01):
        self.l2 = l2


# ---- eager-execution context ---------------------------------------------------------------------------
class _Ctx(object):
    """feed: Input name -> numpy batch; training: value injected into ``call(training=...)``;
    rng/grad: when rng is set every add_weight draws O(1) values from it (so that zero-initialised
    terms carry signal) and, with grad, the weights are autograd leaves; layers: creation-ordered
    registry; depth: >0 while inside another layer's build/call (nested layers)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.feed, self.training, self.rng, self.grad, self.scale = {}, None, None, False, 1.0
        self.layers, self.uid, self.depth = [], {}, 0


CTX = _Ctx()


def to_snake_case(name):
    """keras.utils.generic_utils.to_snake_case (public Keras behaviour)."""
    import re
    s1 = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    s2 = re.sub("([a-z])([A-Z])", r"\1_\2", s1).lower()
    return "private" + s2 if s2[0] == "_" else s2


def _unique_name(base):
    n = CTX.uid.get(base, 0)
    CTX.uid[base] = n + 1
    return base if n == 0 else "%s_%d" % (base, n)


def _flat(x):
    if isinstance(x, (list, tuple)):
        out = []
        for e in x:
            out.extend(_flat(e))
        return out
    return [x]


# ---- keras layers ------------------------------------------------------------------------------------
def _shape_of(x):
    if isinstance(x, (list, tuple)):
        return [_shape_of(e) for e in x]
    return tuple([None] + list(x.shape[1:]))


class Layer(object):
    def __init__(self, name=None, trainable=True, dtype=None, **kwargs):
        self.name = name or _unique_name(to_snake_case(self.__class__.__name__))
        self.built = False
        self.trainable = trainable
        self._w = []
        self._nested = CTX.depth > 0
        CTX.layers.append(self)
        if not hasattr(self, "supports_masking"):
            self.supports_masking = False

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, regularizer=None,
                   trainable=True, **kw):
        shape = tuple(int(s) for s in shape)
        if CTX.rng is not None:
            # O(1) values instead of the requested initialiser (zeros / 1e-4): every term carries signal
            if name == "embeddings" or len(shape) < 2:
                std = 0.3 if name == "embeddings" else 0.1
            else:
                std = 1.0 / np.sqrt(float(np.prod(shape[:-1])))
            t = torch.as_tensor(CTX.rng.normal(0, std * CTX.scale, size=shape).astype(np.float32))
        else:
            init = initializer() if isinstance(initializer, type) else (initializer or Zeros())
            t = init(shape).float()
        if CTX.grad and trainable:
            t.requires_grad_(True)
        self._w.append((name, t))
        return t

    @property
    def weights(self):
        return list(self._w)

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, **kwargs):
        import inspect
        CTX.depth += 1
        try:
            if not self.built:
                self.build(_shape_of(inputs))
                self.built = True
            try:
                params = inspect.signature(self.call).parameters
            except (TypeError, ValueError):
                params = {}
            expects_mask = "mask" in params
            explicit_cm = type(self).compute_mask is not Layer.compute_mask
            # tf.keras base_layer._get_input_masks
            input_masks = None
            if self.supports_masking or expects_mask:
                if kwargs.get("mask") is not None:
                    input_masks = kwargs["mask"]
                else:
                    ms = [getattr(t, "_keras_mask", None) for t in _flat(inputs)]
                    if any(m is not None for m in ms):
                        input_masks = ms if isinstance(inputs, (list, tuple)) else ms[0]
                        if expects_mask:
                            kwargs["mask"] = input_masks
            if "training" in params and "training" not in kwargs and CTX.training is not None:
                kwargs["training"] = CTX.training
            out = self.call(inputs, **kwargs)
            # a layer that returns its input gets an identity copy (Keras wraps it in tf.identity)
            if isinstance(out, torch.Tensor) and any(out is t for t in _flat(inputs)):
                out = out.view_as(out)
            # tf.keras base_layer._set_mask_metadata
            if (self.supports_masking or explicit_cm) and isinstance(out, torch.Tensor) \
                    and getattr(out, "_keras_mask", None) is None:
                om = self.compute_mask(inputs, input_masks)
                if om is not None:
                    out._keras_mask = om
            self._last_in, self._last_out = inputs, out
            return out
        finally:
            CTX.depth -= 1

    def call(self, inputs, **kwargs):
        return inputs

    def get_config(self):
        return {"name": self.name}

    def compute_mask(self, inputs, mask=None):
        return mask if self.supports_masking else None


class Activation(Layer):
    def __init__(self, activation, **kw):
        Layer.__init__(self, **kw)
        self.activation = activation

    def call(self, x, **kwargs):       # NB: no `training` kwarg -> the reference's TypeError fallback runs
        return {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
                "linear": lambda v: v, None: lambda v: v}[self.activation](x)


class Dropout(Layer):
    def __init__(self, rate, seed=None, **kw):
        Layer.__init__(self, **kw)

    def call(self, x, training=None, **kwargs):
        return x


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, **kw):
        Layer.__init__(self, **kw)
        self.epsilon, self.center, self.scale = epsilon, center, scale
        self.moving_mean = self.moving_variance = None

    def call(self, x, training=None, **kwargs):
        n = x.shape[-1]
        if self.moving_mean is None:
            if CTX.rng is not None:
                self.moving_mean = torch.as_tensor(CTX.rng.normal(0, 0.2, size=n).astype(np.float32))
                self.moving_variance = torch.as_tensor(np.abs(CTX.rng.normal(1.0, 0.2, size=n)).astype(np.float32))
            else:
                self.moving_mean, self.moving_variance = torch.zeros(n), torch.ones(n)
        if training:
            red = tuple(range(x.dim() - 1))
            mean, var = x.mean(dim=red), x.var(dim=red, unbiased=False)
        else:
            mean, var = self.moving_mean, self.moving_variance
        return (x - mean) / torch.sqrt(var + self.epsilon)


class Flatten(Layer):
    def call(self, x, **kw):
        return x.flatten(1)


class Add(Layer):
    def call(self, xs, **kw):
        # keras.layers.merge._Merge.call: lower-rank operands are expanded at axis 1 up to the max rank
        nd = max(x.dim() for x in xs)
        ys = []
        for x in xs:
            while x.dim() < nd:
                x = x.unsqueeze(1)
            ys.append(x)
        out = ys[0]
        for x in ys[1:]:
            out = out + x
        return out


class _Dummy(Layer):
    def __init__(self, *a, **kw):
        Layer.__init__(self)


def Input(shape=None, name=None, dtype=None, **kw):
    """Eager stand-in for tf.keras.layers.Input: the batch CTX.feed holds under ``name``, shaped
    [B] + shape and cast to ``dtype``."""
    if name not in CTX.feed:
        raise KeyError("no data fed for Input %r" % (name,))
    a = np.asarray(CTX.feed[name])
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    dt = as_dtype(dtype or "float32")
    if dt == string:
        t = constant(a.reshape((a.shape[0],) + shape))
    else:
        t = torch.as_tensor(a.reshape((a.shape[0],) + shape)).to(dt.torch)
    t._keras_input_name = name
    return t


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, embeddings_initializer=None, embeddings_regularizer=None,
                 mask_zero=False, name=None, **kw):
        Layer.__init__(self, name=name)
        self.input_dim, self.output_dim, self.mask_zero = int(input_dim), int(output_dim), mask_zero
        self.supports_masking = mask_zero
        self.embeddings_initializer = embeddings_initializer
        self.embeddings_regularizer = embeddings_regularizer

    def build(self, input_shape):
        self.embeddings = self.add_weight(name="embeddings", shape=(self.input_dim, self.output_dim),
                                          initializer=self.embeddings_initializer)
        self.built = True

    def call(self, inputs, **kw):
        idx = inputs.to(torch.int64)
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self.input_dim):
            raise IndexError("InvalidArgument: indices out of range [0, %d)" % self.input_dim)   # TF-CPU behaviour
        return self.embeddings[idx]

    def compute_mask(self, inputs, mask=None):
        if not self.mask_zero:
            return None
        return inputs != 0


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, name=None, **kw):
        Layer.__init__(self, name=name)
        self.units, self.activation, self.use_bias = int(units), activation, use_bias
        self.kernel_initializer = kernel_initializer
        self.supports_masking = True

    def build(self, input_shape):
        self.kernel = self.add_weight(name="kernel", shape=(int(input_shape[-1]), self.units),
                                      initializer=self.kernel_initializer or glorot_uniform())
        self.bias = self.add_weight(name="bias", shape=(self.units,), initializer=Zeros()) if self.use_bias else None
        self.built = True

    def call(self, inputs, **kw):
        y = torch.matmul(inputs, self.kernel)
        if self.bias is not None:
            y = y + self.bias
        return {None: lambda v: v, "linear": lambda v: v, "relu": torch.relu, "sigmoid": torch.sigmoid,
                "tanh": torch.tanh}[self.activation](y)


class Lambda(Layer):
    def __init__(self, function, name=None, **kw):
        Layer.__init__(self, name=name)
        self.function = function

    def call(self, inputs, **kw):
        return self.function(inputs)


class Concatenate(Layer):
    def __init__(self, axis=-1, name=None, **kw):
        Layer.__init__(self, name=name)
        self.axis = axis

    def call(self, inputs, **kw):
        return torch.cat(list(inputs), dim=self.axis)


class Model(object):
    """tf.keras.models.Model of an eagerly executed graph: holds the already computed outputs and
    the creation-ordered layer registry of this build."""

    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs, self.name = inputs, outputs, name
        self.layers = list(CTX.layers)


# ---- lookup ops --------------------------------------------------------------------------------------
class TextFileInitializer(object):
    def __init__(self, filename, key_dtype, key_index, value_dtype, value_index, delimiter="\t", **kw):
        import csv
        self.table = {}
        with open(filename, newline="") as fh:
            for row in csv.reader(fh, delimiter=delimiter):
                if len(row) > max(key_index, value_index):
                    self.table[row[key_index].encode()] = int(row[value_index])


class StaticHashTable(object):
    def __init__(self, initializer, default_value=0, **kw):
        self.table, self.default = initializer.table, default_value

    def lookup(self, x):
        return torch.as_tensor(np.vectorize(lambda s: self.table.get(s, self.default), otypes=[np.int64])(x.arr))


def _to_hash_bucket_fast(x, num_buckets, name=None):
    from oracle import farmhash       # NOT a pin: TensorFlow's FarmHash is unavailable here
    return torch.as_tensor(np.vectorize(lambda s: farmhash.fingerprint64(s) % num_buckets, otypes=[np.int64])(x.arr))


# ---- install -----------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(reference_root="/root/reference"):
    torch.Tensor.get_shape = lambda self: self.shape
    g = globals()
    tf = _mod("tensorflow", __version__="2.10.0", **{k: g[k] for k in (
        "float32", "int32", "int64", "string", "as_dtype", "constant", "zeros", "as_string", "split", "matmul",
        "reshape", "transpose", "concat", "expand_dims", "squeeze", "stack", "tile", "cast", "where", "ones_like",
        "zeros_like", "sequence_mask", "tensordot", "einsum", "reduce_sum", "reduce_mean", "reduce_max", "divide",
        "multiply", "not_equal", "sigmoid", "square")})
    tf.bool = bool_
    tf.nn = _mod("tensorflow.nn", conv1d=_conv1d, bias_add=_bias_add, softmax=_softmax, relu=torch.relu,
                 tanh=torch.tanh)
    tf.strings = _mod("tensorflow.strings", to_hash_bucket_fast=_to_hash_bucket_fast)
    K = _mod("tensorflow.keras.backend", ndim=lambda x: x.dim(),
             repeat_elements=lambda x, rep, axis: torch.repeat_interleave(x, rep, dim=axis),
             concatenate=lambda xs, axis=-1: torch.cat(list(xs), dim=axis),
             all=lambda x, axis=None, keepdims=False: torch.all(x, dim=axis, keepdim=keepdims),
             batch_dot=None)
    layers = _mod("tensorflow.keras.layers", Layer=Layer, Activation=Activation, Dropout=Dropout,
                  BatchNormalization=BatchNormalization, Flatten=Flatten, Add=Add, Lambda=Lambda, Dense=Dense,
                  Concatenate=Concatenate, Conv2D=_Dummy, MaxPooling2D=_Dummy, LSTM=_Dummy, Embedding=Embedding,
                  Input=Input)
    inits = _mod("tensorflow.keras.initializers", RandomNormal=RandomNormal, Zeros=Zeros, Ones=Ones,
                 TruncatedNormal=TruncatedNormal, glorot_normal=glorot_normal, glorot_uniform=glorot_uniform)
    regs = _mod("tensorflow.keras.regularizers", l2=l2)
    models = _mod("tensorflow.keras.models", Model=Model)
    keras = _mod("tensorflow.keras", backend=K, layers=layers, initializers=inits, regularizers=regs, models=models)
    tf.keras = keras
    py = _mod("tensorflow.python")
    ops = _mod("tensorflow.python.ops")
    v2 = _mod("tensorflow.python.ops.init_ops_v2", Zeros=Zeros, Ones=Ones, Constant=Constant,
              TruncatedNormal=TruncatedNormal, glorot_normal=glorot_normal, glorot_uniform=glorot_uniform)
    lk = _mod("tensorflow.python.ops.lookup_ops", TextFileInitializer=TextFileInitializer,
              StaticHashTable=StaticHashTable)
    pl = _mod("tensorflow.python.layers", utils=types.ModuleType("utils"))
    py.ops, ops.init_ops_v2, ops.lookup_ops, py.layers = ops, v2, lk, pl
    tf.python = py
    # the reference package, without executing deepctr/__init__.py (it starts a PyPI version check)
    pkg = types.ModuleType("deepctr")
    pkg.__path__ = [reference_root + "/deepctr"]
    sys.modules["deepctr"] = pkg
    # the model packages, without their __init__ (it imports all 30 builders and their Keras symbols)
    for n, sub in (("deepctr.models", "/deepctr/models"), ("deepctr.models.sequence", "/deepctr/models/sequence")):
        mp = types.ModuleType(n)
        mp.__path__ = [reference_root + sub]
        sys.modules[n] = mp
    for n in ("deepctr.contrib", "deepctr.contrib.rnn", "deepctr.contrib.rnn_v2", "deepctr.contrib.utils"):
        m = _mod(n, dynamic_rnn=None, QAAttGRUCell=None, VecAttGRUCell=None)
        if n == "deepctr.contrib":
            m.__path__ = []
    return importlib.import_module("deepctr.layers")

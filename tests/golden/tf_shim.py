'''
A minimal ``tensorflow`` stand-in (TEST INFRASTRUCTURE ONLY) that lets the reference's own, UNMODIFIED
``composer/models/transformer.py`` execute in this container, where TensorFlow cannot be installed.

What it is for: pinning ``oracle/transformer_oracle.py``.  The reference model is ~45 TensorFlow entry points called
from pure Python.  Each of them is restated here on torch-CPU float64 with TensorFlow's published semantics (one
line each; the non-obvious ones cite the TF / Keras source they follow).  With this module installed as
``tensorflow`` in ``sys.modules``, ``composer.models.transformer.Transformer`` -- the reference's class, from the
reference's file -- runs forward, ``past=`` decoding, and its own ``train()`` loop (GradientTape + Adam), and
``tests/golden/make_model_golden.py`` records what it produces.  The composition of the ops (mask arithmetic,
the residual wiring, head split / merge, cache concatenation, position ids, tied logits, loss reduction) is then
the reference's code, not a restatement; what remains a restatement is the per-op arithmetic below.

Arithmetic is float64 throughout (``tf.float32`` maps to torch.float64) so that comparisons with the oracle are not
limited by rounding.  Nothing under ``composer_b200/`` imports this file.
'''

import contextlib
import math
import sys
import types

import numpy as np
import torch

FLOAT = torch.float64


class TensorShape(list):
    '''``tf.TensorShape``: only ``as_list`` and indexing are used (transformer.py:30).'''

    def as_list(self):
        return list(self)


class Tensor(torch.Tensor):
    '''torch tensor with the two TF-only attributes the reference touches: ``shape.as_list()`` and ``numpy()``.'''

    @property
    def shape(self):
        return TensorShape(self.size())

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()

    def __format__(self, spec):       # an EagerTensor scalar formats like its value (transformer.py:939)
        return format(self.detach().as_subclass(torch.Tensor).item(), spec) if self.dim() == 0 else repr(self)

    def __float__(self):
        return float(self.detach().as_subclass(torch.Tensor).item())


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        out = x if dtype is None else x.to(dtype)
    else:
        out = torch.as_tensor(np.asarray(x))
        if out.is_floating_point():
            out = out.to(FLOAT)
        if dtype is not None:
            out = out.to(dtype)
    return out.as_subclass(Tensor)


_DTYPES = {'float32': FLOAT, 'float64': FLOAT, 'int32': torch.int64, 'int64': torch.int64, 'bool': torch.bool}


class DType:
    def __init__(self, name):
        self.name = name
        self.torch = _DTYPES[name]

    def __repr__(self):
        return 'tf.' + self.name


def _dtype(d):
    if d is None:
        return None
    if isinstance(d, DType):
        return d.torch
    if isinstance(d, torch.dtype):
        return d
    return _DTYPES[str(d)]


# ---------------------------------------------------------------------------
# tf.* functions (the ones transformer.py calls)
# ---------------------------------------------------------------------------

def cast(x, dtype):
    return _t(x, _dtype(dtype))


def shape(x):
    return TensorShape(_t(x).size())


def reshape(x, new_shape):
    return _t(x).reshape([int(s) for s in new_shape])


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.permute(*reversed(range(x.dim())))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def tf_range(start, limit=None, delta=1, dtype=None):
    if limit is None:
        start, limit = 0, start
    return _t(torch.arange(int(start), int(limit), int(delta), dtype=_dtype(dtype) or torch.int64))


def ones(shape_, dtype=None):
    return _t(torch.ones([int(s) for s in shape_], dtype=_dtype(dtype) or FLOAT))


def band_part(x, num_lower, num_upper):
    '''tf.linalg.band_part: keep element (m, n) when (num_lower < 0 or m - n <= num_lower) and (num_upper < 0 or n - m <= num_upper).'''

    x = _t(x)
    m = torch.arange(x.shape[-2])[:, None]
    n = torch.arange(x.shape[-1])[None, :]
    keep = torch.ones(x.shape[-2], x.shape[-1], dtype=torch.bool)
    if num_lower >= 0:
        keep &= (m - n) <= num_lower
    if num_upper >= 0:
        keep &= (n - m) <= num_upper
    return x * keep.to(x.dtype)


def split(x, num, axis=0):
    return [_t(p) for p in torch.chunk(_t(x), num, dim=axis)]


def unstack(x, axis=0):
    return [_t(p) for p in torch.unbind(_t(x), dim=axis)]


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis).as_subclass(Tensor)


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis).as_subclass(Tensor)


def gather(params, indices):
    '''tf.gather on axis 0.  The CPU kernel raises on an out-of-range index (the GPU kernel returns zeros).'''

    params, indices = _t(params), _t(indices).long()
    if indices.numel() and (int(indices.min()) < 0 or int(indices.max()) >= params.shape[0]):
        raise IndexError('indices out of range in tf.gather (InvalidArgumentError on TF-CPU)')
    return params[indices]


def softmax(logits, axis=-1):
    return torch.softmax(_t(logits), dim=axis)


def pad(x, paddings):
    flat = []
    for before, after in reversed(list(paddings)):
        flat += [int(before), int(after)]
    return torch.nn.functional.pad(_t(x), flat)


def reduce_mean(x, axis=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis)


def argmax(x, axis=None):
    return torch.argmax(_t(x), dim=axis)


def is_tensor(x):
    return isinstance(x, torch.Tensor)


# ---------------------------------------------------------------------------
# Variables, GradientTape
# ---------------------------------------------------------------------------

def _variable(value, name=None):
    tensor = _t(value).detach().clone()
    if tensor.is_floating_point():
        tensor.requires_grad_(True)
    tensor = tensor.as_subclass(Tensor)
    tensor.var_name = name
    return tensor


class Variable:
    '''``tf.Variable`` as the train loop uses it: an integer counter (``checkpoint.step``, ``checkpoint.epoch``).'''

    def __init__(self, value):
        self.value = value

    def assign_add(self, delta):
        self.value += delta

    def assign(self, value):
        self.value = value

    def numpy(self):
        return self.value

    def __int__(self):
        return int(self.value)


class GradientTape:
    '''tape.gradient(loss, variables) == d loss / d variable, here by torch autograd over the same graph.'''

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        sources = list(sources)
        grads = torch.autograd.grad(target, sources, allow_unused=True)
        return [None if g is None else g.as_subclass(Tensor) for g in grads]


# ---------------------------------------------------------------------------
# tf.keras
# ---------------------------------------------------------------------------

class TruncatedNormal:
    '''Values are overwritten by the golden generator; drawn here only so that shapes exist.'''

    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape_, dtype=None):
        generator = torch.Generator().manual_seed(0)
        return torch.nn.init.trunc_normal_(torch.empty(*shape_, dtype=FLOAT), self.mean, self.stddev,
                                           self.mean - 2 * self.stddev, self.mean + 2 * self.stddev, generator=generator)


class GlorotUniform:
    def __call__(self, shape_, dtype=None):
        raise NotImplementedError('relative attention is broken in the reference (transformer.py:281-286)')


def zeros_initializer():
    return lambda shape_, dtype=None: torch.zeros(*shape_, dtype=FLOAT)


def ones_initializer():
    return lambda shape_, dtype=None: torch.ones(*shape_, dtype=FLOAT)


class Layer:
    '''keras.layers.Layer: name, lazy ``build(input_shape)`` on the first call, ``add_weight``, variable tracking.'''

    def __init__(self, name=None, **kwargs):
        self.name = name or type(self).__name__.lower()
        self.built = False
        self._weights = []

    def add_weight(self, name=None, shape=None, initializer=None, **kwargs):
        value = initializer(list(shape)) if initializer is not None else torch.zeros(*shape, dtype=FLOAT)
        variable = _variable(value, name)
        self._weights.append(variable)
        return variable

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, *args, **kwargs):
        # Keras converts array-like inputs to tensors before ``call``
        if isinstance(inputs, (np.ndarray, torch.Tensor)):
            inputs = _t(inputs)
        elif isinstance(inputs, (list, tuple)):
            inputs = type(inputs)(_t(i) if isinstance(i, (np.ndarray, torch.Tensor)) else i for i in inputs)
        if not self.built:
            if isinstance(inputs, (list, tuple)):
                input_shape = [TensorShape(i.size()) if isinstance(i, torch.Tensor) else None for i in inputs]
            else:
                input_shape = TensorShape(inputs.size()) if isinstance(inputs, torch.Tensor) else None
            self.build(input_shape)
            self.built = True
        return self.call(inputs, *args, **kwargs)

    def _sublayers(self):
        for value in self.__dict__.values():
            if isinstance(value, Layer):
                yield value
            elif isinstance(value, (list, tuple)):
                for item in value:
                    if isinstance(item, Layer):
                        yield item

    def named_variables(self, prefix=''):
        '''(path, variable) pairs, paths as Keras names them: <layer name>/.../<weight name>.'''

        for variable in self._weights:
            yield prefix + variable.var_name, variable
        for layer in self._sublayers():
            yield from layer.named_variables(prefix + layer.name + '/')

    @property
    def trainable_variables(self):
        return [variable for _, variable in self.named_variables()]


class Model(Layer):
    def compile(self, optimizer=None, loss=None, metrics=None):
        self.optimizer, self.loss = optimizer, loss


class Dropout(Layer):
    '''keras.layers.Dropout: identity unless training; tf.nn.dropout scales the kept elements by 1 / (1 - rate).'''

    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate

    def call(self, inputs, training=False):
        if not training or self.rate == 0:
            return inputs
        keep = (torch.rand(inputs.size(), dtype=FLOAT) >= self.rate).to(FLOAT)
        return inputs * keep / (1.0 - self.rate)


class LayerNormalization(Layer):
    '''
    keras.layers.LayerNormalization(axis=-1): mean and (biased) variance over the last axis, then
    tf.nn.batch_normalization: (x - mean) * rsqrt(variance + epsilon) * gamma + beta.
    '''

    def __init__(self, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight('gamma', [input_shape[-1]], ones_initializer())
        self.beta = self.add_weight('beta', [input_shape[-1]], zeros_initializer())

    def call(self, inputs):
        mean = inputs.mean(dim=-1, keepdim=True)
        variance = ((inputs - mean) ** 2).mean(dim=-1, keepdim=True)
        return (inputs - mean) * torch.rsqrt(variance + self.epsilon) * self.gamma + self.beta


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, embeddings_initializer=None, **kwargs):
        super().__init__(**kwargs)
        self.input_dim, self.output_dim, self.initializer = input_dim, output_dim, embeddings_initializer

    def build(self, input_shape):
        self.embeddings = self.add_weight('embeddings', [self.input_dim, self.output_dim], self.initializer)

    def call(self, inputs):
        return gather(self.embeddings, inputs)


class SparseCategoricalCrossentropy:
    '''from_logits=True, reduction SUM_OVER_BATCH_SIZE: the mean over every label of -log softmax(logits)[label].'''

    def __init__(self, from_logits=False):
        assert from_logits

    def __call__(self, y_true, y_pred):
        logp = torch.log_softmax(_t(y_pred), dim=-1)
        picked = torch.gather(logp, -1, _t(y_true).long().unsqueeze(-1)).squeeze(-1)
        return -picked.mean()


class Adam:
    '''
    TF-2 Keras Adam (keras/optimizer_v2/adam.py, ``_resource_apply_dense`` -> ``ResourceApplyAdam``, amsgrad off):
        lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t);  m = beta_1 m + (1 - beta_1) g;
        v = beta_2 v + (1 - beta_2) g^2;  var -= lr_t * m / (sqrt(v) + epsilon),  epsilon = 1e-7, t from 1.
    '''

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self.slots = {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        lr_t = self.learning_rate * math.sqrt(1 - self.beta_2 ** t) / (1 - self.beta_1 ** t)
        with torch.no_grad():
            for grad, variable in grads_and_vars:
                if grad is None:
                    continue
                grad = grad.as_subclass(torch.Tensor)
                m, v = self.slots.setdefault(id(variable), (torch.zeros_like(grad), torch.zeros_like(grad)))
                m.mul_(self.beta_1).add_(grad, alpha=1 - self.beta_1)
                v.mul_(self.beta_2).addcmul_(grad, grad, value=1 - self.beta_2)
                variable.as_subclass(torch.Tensor).sub_(lr_t * m / (v.sqrt() + self.epsilon))


class Mean:
    def __init__(self):
        self.total, self.count = 0.0, 0

    def update_state(self, value):
        self.total += float(value)
        self.count += 1

    def result(self):
        return self.total / max(self.count, 1)


class SparseCategoricalAccuracy:
    def __init__(self):
        self.correct, self.count = 0, 0

    def update_state(self, y_true, y_pred):
        hits = (torch.argmax(_t(y_pred), dim=-1) == _t(y_true).long())
        self.correct += int(hits.sum())
        self.count += hits.numel()

    def result(self):
        return self.correct / max(self.count, 1)


# ---------------------------------------------------------------------------
# tf.train / tf.summary: the train loop's bookkeeping, recorded instead of written
# ---------------------------------------------------------------------------

SCALARS = []     # (tag, value, step) in the order the reference's train loop wrote them


class _Writer:
    @contextlib.contextmanager
    def as_default(self):
        yield self


class Checkpoint:
    def __init__(self, **objects):
        self.__dict__.update(objects)

    def restore(self, path):
        return self

    def expect_partial(self):
        return self


class CheckpointManager:
    def __init__(self, checkpoint, directory, max_to_keep=None):
        self.latest_checkpoint = None
        self.saves = 0

    def save(self):
        self.saves += 1
        return 'ckpt-%d' % self.saves


def install():
    '''Registers the stand-in as ``tensorflow`` (and the ``tensorflow.keras`` submodules transformer.py imports).'''

    tf = types.ModuleType('tensorflow')
    tf.__shim__ = True
    for name in ('float32', 'float64', 'int32', 'int64', 'bool'):
        setattr(tf, name, DType(name))
    tf.newaxis = None
    tf.Tensor = Tensor
    tf.cast, tf.shape, tf.reshape, tf.transpose, tf.matmul, tf.range, tf.ones = cast, shape, reshape, transpose, matmul, tf_range, ones
    tf.split, tf.unstack, tf.concat, tf.stack, tf.gather, tf.pad = split, unstack, concat, stack, gather, pad
    tf.tanh = lambda x: torch.tanh(_t(x))
    tf.pow = lambda x, y: torch.pow(_t(x), y)
    tf.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
    tf.equal = lambda a, b: _t(a) == (b if not isinstance(b, torch.Tensor) else _t(b))
    tf.reduce_mean, tf.argmax, tf.is_tensor = reduce_mean, argmax, is_tensor
    tf.zeros_initializer = zeros_initializer
    tf.Variable, tf.GradientTape = Variable, GradientTape
    tf.matrix_band_part = band_part

    tf.math = types.ModuleType('tensorflow.math')
    tf.math.equal = tf.equal
    tf.math.rsqrt = lambda x: torch.rsqrt(_t(x))
    tf.linalg = types.ModuleType('tensorflow.linalg')
    tf.linalg.band_part = band_part
    tf.nn = types.ModuleType('tensorflow.nn')
    tf.nn.softmax = softmax

    tf.summary = types.ModuleType('tensorflow.summary')
    tf.summary.create_file_writer = lambda path: _Writer()
    tf.summary.scalar = lambda tag, value, step=None: SCALARS.append((tag, float(value), int(step)))
    tf.train = types.ModuleType('tensorflow.train')
    tf.train.Checkpoint, tf.train.CheckpointManager = Checkpoint, CheckpointManager
    tf.train.latest_checkpoint = lambda directory: None

    # only touched as a default argument when composer/models/__init__.py is imported (:239)
    tf.data = types.SimpleNamespace(experimental=types.SimpleNamespace(AUTOTUNE=-1), Dataset=None)

    keras = types.ModuleType('tensorflow.keras')
    keras.Model = Model
    layers = types.ModuleType('tensorflow.keras.layers')
    layers.Layer, layers.Dropout, layers.LayerNormalization, layers.Embedding = Layer, Dropout, LayerNormalization, Embedding
    optimizers = types.ModuleType('tensorflow.keras.optimizers')
    optimizers.Adam = Adam
    losses = types.ModuleType('tensorflow.keras.losses')
    losses.SparseCategoricalCrossentropy = SparseCategoricalCrossentropy
    initializers = types.ModuleType('tensorflow.keras.initializers')
    initializers.TruncatedNormal, initializers.GlorotUniform = TruncatedNormal, GlorotUniform
    metrics = types.ModuleType('tensorflow.keras.metrics')
    metrics.Mean, metrics.SparseCategoricalAccuracy = Mean, SparseCategoricalAccuracy
    keras.layers, keras.optimizers, keras.losses, keras.initializers, keras.metrics = layers, optimizers, losses, initializers, metrics
    tf.keras = keras

    sys.modules['tensorflow'] = tf
    for name, module in (('keras', keras), ('keras.layers', layers), ('keras.optimizers', optimizers),
                         ('keras.losses', losses), ('keras.initializers', initializers), ('keras.metrics', metrics),
                         ('math', tf.math), ('linalg', tf.linalg), ('nn', tf.nn), ('summary', tf.summary),
                         ('train', tf.train)):
        sys.modules['tensorflow.' + name] = module
    return tf

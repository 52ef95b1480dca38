'''
CPU restatement of the reference Transformer (TEST INFRASTRUCTURE ONLY).

This file restates, operation by operation, the arithmetic of the reference
model in ``composer/models/transformer.py`` so that the CUDA path can be
checked against it.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py``.  Nothing under ``composer_b200/`` may import it.

PARITY STATUS: pinned to the reference's own model code, not to TensorFlow's kernels.  The reference's
arithmetic lives in TensorFlow, an un-vendored, un-pinned dependency (``environment.yml:13``:
``tensorflow-gpu``) that is not installable in this image, and the reference's own tests
(``tests/test_sequences.py``) never touch the model.  What pins this file instead:
``tests/golden/model_golden.npz`` holds outputs of the reference's UNMODIFIED
``composer/models/transformer.py`` (logits, presents, greedy ``past=`` decoding, and the losses, accuracies,
gradients and final variables of its own ``train()`` loop) executed with ``tests/golden/tf_shim.py`` standing in
for the ~45 TensorFlow entry points it calls (each restated there in one line with TF's published semantics,
fp64); ``tests/test_oracle.py::test_oracle_matches_reference_transformer`` requires agreement to 1e-11 (logits,
loss) / fp32 storage precision (gradients, variables), and ``test_model_golden_is_what_the_reference_computes``
re-runs the reference live where ``/root/reference`` exists.  The composition of the model is therefore the
reference's code; the per-op arithmetic (softmax, LayerNormalization, Adam, ...) remains a restatement of the
dependency's published algorithm.  ``tools/dump_tf_reference.py`` produces goldens from real TensorFlow wherever
it exists (``tests/golden/tf_reference.npz``, test skipped until then).

Every function cites the reference lines it follows (paths relative to the
reference root).  torch (CPU) is used as the array library so that
``tape.gradient`` (transformer.py:920) can be restated with autograd; pass
``dtype=torch.float64`` for checks and ``torch.float32`` for timing.

Parameters are a flat ``dict`` keyed by the Keras variable names:

    wte/weight [V, E]                      transformer.py:116, 667
    wpe/embeddings [window, E]             transformer.py:675-679
    h_{i}/ln_1/gamma, beta [E]             i = 1..L, transformer.py:551, 692
    h_{i}/attn/c_attn/weight [E, 3E], bias [1, 3E]    transformer.py:257-262, 189-190
    h_{i}/attn/c_proj/weight [E, E],  bias [1, E]     transformer.py:264-269
    h_{i}/ln_2/gamma, beta [E]             transformer.py:563
    h_{i}/mlp/c_fc/weight [E, 4E],  bias [1, 4E]      transformer.py:482-487
    h_{i}/mlp/c_proj/weight [4E, E], bias [1, E]      transformer.py:489-494
    ln_f/gamma, beta [E]                   transformer.py:694
'''

import math
from collections import OrderedDict

import numpy as np
import torch


class OracleConfig:
    '''Hyperparameters, named as in ``default_config.yml:32-45`` / ``Transformer.__init__`` (transformer.py:610-614).'''

    def __init__(self, vocab_size=390, embedding_size=256, window_size=1024, decoder_layers_count=8,
                 attention_head_count=16, attention_dropout_rate=0.1, residual_dropout_rate=0.1,
                 layer_normalization_epsilon=1e-5, scale_attention=True, use_layer_normalization=True,
                 initializer_mean=0.0, initializer_stddev=0.02):
        assert embedding_size % attention_head_count == 0  # transformer.py:255
        self.vocab_size = vocab_size
        self.embedding_size = embedding_size
        self.window_size = window_size
        self.decoder_layers_count = decoder_layers_count
        self.attention_head_count = attention_head_count
        self.attention_dropout_rate = attention_dropout_rate
        self.residual_dropout_rate = residual_dropout_rate
        self.layer_normalization_epsilon = layer_normalization_epsilon
        self.scale_attention = scale_attention
        self.use_layer_normalization = use_layer_normalization
        self.initializer_mean = initializer_mean
        self.initializer_stddev = initializer_stddev


def parameter_shapes(cfg):
    '''Ordered name -> shape map in Keras creation order (see module docstring).'''

    E, V, W = cfg.embedding_size, cfg.vocab_size, cfg.window_size
    shapes = OrderedDict()
    shapes['wte/weight'] = (V, E)
    shapes['wpe/embeddings'] = (W, E)
    for i in range(1, cfg.decoder_layers_count + 1):
        p = 'h_%d/' % i
        shapes[p + 'ln_1/gamma'] = (E,)
        shapes[p + 'ln_1/beta'] = (E,)
        shapes[p + 'attn/c_attn/weight'] = (E, 3 * E)
        shapes[p + 'attn/c_attn/bias'] = (1, 3 * E)
        shapes[p + 'attn/c_proj/weight'] = (E, E)
        shapes[p + 'attn/c_proj/bias'] = (1, E)
        shapes[p + 'ln_2/gamma'] = (E,)
        shapes[p + 'ln_2/beta'] = (E,)
        shapes[p + 'mlp/c_fc/weight'] = (E, 4 * E)
        shapes[p + 'mlp/c_fc/bias'] = (1, 4 * E)
        shapes[p + 'mlp/c_proj/weight'] = (4 * E, E)
        shapes[p + 'mlp/c_proj/bias'] = (1, E)
    shapes['ln_f/gamma'] = (E,)
    shapes['ln_f/beta'] = (E,)
    return shapes


def init_parameters(cfg, seed=0):
    '''
    Random initial parameters with the reference's distributions
    (transformer.py:115, 188-190, 670-673; Keras LayerNormalization defaults):
    truncated normal(mean, stddev) re-drawn outside two standard deviations
    for every weight / embedding, zeros for biases and beta, ones for gamma.
    TF's random stream cannot be reproduced, only its distribution.
    Returns float32 numpy arrays.
    '''

    rng = np.random.default_rng(seed)

    def truncated_normal(shape):
        out = rng.standard_normal(shape)
        bad = np.abs(out) > 2.0
        while bad.any():
            out[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(out) > 2.0
        return (cfg.initializer_mean + cfg.initializer_stddev * out).astype(np.float32)

    params = OrderedDict()
    for name, shape in parameter_shapes(cfg).items():
        if name.endswith('/gamma'):
            params[name] = np.ones(shape, dtype=np.float32)
        elif name.endswith('/beta') or name.endswith('/bias'):
            params[name] = np.zeros(shape, dtype=np.float32)
        else:
            params[name] = truncated_normal(shape)
    return params


def to_torch(params, dtype=torch.float64, requires_grad=False):
    out = OrderedDict()
    for name, value in params.items():
        tensor = torch.as_tensor(np.asarray(value)).to(dtype).clone()
        tensor.requires_grad_(requires_grad)
        out[name] = tensor
    return out


# --------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------

def gelu(x):
    '''transformer.py:35-40 — tanh form of GELU.'''

    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def layer_normalization(x, gamma, beta, epsilon):
    '''
    Keras ``LayerNormalization(epsilon=...)`` over the last axis
    (transformer.py:551, 563, 694): biased variance, epsilon inside the rsqrt.
    '''

    mean = x.mean(dim=-1, keepdim=True)
    variance = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(variance + epsilon) * gamma + beta


def dropout(x, rate, keep_mask):
    '''
    Keras ``Dropout(rate)`` in training mode: kept elements scaled by
    1 / (1 - rate).  ``keep_mask`` None means inference (identity), which is
    also what ``training=False`` gives in the reference.  The mask is an input
    because TF's random stream is not reproducible; the CUDA path exports the
    masks it used so both sides apply the same ones.
    '''

    if keep_mask is None or rate == 0.0:
        return x
    return x * keep_mask.to(x.dtype) / (1.0 - rate)


def embedding_lookup(weight, ids):
    '''transformer.py:137-138 (``tf.gather``) and the Keras ``Embedding`` at :675-679.'''

    return weight[ids]


def embedding_linear(weight, hidden):
    '''transformer.py:139-144 — tied output projection ``h · wteᵀ`` (no bias).'''

    flat = hidden.reshape(-1, weight.shape[1])
    logits = flat @ weight.t()
    return logits.reshape(*hidden.shape[:-1], weight.shape[0])


def conv1d(x, weight, bias):
    '''transformer.py:194-209 — ``reshape([-1, in]) @ W[in, out] + b[1, out]``.'''

    batch, sequence = x.shape[:2]
    flat = x.reshape(-1, weight.shape[0])
    flat = flat @ weight + bias
    return flat.reshape(batch, sequence, weight.shape[1])


def causal_attention_mask(nd, ns, dtype):
    '''transformer.py:290-301 — ones where ``i >= j - ns + nd`` (lower-right anchored).'''

    i = torch.arange(nd)[:, None]
    j = torch.arange(ns)
    return (i >= j - ns + nd).to(dtype)


def split_heads(x, head_count):
    '''transformer.py:385-395 — [B, T, E] -> [B, H, T, E/H].'''

    batch, sequence, features = x.shape
    return x.reshape(batch, sequence, head_count, features // head_count).permute(0, 2, 1, 3)


def merge_heads(x):
    '''transformer.py:373-383 — [B, H, T, d] -> [B, T, H*d].'''

    batch, heads, sequence, depth = x.shape
    return x.permute(0, 2, 1, 3).reshape(batch, sequence, heads * depth)


def multihead_attention(q, k, v, cfg, attention_keep_mask=None):
    '''
    transformer.py:331-371: ``w = q kᵀ``; scale by rsqrt(d_k) (before the
    mask, :345-348); ``w = w*b - 1e4*(1-b)`` (:351-354); softmax over the
    last axis (:360); dropout on the probabilities (:361); ``w v`` (:367).
    '''

    w = q @ k.transpose(-1, -2)
    if cfg.scale_attention:
        w = w * (1.0 / math.sqrt(k.shape[-1]))

    nd, ns = w.shape[-2:]
    b = causal_attention_mask(nd, ns, w.dtype).reshape(1, 1, nd, ns)
    w = w * b - 1e4 * (1 - b)
    w = torch.softmax(w, dim=-1)
    w = dropout(w, cfg.attention_dropout_rate, attention_keep_mask)
    return w @ v


def attention(x, params, prefix, cfg, layer_past=None, masks=None):
    '''
    transformer.py:397-448.  Returns (output [B, T, E], present [2, B, H, t+T, d]).
    ``masks`` may hold ``'attn_w'`` ([B, H, T, t+T]) and ``'attn_resid'`` ([B, T, E]).
    '''

    masks = masks or {}
    qkv = conv1d(x, params[prefix + 'attn/c_attn/weight'], params[prefix + 'attn/c_attn/bias'])
    query, key, value = torch.split(qkv, cfg.embedding_size, dim=2)
    query = split_heads(query, cfg.attention_head_count)
    key = split_heads(key, cfg.attention_head_count)
    value = split_heads(value, cfg.attention_head_count)

    if layer_past is not None:
        past_key, past_value = layer_past[0], layer_past[1]
        key = torch.cat([past_key, key], dim=-2)
        value = torch.cat([past_value, value], dim=-2)

    present = torch.stack([key, value], dim=0)
    a = multihead_attention(query, key, value, cfg, masks.get('attn_w'))
    a = merge_heads(a)
    a = conv1d(a, params[prefix + 'attn/c_proj/weight'], params[prefix + 'attn/c_proj/bias'])
    a = dropout(a, cfg.residual_dropout_rate, masks.get('attn_resid'))
    return a, present


def multilayer_perceptron(x, params, prefix, cfg, masks=None):
    '''transformer.py:498-507 — c_fc, gelu, c_proj, dropout.'''

    masks = masks or {}
    h = gelu(conv1d(x, params[prefix + 'mlp/c_fc/weight'], params[prefix + 'mlp/c_fc/bias']))
    h = conv1d(h, params[prefix + 'mlp/c_proj/weight'], params[prefix + 'mlp/c_proj/bias'])
    return dropout(h, cfg.residual_dropout_rate, masks.get('mlp'))


def decoder_block(x, params, index, cfg, layer_past=None, masks=None):
    '''
    transformer.py:574-597.  Note that ``ln_1`` overwrites ``x`` (:583-584), so
    the attention skip connection adds to the *normalised* activations:

        x1 = LN1(x);  x2 = x1 + Attn(x1);  x3 = x2 + MLP(LN2(x2))
    '''

    prefix = 'h_%d/' % index
    eps = cfg.layer_normalization_epsilon
    if cfg.use_layer_normalization:
        x = layer_normalization(x, params[prefix + 'ln_1/gamma'], params[prefix + 'ln_1/beta'], eps)

    a, present = attention(x, params, prefix, cfg, layer_past, masks)
    x = x + a

    m = x
    if cfg.use_layer_normalization:
        m = layer_normalization(x, params[prefix + 'ln_2/gamma'], params[prefix + 'ln_2/beta'], eps)

    m = multilayer_perceptron(m, params, prefix, cfg, masks)
    x = x + m
    return x, present


def transformer_call(params, ids, cfg, past=None, dropout_masks=None, return_hidden=False):
    '''
    ``Transformer.call`` (transformer.py:696-833) for the arguments this path
    uses: integer ``ids`` [B, T], optional ``past`` (tuple of L presents), no
    attention/head masks, no token types.  ``dropout_masks`` None is
    ``training=False``; otherwise a dict with ``'embd'`` and ``(layer, site)``
    keys, site in {'attn_w', 'attn_resid', 'mlp'}, layer 1-based.
    Returns (logits [B, T, V], presents).
    '''

    ids = torch.as_tensor(np.asarray(ids)).long()
    if past is not None:
        ids = ids[:, -1:]                                               # :735-737
        past_length = past[0][0].shape[-2]                              # :765
    else:
        past_length = 0                                                 # :762
        past = [None] * cfg.decoder_layers_count

    sequence = ids.shape[-1]
    position_ids = torch.arange(past_length, sequence + past_length)    # :770
    if int(position_ids[-1]) >= cfg.window_size:
        # TF-CPU's gather raises on out-of-range indices into wpe (window_size rows).
        raise IndexError('position %d is outside wpe (window_size=%d)' % (int(position_ids[-1]), cfg.window_size))

    hidden = embedding_lookup(params['wte/weight'], ids) \
        + embedding_lookup(params['wpe/embeddings'], position_ids)[None]    # :783-793
    masks = dropout_masks or {}
    hidden = dropout(hidden, cfg.residual_dropout_rate, masks.get('embd'))  # :794

    presents = []
    for layer in range(1, cfg.decoder_layers_count + 1):                    # :800-809
        layer_masks = {site: masks[(layer, site)] for site in ('attn_w', 'attn_resid', 'mlp')
                       if (layer, site) in masks}
        hidden, present = decoder_block(hidden, params, layer, cfg, past[layer - 1], layer_masks)
        presents.append(present)

    hidden = layer_normalization(hidden, params['ln_f/gamma'], params['ln_f/beta'],
                                 cfg.layer_normalization_epsilon)           # :811
    logits = embedding_linear(params['wte/weight'], hidden)                 # :818
    if return_hidden:
        return logits, tuple(presents), hidden
    return logits, tuple(presents)


# --------------------------------------------------------------------------
# Loss, metrics, optimizer, sampling
# --------------------------------------------------------------------------

def sparse_categorical_crossentropy(labels, logits):
    '''
    ``SparseCategoricalCrossentropy(from_logits=True)`` with Keras' default
    reduction (transformer.py:888, 918): mean over every position of
    ``logsumexp(z) - z[y]``.
    '''

    labels = torch.as_tensor(np.asarray(labels)).long()
    flat = logits.reshape(-1, logits.shape[-1])
    picked = flat.gather(1, labels.reshape(-1, 1)).squeeze(1)
    return (torch.logsumexp(flat, dim=-1) - picked).mean()


def batch_accuracy(labels, logits):
    '''transformer.py:924-926 — mean(argmax(logits) == y).'''

    labels = torch.as_tensor(np.asarray(labels)).long()
    return (logits.argmax(dim=-1) == labels).to(torch.float64).mean()


def loss_and_gradients(params_np, ids, labels, cfg, dtype=torch.float64, dropout_masks=None):
    '''
    One ``GradientTape`` evaluation (transformer.py:916-920): returns
    (loss, accuracy, logits, {name: gradient}) as numpy values.
    '''

    params = to_torch(params_np, dtype, requires_grad=True)
    logits, _ = transformer_call(params, ids, cfg, dropout_masks=dropout_masks)
    loss = sparse_categorical_crossentropy(labels, logits)
    loss.backward()
    # (a variable the configuration does not use, e.g. ln_1 / ln_2 without LayerNorm, has no gradient: zeros)
    grads = OrderedDict((name, (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy())
                        for name, p in params.items())
    return float(loss.detach()), float(batch_accuracy(labels, logits.detach())), logits.detach().numpy(), grads


class AdamState:
    '''
    TF-2 Keras ``Adam(learning_rate)`` with its defaults (transformer.py:887):
    beta_1 0.9, beta_2 0.999, epsilon 1e-7, no amsgrad.  Per step t = 1, 2, ...

        m <- b1 m + (1-b1) g ;  v <- b2 v + (1-b2) g^2
        lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
        theta <- theta - lr_t * m / (sqrt(v) + eps)

    i.e. epsilon is added to sqrt(v), not to the bias-corrected sqrt(v-hat)
    (this is what distinguishes it from ``torch.optim.Adam``).
    '''

    def __init__(self, params, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        self.m = OrderedDict((k, np.zeros_like(v, dtype=np.float64)) for k, v in params.items())
        self.v = OrderedDict((k, np.zeros_like(v, dtype=np.float64)) for k, v in params.items())

    def apply(self, params, grads):
        '''Updates ``params`` (dict of numpy arrays) in place, returns it.'''

        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        for name in params:
            g = np.asarray(grads[name], dtype=np.float64)
            self.m[name] = self.b1 * self.m[name] + (1 - self.b1) * g
            self.v[name] = self.b2 * self.v[name] + (1 - self.b2) * g * g
            update = lr_t * self.m[name] / (np.sqrt(self.v[name]) + self.eps)
            params[name] = (np.asarray(params[name], dtype=np.float64) - update).astype(params[name].dtype)
        return params


def next_token_distribution(logits_last, temperature):
    '''
    cli.py:670-673 — ``tf.random.categorical(logits / temperature, 1)`` draws
    from softmax(logits / temperature); this returns that distribution for
    the last position (float64 numpy, rows sum to 1).
    '''

    z = np.asarray(logits_last, dtype=np.float64) / temperature
    z = z - z.max(axis=-1, keepdims=True)
    p = np.exp(z)
    return p / p.sum(axis=-1, keepdims=True)


def generate(params_np, prompt_ids, length, cfg, temperature=1.0, dtype=torch.float64,
             greedy=False, uniforms=None, use_cache=True):
    '''
    Autoregressive decoding with the model's ``past=`` interface
    (transformer.py:735-770, 423-426): the prompt is run once, then one token
    per step with the cached keys/values.  ``use_cache=False`` recomputes the
    whole sequence every step instead (must give the same logits).

    Note: the reference CLI loop (cli.py:663-676) calls ``model(x)`` without
    ``past`` and feeds back only the last id, i.e. it discards the context;
    the north-star asks for KV-cache decoding, which is what this restates.

    Token choice: ``greedy`` takes argmax; otherwise inverse-CDF sampling from
    softmax(logits / temperature) with ``uniforms[b, step]`` in [0, 1) (the
    CUDA sampler can be driven with the same uniforms).
    Returns (ids [B, length], per-step last-position logits [B, length, V]).
    '''

    params = to_torch(params_np, dtype)
    ids = torch.as_tensor(np.asarray(prompt_ids)).long()
    batch = ids.shape[0]
    out_ids = np.zeros((batch, length), dtype=np.int64)
    out_logits = np.zeros((batch, length, cfg.vocab_size), dtype=np.float64)

    with torch.no_grad():
        past = None
        context = ids
        for step in range(length):
            if use_cache:
                logits, past = transformer_call(params, context, cfg, past=past)
            else:
                logits, _ = transformer_call(params, context, cfg)
            last = logits[:, -1, :].double().numpy()
            out_logits[:, step] = last
            if greedy:
                chosen = last.argmax(axis=-1)
            else:
                cdf = np.cumsum(next_token_distribution(last, temperature), axis=-1)
                u = uniforms[:, step:step + 1]
                chosen = np.minimum((cdf <= u).sum(axis=-1), cfg.vocab_size - 1)
            out_ids[:, step] = chosen
            nxt = torch.as_tensor(chosen).long()[:, None]
            context = nxt if use_cache else torch.cat([context, nxt], dim=1)

    return out_ids, out_logits

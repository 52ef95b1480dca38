'''
The Transformer decoder of Composer, driven on a B200 through the C ABI of
``libcomposer_b200`` (``include/composer_b200.h``).

This class mirrors the reference's ``composer.models.Transformer``
(composer/models/transformer.py:599-960): same constructor arguments, the same
``__call__`` / ``train`` / ``compile`` / ``build`` / ``load_from_checkpoint``
surface used by the CLI (composer/cli.py:579-589, 635-676), the same Keras
variable names for its weights.  PyTorch only owns device memory, streams and
the NCCL process group; every arithmetic operation is a hand-written sm_100a
kernel.  There is no CPU path: constructing the model without a CUDA device or
without the built library raises.
'''

import ctypes
import json
import logging
import math
import os
import time
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch

from composer_b200 import ModelSaveFrequencyMode, _lib, parallel, tf_checkpoint
from composer_b200.models.base import BaseModel

CHECKPOINT_INDEX = 'checkpoint.json'


def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Presents:
    '''
    The ``presents`` / ``past`` of the model call (transformer.py:423-432, 800-809): behaves like the reference's tuple
    of ``decoder_layers_count`` tensors ``[2, batch, heads, t, d_h]`` (index it, iterate it, ``len``), and is one
    bf16 CUDA tensor ``[layers, 2, batch, heads, capacity, d_h]`` underneath, so that a decode step appends the new
    key / value rows in place instead of re-concatenating the whole cache.
    '''

    def __init__(self, owner, storage, length):
        self.owner, self.storage, self.length = owner, storage, int(length)

    @property
    def capacity(self):
        return self.storage.shape[4]

    @property
    def batch(self):
        return self.storage.shape[2]

    @classmethod
    def allocate(cls, model, batch, length):
        # room for 64 .. 127 more positions (a step that finds the cache full moves it to a larger one), never more
        # than wpe can address
        capacity = max(min(model.window_size, (length + 127) // 64 * 64), length)
        heads = model.attention_head_count
        storage = torch.empty((model.decoder_layers_count, 2, batch, heads, capacity, model.embedding_size // heads),
                              dtype=torch.bfloat16, device=model.device)
        return cls(model, storage, length)

    @classmethod
    def from_tensors(cls, model, past):
        layers = [torch.as_tensor(p) for p in past]
        if len(layers) != model.decoder_layers_count:
            raise ValueError('past has %d layers, the model %d' % (len(layers), model.decoder_layers_count))
        length = layers[0].shape[-2]
        presents = cls.allocate(model, layers[0].shape[1], length + 1)
        for index, layer in enumerate(layers):
            presents.storage[index, :, :, :, :length] = layer.to(device=model.device, dtype=torch.bfloat16)
        presents.length = length
        return presents

    def __len__(self):
        return self.storage.shape[0]

    def __getitem__(self, layer):
        return self.storage[layer, :, :, :, :self.length]

    def __iter__(self):
        return (self[layer] for layer in range(len(self)))


class Transformer(BaseModel):
    '''
    A Transformer-decoder model that generates music as a sequence of MIDI-like
    events (see :mod:`composer_b200.dataset.sequence`).

    Constructor arguments are those of the reference (transformer.py:610-614),
    in the same order, so that ``cli.create_model`` (cli.py:123-132) maps
    ``transformer.model.*`` positionally.
    '''

    def __init__(self, vocab_size, embedding_size, window_size, decoder_layers_count,
                 attention_head_count, use_relative_attention=False, initializer_mean=0,
                 initializer_stddev=0.02, attention_dropout_rate=0.1, residual_dropout_rate=0.1,
                 layer_normalization_epsilon=1e-5, scale=True, use_layer_normalization=True,
                 output_hidden_states=False, output_attention_weights=False, device=None, seed=0,
                 process_group=None):
        if use_relative_attention:
            # The reference's relative attention cannot run (``self.depth`` is undefined, transformer.py:281-286).
            raise NotImplementedError('use_relative_attention is broken in the reference and not provided here.')
        if output_hidden_states or output_attention_weights:
            raise NotImplementedError('output_hidden_states / output_attention_weights are not on the '
                                      'train/generate path and are not provided by the B200 build.')
        if not torch.cuda.is_available():
            raise RuntimeError('composer_b200.models.Transformer needs a CUDA device (B200); there is no CPU path.')

        _lib.load()
        self.vocab_size = int(vocab_size)
        self.embedding_size = int(embedding_size)
        self.window_size = int(window_size)
        self.decoder_layers_count = int(decoder_layers_count)
        self.attention_head_count = int(attention_head_count)
        self.attention_dropout_rate = float(attention_dropout_rate)
        self.residual_dropout_rate = float(residual_dropout_rate)
        self.layer_normalization_epsilon = float(layer_normalization_epsilon)
        self.scale = bool(scale)
        self.use_layer_normalization = bool(use_layer_normalization)
        self.initializer_mean = float(initializer_mean)
        self.initializer_stddev = float(initializer_stddev)
        self.seed = int(seed)
        self.process_group = process_group
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.learning_rate = 1e-3

        self._config = _lib.Config(self.vocab_size, self.embedding_size, self.window_size,
                                   self.decoder_layers_count, self.attention_head_count,
                                   self.attention_dropout_rate, self.residual_dropout_rate,
                                   self.layer_normalization_epsilon, int(self.scale),
                                   int(self.use_layer_normalization))
        handle = ctypes.c_void_p()
        _lib.call('cb200_engine_create', ctypes.byref(self._config), ctypes.byref(handle))
        self._engine = handle
        self._layout = self._query_layout()

        with torch.cuda.device(self.device):
            count = _lib.call('cb200_param_elems', ctypes.byref(self._config))
            self._params = torch.zeros(count, dtype=torch.float32, device=self.device)
            self._shadow = torch.zeros(_lib.call('cb200_shadow_elems', self._engine), dtype=torch.bfloat16,
                                       device=self.device)
        self._grads = self._adam_m = self._adam_v = None
        self._workspace = None
        self._bound = None          # (max_B, max_T, training)
        self._adam_t = 0
        self._global_step = 1       # tf.Variable(1), transformer.py:890
        self._epoch = 1
        self._loss_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._correct_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._decode_state = None
        self._step_ws = None
        self.set_weights(self.initial_weights(self.seed))

    def __del__(self):
        engine = getattr(self, '_engine', None)
        if engine:
            try:
                _lib.call('cb200_engine_destroy', engine)
            except Exception:   # interpreter shutdown
                pass
            self._engine = None

    # ------------------------------------------------------------------
    # Variables
    # ------------------------------------------------------------------
    def _query_layout(self):
        layout = OrderedDict()
        count = _lib.call('cb200_param_tensor_count', ctypes.byref(self._config))
        name = ctypes.create_string_buffer(128)
        offset, rows, cols = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32()
        for index in range(count):
            _lib.call('cb200_param_tensor_info', ctypes.byref(self._config), index, name, 128, ctypes.byref(offset),
                      ctypes.byref(rows), ctypes.byref(cols))
            layout[name.value.decode()] = (offset.value, rows.value, cols.value)
        return layout

    @property
    def variable_names(self):
        return list(self._layout)

    def _shape_of(self, name):
        _, rows, cols = self._layout[name]
        if name.endswith('/gamma') or name.endswith('/beta'):
            return (cols,)                       # Keras LayerNormalization variables are [E]
        return (rows, cols)                      # biases are [1, out] (transformer.py:189-190)

    def initial_weights(self, seed=0):
        '''
        Fresh variables with the reference's initializers (transformer.py:115, 188-190, 670-673):
        truncated normal(mean, stddev) for weights and embeddings, zeros for
        biases / beta, ones for gamma.  Returns ``{keras name: float32 array}``.
        '''

        rng = np.random.default_rng(seed)

        def truncated_normal(shape):
            out = rng.standard_normal(shape)
            bad = np.abs(out) > 2.0
            while bad.any():
                out[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(out) > 2.0
            return (self.initializer_mean + self.initializer_stddev * out).astype(np.float32)

        weights = OrderedDict()
        for name in self._layout:
            shape = self._shape_of(name)
            if name.endswith('/gamma'):
                weights[name] = np.ones(shape, dtype=np.float32)
            elif name.endswith('/beta') or name.endswith('/bias'):
                weights[name] = np.zeros(shape, dtype=np.float32)
            else:
                weights[name] = truncated_normal(shape)
        return weights

    def set_weights(self, weights):
        '''Loads ``{keras name: array}`` into the fp32 arena and refreshes the bf16 shadows.'''

        host = np.zeros(self._params.numel(), dtype=np.float32)
        for name, (offset, rows, cols) in self._layout.items():
            value = np.asarray(weights[name], dtype=np.float32)
            if value.size != rows * cols:
                raise ValueError('variable %s has %d elements, expected %d' % (name, value.size, rows * cols))
            host[offset:offset + rows * cols] = value.reshape(-1)
        self._params.copy_(torch.from_numpy(host))
        self._refresh_shadows()

    def get_weights(self):
        '''Returns ``{keras name: float32 array}`` (a host copy of the fp32 master weights).'''

        host = self._params.detach().cpu().numpy()
        return OrderedDict((name, host[offset:offset + rows * cols].reshape(self._shape_of(name)).copy())
                           for name, (offset, rows, cols) in self._layout.items())

    def get_gradients(self):
        host = self._grads.detach().cpu().numpy()
        return OrderedDict((name, host[offset:offset + rows * cols].reshape(self._shape_of(name)).copy())
                           for name, (offset, rows, cols) in self._layout.items())

    def count_params(self):
        return sum(rows * cols for _, rows, cols in self._layout.values())

    @property
    def gradient_arena(self):
        '''The flat fp32 gradient tensor (what data-parallel training all-reduces).'''

        return self._grads

    def _refresh_shadows(self):
        # The transpose job table lives in the workspace, so something must be bound; binding refreshes.
        if self._bound is None:
            self._bind(1, min(self.window_size, 64), training=False)
        else:
            with torch.cuda.device(self.device):
                _lib.call('cb200_refresh_shadows', self._engine, _stream())

    # ------------------------------------------------------------------
    # Memory binding
    # ------------------------------------------------------------------
    def _bind(self, batch, sequence, training):
        if self._bound is not None:
            max_b, max_t, bound_training = self._bound
            if batch * sequence <= max_b * max_t and (bound_training or not training):
                return
            batch_tokens = max(batch * sequence, max_b * max_t)
            training = training or bound_training
            sequence = max(sequence, max_t)
            batch = (batch_tokens + sequence - 1) // sequence

        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if training and self._grads is None:
                self._grads = torch.zeros_like(self._params)
                self._adam_m = torch.zeros_like(self._params)
                self._adam_v = torch.zeros_like(self._params)
            need = _lib.call('cb200_workspace_bytes', self._engine, batch, sequence, int(training))
            self._workspace = None
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.call('cb200_engine_bind', self._engine, _ptr(self._params), _ptr(self._grads), _ptr(self._adam_m),
                      _ptr(self._adam_v), _ptr(self._shadow), _ptr(self._workspace), need, batch, sequence,
                      int(training))
            _lib.call('cb200_refresh_shadows', self._engine, _stream())
        self._bound = (batch, sequence, training)

    def build(self, input_shape=None):
        '''Keras-compatibility no-op (cli.py:640): variables exist from construction.'''

        return self

    def compile(self, learning_rate):
        '''transformer.py:835-844 — remembers the optimizer's learning rate.'''

        self.learning_rate = float(learning_rate)
        return self

    def reset_states(self):
        self._decode_state = None

    # ------------------------------------------------------------------
    # Forward / training step
    # ------------------------------------------------------------------
    def _as_ids(self, array):
        tensor = array if isinstance(array, torch.Tensor) else torch.as_tensor(np.asarray(array))
        if tensor.dim() == 1:
            tensor = tensor[None]
        if tensor.dim() != 2:
            raise ValueError('expected integer ids of shape [batch, sequence], got %r' % (tuple(tensor.shape),))
        return tensor.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()

    def __call__(self, inputs, past=None, training=False, use_cache=True):
        '''
        ``Transformer.call`` (transformer.py:696-833) for integer ``inputs``
        [batch, sequence].  Returns ``(logits, presents)`` like the reference
        (``(logits,)`` with ``use_cache=False``): logits is a float32 CUDA
        tensor [batch, sequence, vocab] and ``presents`` a :class:`Presents`,
        a sequence of ``decoder_layers_count`` tensors [2, batch, heads, t, d_h]
        (transformer.py:430-432) that can be handed back as ``past``.

        With ``past`` only the last token of ``inputs`` is used (:735-737); it
        is embedded at position ``t`` = the length of the past (:765-770), its
        key / value rows are appended and the logits have shape [batch, 1, vocab].
        A :class:`Presents` returned by this model is extended IN PLACE (the
        reference re-concatenates the whole cache every step, :423-426): the
        object passed in and the one returned share storage, and stepping from
        the same past twice overwrites position ``t``.  Any other sequence of
        [2, batch, heads, t, d_h] tensors is copied into a fresh cache first.
        '''

        ids = self._as_ids(inputs)
        if past is None:
            batch, sequence = ids.shape
            self._bind(batch, sequence, training=False)
            logits = torch.empty((batch, sequence, self.vocab_size), dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                if not use_cache:
                    _lib.call('cb200_forward', self._engine, _ptr(ids), None, batch, sequence, 0, self.seed, 0, 0.0,
                              None, None, _ptr(logits), _stream())
                    return (logits,)
                presents = Presents.allocate(self, batch, sequence)
                _lib.call('cb200_prefill', self._engine, _ptr(ids), batch, sequence, _ptr(presents.storage),
                          presents.capacity, _ptr(logits), _stream())
            return logits, presents

        ids = ids[:, -1].contiguous()                                       # transformer.py:735-737
        batch = ids.shape[0]
        if not (isinstance(past, Presents) and past.owner is self and past.length < past.capacity):
            past = Presents.from_tensors(self, past)
        if past.batch != batch:
            raise ValueError('past holds %d sequences, inputs %d' % (past.batch, batch))
        if past.length >= self.window_size:
            # TF-CPU's gather raises on an index outside wpe (window_size rows, transformer.py:675-679, 770)
            raise IndexError('position %d is outside wpe (window_size=%d)' % (past.length, self.window_size))
        if self._bound is None:
            self._bind(1, min(self.window_size, 64), training=False)
        logits = torch.empty((batch, 1, self.vocab_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            workspace = self._step_workspace(batch)
            _lib.call('cb200_decode_step', self._engine, _ptr(past.storage), past.capacity, _ptr(workspace),
                      workspace.numel(), _ptr(ids), batch, past.length, _ptr(logits), _stream())
        past.length += 1
        return (logits, past) if use_cache else (logits,)

    def _step_workspace(self, batch):
        if self._step_ws is None or self._step_ws[0] < batch:
            need = _lib.call('cb200_decode_workspace_bytes', self._engine, batch)
            self._step_ws = (batch, torch.empty(need, dtype=torch.uint8, device=self.device))
        return self._step_ws[1]

    def forward_loss(self, x, y, training=False, step=0, return_logits=False):
        '''Forward + summed loss + correct count on the device; returns (loss_sum, correct[, logits]) tensors.'''

        ids, labels = self._as_ids(x), self._as_ids(y)
        batch, sequence = ids.shape
        self._bind(batch, sequence, training=training)
        logits = None
        if return_logits:
            logits = torch.empty((batch, sequence, self.vocab_size), dtype=torch.float32, device=self.device)
        self._loss_dev.zero_()
        self._correct_dev.zero_()
        with torch.cuda.device(self.device):
            # dropout masks differ between data-parallel ranks (everything else keyed by the seed does not)
            _lib.call('cb200_forward', self._engine, _ptr(ids), _ptr(labels), batch, sequence, int(training),
                      parallel.rank_dropout_seed(self.seed, self._rank()), step, 1.0 / (batch * sequence),
                      _ptr(self._loss_dev), _ptr(self._correct_dev), _ptr(logits), _stream())
        self._last_ids = (ids, labels)   # keep the device buffers alive until backward has run
        if return_logits:
            return self._loss_dev, self._correct_dev, logits
        return self._loss_dev, self._correct_dev

    def _rank(self):
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_rank(self.process_group)
        return 0

    def sync_replicas(self):
        '''
        Data-parallel replicas start from rank 0's variables and optimizer state (a restore or a seed that
        differs between ranks would otherwise leave them apart for good: only gradients are exchanged).
        '''

        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        if torch.distributed.get_world_size(self.process_group) == 1:
            return
        torch.distributed.broadcast(self._params, src=0, group=self.process_group)
        if self._adam_m is not None:
            torch.distributed.broadcast(self._adam_m, src=0, group=self.process_group)
            torch.distributed.broadcast(self._adam_v, src=0, group=self.process_group)
        counters = torch.tensor([self._adam_t, self._global_step, self._epoch], dtype=torch.int64, device=self.device)
        torch.distributed.broadcast(counters, src=0, group=self.process_group)
        self._adam_t, self._global_step, self._epoch = (int(v) for v in counters.tolist())
        self._refresh_shadows()

    def _gradient_buckets(self):
        return parallel.gradient_buckets(self._layout, self.decoder_layers_count)

    def backward(self, overlap_allreduce=True):
        '''
        tape.gradient (transformer.py:920).  With more than one rank the flat
        gradient arena is summed over ranks with NCCL; each bucket (one decoder
        block) is all-reduced asynchronously as soon as its backward stage has
        been enqueued, so the collective overlaps the remaining backward work.
        Returns the world size (the 1/world mean is folded into Adam).
        '''

        world = 1
        group = self.process_group
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(group)
        with torch.cuda.device(self.device):
            _lib.call('cb200_zero_grads', self._engine, _stream())
            if world == 1:
                _lib.call('cb200_backward', self._engine, -1, _stream())
                return world
            if not overlap_allreduce:
                _lib.call('cb200_backward', self._engine, -1, _stream())
                torch.distributed.all_reduce(self._grads, group=group)
                return world
            # bucket i is complete once backward stage i has been enqueued (stage 0 = ln_f; the last = embeddings)
            parallel.allreduce_buckets(
                self._grads, self._gradient_buckets(), group,
                after_bucket=lambda stage: _lib.call('cb200_backward', self._engine, stage, _stream()))
        return world

    def apply_gradients(self, learning_rate=None, world=1):
        '''optimizers.Adam.apply_gradients with TF-2 Keras defaults (transformer.py:887, 921).'''

        self._adam_t += 1
        lr = self.learning_rate if learning_rate is None else learning_rate
        with torch.cuda.device(self.device):
            _lib.call('cb200_adam_step', self._engine, lr, 0.9, 0.999, 1e-7, self._adam_t, 1.0 / world, _stream())

    def train_step(self, x, y, learning_rate=None, training=True):
        '''
        One iteration of the reference's hot loop (transformer.py:914-926):
        forward, loss, gradients, Adam.  Returns device tensors
        ``(loss_sum, correct_count)`` for this rank's batch; nothing is synchronised.
        '''

        loss, correct = self.forward_loss(x, y, training=training, step=self._adam_t)
        world = self.backward()
        self.apply_gradients(learning_rate, world)
        return loss, correct

    def evaluate(self, dataset, verbose=0):
        '''
        Mean loss and accuracy over ``dataset`` without dropout (what the reference's
        ``evaluate`` command asks of Keras, cli.py:606-615).  Returns ``(loss, accuracy)``.
        '''

        loss_total = torch.zeros(1, dtype=torch.float64, device=self.device)
        correct_total = torch.zeros(1, dtype=torch.float64, device=self.device)
        tokens = 0
        for x, y in dataset:
            loss_sum, correct = self.forward_loss(x, y, training=False)
            loss_total += loss_sum.double()
            correct_total += correct.double()
            tokens += int(np.prod(np.shape(x)))
        if tokens == 0:
            return float('nan'), float('nan')
        return float(loss_total) / tokens, float(correct_total) / tokens

    def summary_lines(self):
        '''Variable table in Keras order (the reference's ``summary`` command prints ``model.summary()``).'''

        lines = ['Model: "transformer"', '%-40s %-16s %12s' % ('Variable', 'Shape', 'Param #'), '=' * 70]
        for name in self._layout:
            shape = self._shape_of(name)
            lines.append('%-40s %-16s %12d' % (name, str(tuple(shape)), int(np.prod(shape))))
        lines.append('=' * 70)
        lines.append('Total params: {:,}'.format(self.count_params()))
        return lines

    # ------------------------------------------------------------------
    # Generation (cli.py:663-676 with the model's past= semantics)
    # ------------------------------------------------------------------
    def generate(self, prompt_ids, length, temperature=1.0, seed=None, sequence_index_base=0,
                 return_uniforms=False, return_last_logits=False):
        '''
        Autoregressively samples ``length`` new event ids after ``prompt_ids``
        ([batch, prompt_length], every row the same length) with a KV cache.
        ``temperature <= 0`` selects argmax.  Returns an int32 CUDA tensor [batch, length].

        The positional table has ``window_size`` rows (transformer.py:675-679), so one cached pass covers
        ``prompt_length + length - 1 <= window_size`` positions.  The reference's loop never runs out of
        positions because it feeds every token back at position 0 without any context (cli.py:663-676), and its
        default invocation (prompt 10, length 1024, window 1024) must keep working: a longer request is served
        in windows, each one re-primed with the last ``window_size // 2`` tokens at positions 0.. (the usual
        sliding-window continuation).  The diagnostics outputs are only defined for a single window.
        '''

        prompt = self._as_ids(prompt_ids)
        batch, prompt_length = prompt.shape
        seed = self.seed if seed is None else int(seed)
        if prompt_length > self.window_size:
            raise ValueError('the prompt has %d tokens but position embeddings exist only for window_size = %d'
                             % (prompt_length, self.window_size))
        if prompt_length - 1 + length <= self.window_size:
            return self._generate_window(prompt, length, temperature, seed, sequence_index_base, return_uniforms,
                                         return_last_logits)
        if return_uniforms or return_last_logits:
            raise ValueError('prompt_length + length - 1 = %d exceeds window_size = %d: uniforms / logits are only '
                             'returned for a generation that fits one window'
                             % (prompt_length - 1 + length, self.window_size))
        logging.warning('Generating %d events after a %d-event prompt needs more than window_size = %d positions: '
                        'continuing in windows re-primed with the last %d events.'
                        % (length, prompt_length, self.window_size, max(1, self.window_size // 2)))
        pieces, context, remaining, window = [], prompt, length, 0
        while remaining > 0:
            count = min(remaining, self.window_size - context.shape[1] + 1)
            ids = self._generate_window(context, count, temperature, seed + window, sequence_index_base, False, False)
            pieces.append(ids)
            remaining -= count
            context = torch.cat([context, ids], dim=1)[:, -max(1, self.window_size // 2):].contiguous()
            window += 1
        return torch.cat(pieces, dim=1)

    def _generate_window(self, prompt, length, temperature, seed, sequence_index_base, return_uniforms,
                         return_last_logits):
        '''One cached pass of ``cb200_generate``: prompt_length + length - 1 <= window_size positions.'''

        batch, prompt_length = prompt.shape
        steps = prompt_length - 1 + length
        if prompt_length > 1:
            # the prompt's k, v rows come from one batched forward pass: its workspace has to hold batch x (prompt - 1) tokens
            bound = self._bound
            if bound is None or bound[0] * bound[1] < batch * (prompt_length - 1):
                self._bind(batch, prompt_length - 1, training=False)
        if self._bound is None:
            self._bind(1, min(self.window_size, 64), training=False)
        with torch.cuda.device(self.device):
            t_max = (steps + 63) // 64 * 64
            key = (batch, t_max)
            if self._decode_state is None or self._decode_state[0] != key:
                self._decode_state = None
                cache = torch.empty(_lib.call('cb200_kv_cache_elems', self._engine, batch, t_max),
                                    dtype=torch.bfloat16, device=self.device)
                need = _lib.call('cb200_decode_workspace_bytes', self._engine, batch)
                workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
                self._decode_state = (key, cache, workspace)
            _, cache, workspace = self._decode_state
            out = torch.empty((batch, length), dtype=torch.int32, device=self.device)
            uniforms = torch.empty((batch, steps), dtype=torch.float32, device=self.device) if return_uniforms else None
            last = torch.empty((batch, self.vocab_size), dtype=torch.float32, device=self.device) \
                if return_last_logits else None
            _lib.call('cb200_generate', self._engine, _ptr(cache), t_max, _ptr(workspace), workspace.numel(),
                      _ptr(prompt), batch, prompt_length, length, float(temperature), seed % (1 << 64),
                      int(sequence_index_base), _ptr(out), _ptr(uniforms), _ptr(last), _stream())
        extras = [t for t in (uniforms, last) if t is not None]
        return (out, *extras) if extras else out

    # ------------------------------------------------------------------
    # Checkpoints (own format; TF object-graph checkpoints need TensorFlow)
    # ------------------------------------------------------------------
    def save_checkpoint(self, logdir, max_checkpoints=1):
        '''
        Writes ``ckpt-<step>.npz`` (variables under their Keras names, Adam
        slots, step / epoch counters) and rotates old files like
        ``tf.train.CheckpointManager(max_to_keep)`` (transformer.py:891).
        '''

        logdir = Path(logdir)
        logdir.mkdir(parents=True, exist_ok=True)
        arrays = {'variables/' + k: v for k, v in self.get_weights().items()}
        if self._adam_m is not None:
            arrays['optimizer/m'] = self._adam_m.detach().cpu().numpy()
            arrays['optimizer/v'] = self._adam_v.detach().cpu().numpy()
        arrays['optimizer/iterations'] = np.asarray(self._adam_t, dtype=np.int64)
        arrays['step'] = np.asarray(self._global_step, dtype=np.int64)
        arrays['epoch'] = np.asarray(self._epoch, dtype=np.int64)
        index_path = logdir / CHECKPOINT_INDEX
        index = {'all': [], 'latest': None}
        if index_path.exists():
            index = json.loads(index_path.read_text())
        number = (max([int(Path(p).stem.split('-')[-1]) for p in index['all']] + [0])) + 1
        path = logdir / ('ckpt-%d.npz' % number)
        np.savez(path, **arrays)
        index['all'].append(path.name)
        index['latest'] = path.name
        while max_checkpoints and len(index['all']) > max_checkpoints:
            stale = logdir / index['all'].pop(0)
            if stale.exists():
                stale.unlink()
        index_path.write_text(json.dumps(index))
        return str(path)

    @staticmethod
    def latest_checkpoint(restoredir):
        '''
        The newest checkpoint of ``restoredir``: this package's ``ckpt-N.npz`` (via ``checkpoint.json``) or, in a
        log directory written by the reference itself, the TensorFlow checkpoint prefix its ``checkpoint`` state file
        names (``tf.train.latest_checkpoint``, models/__init__.py:79).
        '''

        index_path = Path(restoredir) / CHECKPOINT_INDEX
        if index_path.exists():
            latest = json.loads(index_path.read_text()).get('latest')
            return str(Path(restoredir) / latest) if latest else None
        return tf_checkpoint.latest_checkpoint(str(restoredir))

    def _restore_tf(self, prefix, with_optimizer):
        '''A checkpoint the reference wrote (``tf.train.Checkpoint(step, epoch, optimizer, model)``, transformer.py:890).'''

        bundle = tf_checkpoint.read_bundle(prefix)
        shapes = {name: self._shape_of(name) for name in self._layout}
        weights, adam_m, adam_v, counters = tf_checkpoint.to_arrays(bundle, list(self._layout), shapes)
        self.set_weights(weights)
        self._global_step = counters.get('step', 1)
        self._epoch = counters.get('epoch', 1)
        if with_optimizer and len(adam_m) == len(self._layout) and len(adam_v) == len(self._layout):
            self._bind(1, min(self.window_size, 64), training=True)
            for slots, arena in ((adam_m, self._adam_m), (adam_v, self._adam_v)):
                host = np.zeros(arena.numel(), dtype=np.float32)
                for name, (offset, rows, cols) in self._layout.items():
                    host[offset:offset + rows * cols] = slots[name].reshape(-1)
                arena.copy_(torch.from_numpy(host))
            self._adam_t = counters.get('iterations', 0)

    def export_tf_checkpoint(self, logdir, number=1):
        '''
        Writes the variables (and Adam slots / counters when they exist) as ``<logdir>/ckpt-<number>`` in
        TensorFlow's checkpoint format under the reference's object-graph names, plus the ``checkpoint`` state
        file, so that a name-based restore on the reference side (``tf.train.load_checkpoint`` /
        ``load_variable``) finds every tensor where its own ``CheckpointManager.save`` puts it.
        '''

        logdir = Path(logdir)
        logdir.mkdir(parents=True, exist_ok=True)
        weights = self.get_weights()
        adam_m = adam_v = None
        if self._adam_m is not None:
            m_host, v_host = self._adam_m.detach().cpu().numpy(), self._adam_v.detach().cpu().numpy()
            adam_m = {n: m_host[o:o + r * c].reshape(self._shape_of(n)) for n, (o, r, c) in self._layout.items()}
            adam_v = {n: v_host[o:o + r * c].reshape(self._shape_of(n)) for n, (o, r, c) in self._layout.items()}
        counters = {'step': self._global_step, 'epoch': self._epoch, 'iterations': self._adam_t}
        prefix = str(logdir / ('ckpt-%d' % number))
        tf_checkpoint.write_bundle(prefix, tf_checkpoint.from_arrays(weights, adam_m, adam_v, counters))
        (logdir / 'checkpoint').write_text('model_checkpoint_path: "ckpt-%d"\nall_model_checkpoint_paths: "ckpt-%d"\n'
                                           % (number, number))
        return prefix

    def _restore(self, path, with_optimizer):
        if not str(path).endswith('.npz'):
            return self._restore_tf(str(path), with_optimizer)
        data = np.load(path)
        self.set_weights({name: data['variables/' + name] for name in self._layout})
        self._global_step = int(data['step'])
        self._epoch = int(data['epoch'])
        if with_optimizer and 'optimizer/m' in data.files:
            self._bind(1, min(self.window_size, 64), training=True)
            self._adam_m.copy_(torch.from_numpy(data['optimizer/m']))
            self._adam_v.copy_(torch.from_numpy(data['optimizer/v']))
            self._adam_t = int(data['optimizer/iterations'])

    def load_from_checkpoint(self, restoredir):
        '''BaseModel.load_from_checkpoint (models/__init__.py:66-90): weights only, errors exit(1).'''

        try:
            path = self.latest_checkpoint(restoredir)
            if path is None:
                raise FileNotFoundError('no checkpoint in %s' % restoredir)
            self._restore(path, with_optimizer=False)
            logging.info('{} model restored from \'{}\' (epoch {}, global step {}).'.format(
                self.__class__.__name__, path, self._epoch, self._global_step))
        except Exception:
            logging.exception('Failed to restore {} model from \'{}\'.'.format(self.__class__.__name__, restoredir))
            raise SystemExit(1)

    # ------------------------------------------------------------------
    # Training loop (transformer.py:846-960)
    # ------------------------------------------------------------------
    def train(self, dataset, input_shape, logdir, restoredir=None, epochs=None,
              learning_rate=1e-3, save_frequency_mode=ModelSaveFrequencyMode.EPOCH,
              save_frequency=1, max_checkpoints=1, show_progress_bar=True, max_steps=None, log_every=10):
        '''
        Fits the model to ``dataset`` (any re-iterable of ``(x, y)`` integer
        batches [batch, window]) with the reference's loop structure: Adam,
        mean sparse-categorical cross-entropy, per-step ``loss`` / ``accuracy``
        scalars, ``epoch_loss`` / ``epoch_accuracy`` per epoch, checkpoints every
        ``save_frequency`` steps or epochs.  Like the reference, ``epochs=N``
        stops when the 1-based epoch counter reaches N (transformer.py:890, 907).

        Differences, all host-side: the per-step scalars are read back every
        ``log_every`` steps in one copy (the reference formats them every step,
        a device sync per step; every step is still written), and ``max_steps``
        (extension) bounds the run for benchmarks and tests.
        '''

        from tqdm import tqdm

        logdir = Path(restoredir) if restoredir is not None else Path(logdir)
        self.compile(learning_rate)
        if restoredir is not None:
            try:
                path = self.latest_checkpoint(logdir)
                if path is None:
                    # tf.train.Checkpoint.restore(None) is a no-op (transformer.py:896-897): start from scratch
                    logging.info('No checkpoint in \'{}\'; initializing from scratch.'.format(logdir))
                else:
                    self._restore(path, with_optimizer=True)
                    logging.info('Model restored from \'{}\'.'.format(path))
            except Exception:
                logging.error('Failed to restore model from \'{}\'.'.format(restoredir))
                raise SystemExit(1)

        rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.sync_replicas()
        writer = None
        if rank == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                writer = SummaryWriter(str(logdir / 'train'))
            except Exception:   # tensorboard missing: scalars go to the log only
                logging.warning('tensorboard is not available; scalars are only logged.')

        save_frequency_mode = ModelSaveFrequencyMode(save_frequency_mode)
        steps_per_epoch = None
        steps_done = 0
        history = []
        # Per-step scalars are written for EVERY step like the reference (transformer.py:933-936), but read back
        # in one device->host copy every ``log_every`` steps from a small ring on the device.
        ring = torch.zeros((max(1, log_every), 2), dtype=torch.float32, device=self.device)
        ring_steps = []

        def flush_ring(progress_bar=None):
            if not ring_steps:
                return
            values = ring[:len(ring_steps)].cpu().numpy()          # host sync, every log_every steps
            for (step_number, tokens), (loss_sum_value, correct_value) in zip(ring_steps, values):
                loss_value, accuracy = float(loss_sum_value) / tokens, float(correct_value) / tokens
                history.append((step_number, loss_value, accuracy))
                if writer is not None:
                    writer.add_scalar('loss', loss_value, step_number)
                    writer.add_scalar('accuracy', accuracy, step_number)
            if progress_bar is not None:
                progress_bar.set_description('- loss: {:.4f} - accuracy: {:.4f}'.format(history[-1][1], history[-1][2]))
            ring_steps.clear()

        stopped_early = False
        while epochs is None or self._epoch < epochs:
            current_epoch = self._epoch
            logging.info('Epoch {}'.format(current_epoch if epochs is None else '{}/{}'.format(current_epoch, epochs)))
            epoch_loss = torch.zeros(1, dtype=torch.float64, device=self.device)
            epoch_correct = torch.zeros(1, dtype=torch.float64, device=self.device)
            epoch_tokens = 0
            epoch_batches = 0
            with tqdm(total=steps_per_epoch, disable=not (show_progress_bar and rank == 0)) as progress_bar:
                for x, y in dataset:
                    loss_sum, correct = self.train_step(x, y, learning_rate)
                    tokens = int(np.prod(np.shape(x)))
                    epoch_loss += loss_sum.double() / tokens
                    epoch_correct += correct.double()
                    epoch_tokens += tokens
                    epoch_batches += 1
                    global_step = self._global_step
                    slot = len(ring_steps)
                    ring[slot, 0] = loss_sum[0]
                    ring[slot, 1] = correct[0].float()
                    ring_steps.append((global_step, tokens))
                    if len(ring_steps) == ring.shape[0]:
                        flush_ring(progress_bar)
                    if save_frequency_mode == ModelSaveFrequencyMode.GLOBAL_STEP and \
                            global_step % save_frequency == 0 and rank == 0:
                        save_path = self.save_checkpoint(logdir, max_checkpoints)
                        progress_bar.write('Saved checkpoint for step {} at {}.'.format(global_step, save_path))
                    self._global_step += 1
                    steps_done += 1
                    progress_bar.update(1)
                    if max_steps is not None and steps_done >= max_steps:
                        stopped_early = True
                        break

                flush_ring(progress_bar)
                if stopped_early:
                    # a bounded run (max_steps) ends inside an epoch: no epoch summaries, no epoch checkpoint, and
                    # the epoch counter stays, so that a resumed run does not skip the rest of this epoch
                    break
                if epoch_batches and writer is not None:
                    writer.add_scalar('epoch_loss', float(epoch_loss) / epoch_batches, current_epoch)
                    writer.add_scalar('epoch_accuracy', float(epoch_correct) / max(epoch_tokens, 1), current_epoch)
                if save_frequency_mode == ModelSaveFrequencyMode.EPOCH and current_epoch % save_frequency == 0 \
                        and rank == 0:
                    save_path = self.save_checkpoint(logdir, max_checkpoints)
                    progress_bar.write('Saved checkpoint for epoch {} at {}.'.format(current_epoch, save_path))
                if steps_per_epoch is None:
                    steps_per_epoch = progress_bar.n
                self._epoch += 1
            if stopped_early:
                break
            if epoch_batches == 0:
                logging.warning('The dataset yielded no batches; stopping.')
                break

        if writer is not None:
            writer.close()
        return history

'''
Self-checks of the CPU oracle (oracle/transformer_oracle.py).  The reference's
own tests never touch the model (PARITY UNPINNED, see the oracle's header), so
the restatement is pinned by internal consistency instead: cached decoding ==
full recompute, autograd == finite differences in fp64, Adam == closed form,
the F5 residual structure and the lower-right causal mask.
'''

import numpy as np
import torch

from oracle import transformer_oracle as oracle


def _tiny(dropout=0.0):
    cfg = oracle.OracleConfig(vocab_size=23, embedding_size=16, window_size=12, decoder_layers_count=2,
                              attention_head_count=4, attention_dropout_rate=dropout, residual_dropout_rate=dropout)
    weights = oracle.init_parameters(cfg, seed=1)
    rng = np.random.default_rng(2)
    for name in weights:
        if name.endswith('/bias') or name.endswith('/beta'):
            weights[name] = (0.1 * rng.standard_normal(weights[name].shape)).astype(np.float32)
    return cfg, weights


def test_parameter_count_default_config():
    cfg = oracle.OracleConfig()
    total = sum(int(np.prod(s)) for s in oracle.parameter_shapes(cfg).values())
    assert total == 6680576


def test_causal_mask_is_lower_right_anchored():
    mask = oracle.causal_attention_mask(2, 5, torch.float64).numpy()
    assert mask.tolist() == [[1, 1, 1, 1, 0], [1, 1, 1, 1, 1]]


def test_cached_decode_equals_full_recompute():
    cfg, weights = _tiny()
    prompt = np.array([[3, 7, 1], [4, 4, 9]])
    ids_a, logits_a = oracle.generate(weights, prompt, 6, cfg, greedy=True, use_cache=True)
    ids_b, logits_b = oracle.generate(weights, prompt, 6, cfg, greedy=True, use_cache=False)
    assert (ids_a == ids_b).all()
    np.testing.assert_allclose(logits_a, logits_b, rtol=1e-9, atol=1e-11)


def test_position_overflow_raises_like_tf_cpu():
    cfg, weights = _tiny()
    params = oracle.to_torch(weights)
    try:
        oracle.transformer_call(params, np.zeros((1, 13), dtype=np.int64), cfg)
    except IndexError:
        return
    raise AssertionError('expected an IndexError for positions beyond window_size')


def test_gradients_match_finite_differences():
    cfg, weights = _tiny()
    rng = np.random.default_rng(3)
    draw = rng.integers(0, cfg.vocab_size, size=(2, 9))
    x, y = draw[:, :-1], draw[:, 1:]
    loss, _, _, grads = oracle.loss_and_gradients(weights, x, y, cfg, dtype=torch.float64)

    def loss_at(params):
        tensors = oracle.to_torch(params, torch.float64)
        logits, _ = oracle.transformer_call(tensors, x, cfg)
        return float(oracle.sparse_categorical_crossentropy(y, logits))

    eps = 1e-5
    for name in ('wte/weight', 'h_1/attn/c_attn/weight', 'h_2/ln_2/gamma', 'h_2/mlp/c_proj/bias', 'wpe/embeddings'):
        base = {k: np.asarray(v, dtype=np.float64) for k, v in weights.items()}
        flat_index = int(rng.integers(0, base[name].size))
        index = np.unravel_index(flat_index, base[name].shape)
        plus = {k: v.copy() for k, v in base.items()}
        minus = {k: v.copy() for k, v in base.items()}
        plus[name][index] += eps
        minus[name][index] -= eps
        numeric = (loss_at(plus) - loss_at(minus)) / (2 * eps)
        np.testing.assert_allclose(grads[name][index], numeric, rtol=2e-4, atol=1e-8)
    assert abs(loss - np.log(cfg.vocab_size)) < 0.5


def test_block_adds_attention_to_the_normalised_stream():
    # F5: x1 = LN1(x); x2 = x1 + Attn(x1).  With attention and MLP weights zeroed the block returns
    # LN1(x) + biases, not x + ...
    cfg, weights = _tiny()
    for name in weights:
        if '/attn/' in name or '/mlp/' in name:
            weights[name] = np.zeros_like(weights[name])
    params = oracle.to_torch(weights)
    x = torch.randn(1, 5, cfg.embedding_size, dtype=torch.float64) * 3 + 1
    out, _ = oracle.decoder_block(x, params, 1, cfg)
    expected = oracle.layer_normalization(x, params['h_1/ln_1/gamma'], params['h_1/ln_1/beta'], 1e-5)
    np.testing.assert_allclose(out.numpy(), expected.numpy(), rtol=1e-12, atol=1e-12)


def test_adam_matches_closed_form_first_step():
    params = {'w': np.array([1.0, -2.0, 0.5], dtype=np.float64)}
    grads = {'w': np.array([0.1, -0.3, 0.0], dtype=np.float64)}
    state = oracle.AdamState(params, learning_rate=1e-3)
    state.apply(params, grads)
    g = grads['w']
    m, v = 0.1 * g, 0.001 * g * g
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expected = np.array([1.0, -2.0, 0.5]) - lr_t * m / (np.sqrt(v) + 1e-7)
    np.testing.assert_allclose(params['w'], expected, rtol=1e-12)
    # epsilon sits outside the bias correction: a zero gradient must not move the weight
    assert params['w'][2] == 0.5


def test_sampling_distribution_rows_sum_to_one():
    p = oracle.next_token_distribution(np.array([[0.0, 1.0, -2.0]]), 0.7)
    np.testing.assert_allclose(p.sum(axis=-1), 1.0)
    assert p[0, 1] > p[0, 0] > p[0, 2]


def test_oracle_matches_tf_reference():
    '''
    Pins the oracle to outputs of the reference itself when ``tests/golden/tf_reference.npz`` exists (written by
    ``tools/dump_tf_reference.py`` where TensorFlow and the reference are available; neither is in this image, hence
    the skip).  Tolerances: fp32 TensorFlow against the fp64 oracle.
    '''
    import os
    import pytest
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tf_reference.npz')
    if not os.path.exists(path):
        pytest.skip('no TensorFlow dump of the reference (tools/dump_tf_reference.py has to run where TensorFlow exists)')
    data = np.load(path)
    vocab, embedding, window, layers, heads = (int(v) for v in data['config'])
    cfg = oracle.OracleConfig(vocab_size=vocab, embedding_size=embedding, window_size=window, decoder_layers_count=layers,
                              attention_head_count=heads, attention_dropout_rate=0.0, residual_dropout_rate=0.0)
    weights = {k[len('param/'):]: data[k] for k in data.files if k.startswith('param/')}
    weights = type(oracle.init_parameters(cfg))((k, weights[k]) for k in oracle.parameter_shapes(cfg))
    x, y = data['x'], data['y']
    loss, _, logits, grads = oracle.loss_and_gradients(weights, x, y, cfg)
    np.testing.assert_allclose(logits, data['logits'], rtol=2e-4, atol=2e-4)
    assert abs(loss - float(data['loss'])) < 1e-4
    for name in weights:
        np.testing.assert_allclose(grads[name], data['grad/' + name], rtol=2e-3, atol=2e-6, err_msg=name)
    adam = oracle.AdamState(weights)
    updated = adam.apply({k: v.copy() for k, v in weights.items()}, grads)
    for name in weights:
        np.testing.assert_allclose(updated[name], data['adam/' + name], rtol=1e-4, atol=2e-5, err_msg=name)
    steps = data['decode_ids'].shape[1]
    ids, step_logits = oracle.generate(weights, x[:, :1], steps, cfg, greedy=True)
    np.testing.assert_allclose(step_logits, data['decode_logits'], rtol=2e-4, atol=2e-4)
    # greedy ids must agree wherever the reference's top-2 margin is not within rounding
    top2 = np.sort(data['decode_logits'], axis=-1)[..., -2:]
    decided = (top2[..., 1] - top2[..., 0]) > 1e-3
    first_undecided = np.where(~decided.all(axis=0))[0]
    upto = int(first_undecided[0]) if first_undecided.size else steps
    assert (ids[:, :upto] == data['decode_ids'][:, :upto]).all()


# ---------------------------------------------------------------------------
# Pin: the oracle against the reference's own transformer.py (executed under tests/golden/tf_shim.py)
# ---------------------------------------------------------------------------

def _golden_case(name):
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    if golden not in sys.path:
        sys.path.insert(0, golden)
    import model_cases
    data = np.load(os.path.join(golden, 'model_golden.npz'))
    index = list(model_cases.CASES).index(name)
    case = model_cases.CASES[name]
    cfg = model_cases.case_config(oracle, case)
    weights = model_cases.case_weights(oracle, cfg, 10 + index)
    batches = model_cases.case_batches(case, 10 + index)
    return case, cfg, weights, batches, {k[len(name) + 1:]: data[k] for k in data.files if k.startswith(name + '/')}


import pytest  # noqa: E402


@pytest.mark.parametrize('name', ['small', 'no_layernorm_no_scale', 'default_heads', 'engine'])
def test_oracle_matches_reference_transformer(name):
    '''
    ``tests/golden/model_golden.npz`` holds what the reference's own ``Transformer`` (composer/models/transformer.py,
    unmodified) computed for these weights and ids: ``call`` logits and presents, greedy ``past=`` decoding, and
    the losses / accuracies / gradients / final variables of its own ``train`` loop.  The oracle must reproduce all
    of it in fp64 (the fixture keeps gradients and variables in fp32, hence 1e-6 there).
    '''

    case, cfg, weights, batches, golden = _golden_case(name)
    x0, y0 = batches[0]
    params = oracle.to_torch(weights, torch.float64)
    logits, presents = oracle.transformer_call(params, x0, cfg)
    reduced = case.get('reduced', False)      # the 'engine' case keeps fp32 logits and norms of the large tensors
    np.testing.assert_allclose(logits.numpy(), golden['logits'], rtol=0, atol=2e-6 if reduced else 1e-11)
    ours = np.stack([np.stack([p[0].numpy(), p[1].numpy()]) if isinstance(p, (tuple, list)) else p.numpy()
                     for p in presents])
    if 'presents' in golden:
        np.testing.assert_allclose(ours, golden['presents'], rtol=0, atol=1e-6)

    steps = golden['decode_ids'].shape[1]
    ids, step_logits = oracle.generate(weights, x0[:, :case['prompt']], steps, cfg, greedy=True)
    np.testing.assert_allclose(step_logits, golden['decode_logits'], rtol=0, atol=2e-6 if reduced else 1e-10)
    assert (ids == golden['decode_ids']).all()

    # the reference's train loop: per-step loss and accuracy, first-step gradients, variables after the last step
    current = type(weights)((k, np.asarray(v, dtype=np.float64)) for k, v in weights.items())
    adam = oracle.AdamState(current)
    for step, (x, y) in enumerate(batches):
        loss, accuracy, _, grads = oracle.loss_and_gradients(current, x, y, cfg)
        assert abs(loss - golden['step_loss'][step]) < 1e-11
        assert abs(accuracy - golden['step_accuracy'][step]) < 1e-12
        if step == 0:
            for variable in weights:
                key = 'grad/' + variable
                if key in golden:
                    np.testing.assert_allclose(grads[variable], golden[key], rtol=2e-6, atol=1e-9, err_msg=variable)
                elif 'grad_norm/' + variable in golden:
                    g = np.asarray(grads[variable])
                    np.testing.assert_allclose(np.linalg.norm(g), golden['grad_norm/' + variable], rtol=1e-9)
                    np.testing.assert_allclose(g.reshape(-1)[::max(1, g.size // 64)][:64], golden['grad_sample/' + variable],
                                               rtol=2e-6, atol=1e-9, err_msg=variable)
                else:   # never built by the reference (ln_1 / ln_2 without LayerNorm): no gradient here either
                    assert '/ln_' in variable and not np.any(grads[variable])
        current = adam.apply(current, grads)
    for variable in weights:
        if 'update_norm/' + variable in golden:
            np.testing.assert_allclose(np.linalg.norm(current[variable] - weights[variable]),
                                       golden['update_norm/' + variable], rtol=1e-6, err_msg=variable)
        if 'trained/' + variable in golden:
            np.testing.assert_allclose(current[variable], golden['trained/' + variable], rtol=0, atol=2e-7,
                                       err_msg=variable)
            # the update itself (3e-3 per step) must agree, not just the weights it is added to
            np.testing.assert_allclose(current[variable] - weights[variable],
                                       golden['trained/' + variable] - weights[variable], rtol=0, atol=3e-7)


def test_model_golden_is_what_the_reference_computes():
    '''Where /root/reference exists: re-run the reference's transformer.py under the shim, compare with the fixture.'''

    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    if golden not in sys.path:
        sys.path.insert(0, golden)
    from reference_shims import load_reference_transformer
    loaded = load_reference_transformer()
    if loaded is None:
        pytest.skip('the reference tree is not available here')
    reference, _ = loaded
    import make_model_golden
    case, cfg, weights, batches, golden_case = _golden_case('small')
    model, _ = make_model_golden.build_reference_model(reference, case, weights)
    logits, presents = model(batches[0][0])
    np.testing.assert_allclose(logits.numpy(), golden_case['logits'], rtol=0, atol=1e-12)
    assert type(model).__module__ == 'composer.models.transformer'
    assert os.path.realpath(sys.modules['composer.models.transformer'].__file__).startswith(
        os.path.realpath(os.environ.get('COMPOSER_REFERENCE_ROOT', '/root/reference')))

'''
The model-parity cases shared by ``make_model_golden.py`` (which runs the reference's own ``transformer.py`` under
``tf_shim.py``) and ``tests/test_oracle.py`` (which runs the oracle on the same weights and inputs).  Weights come
from a recipe, not from the fixture file: ``oracle.init_parameters`` (the reference's initialiser distribution) with
biases, LayerNorm gains / offsets perturbed so that no term of the model is multiplied by 0 or 1.
'''

import numpy as np

# name -> constructor arguments of ``Transformer`` (transformer.py:610-614) that differ from the defaults + shapes
CASES = {
    'small': dict(vocab_size=60, embedding_size=32, window_size=16, decoder_layers_count=2, attention_head_count=4,
                  scale=True, use_layer_normalization=True, batch=2, length=12, prompt=3, train_steps=3),
    'no_layernorm_no_scale': dict(vocab_size=41, embedding_size=24, window_size=12, decoder_layers_count=2,
                                  attention_head_count=3, scale=False, use_layer_normalization=False, batch=3,
                                  length=9, prompt=2, train_steps=2),
    'default_heads': dict(vocab_size=390, embedding_size=32, window_size=24, decoder_layers_count=2,
                          attention_head_count=16, scale=True, use_layer_normalization=True, batch=2, length=20,
                          prompt=4, train_steps=2),
    # a shape the CUDA engine accepts (embedding a multiple of 256, d_h 16): the GPU tests compare the engine with THIS
    # case directly, not only with the oracle.  Stored reduced (fp32 logits, norms of the large gradients / updates).
    'engine': dict(vocab_size=390, embedding_size=256, window_size=64, decoder_layers_count=2, attention_head_count=16,
                   scale=True, use_layer_normalization=True, batch=2, length=48, prompt=5, train_steps=3, reduced=True),
}


def case_weights(oracle, cfg, seed):
    weights = oracle.init_parameters(cfg, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    for name in weights:
        if name.endswith('/bias') or name.endswith('/beta'):
            weights[name] = (0.1 * rng.standard_normal(weights[name].shape)).astype(np.float32)
        elif name.endswith('/gamma'):
            weights[name] = (1.0 + 0.1 * rng.standard_normal(weights[name].shape)).astype(np.float32)
    return weights


def case_config(oracle, case):
    return oracle.OracleConfig(vocab_size=case['vocab_size'], embedding_size=case['embedding_size'],
                               window_size=case['window_size'], decoder_layers_count=case['decoder_layers_count'],
                               attention_head_count=case['attention_head_count'], attention_dropout_rate=0.0,
                               residual_dropout_rate=0.0, scale_attention=case['scale'],
                               use_layer_normalization=case['use_layer_normalization'])


def case_batches(case, seed):
    '''``train_steps`` (x, y) batches of ids, y = x shifted by one (models/__init__.py:304).'''

    rng = np.random.default_rng(seed + 2000)
    batches = []
    for _ in range(case['train_steps']):
        draw = rng.integers(0, case['vocab_size'], size=(case['batch'], case['length'] + 1))
        batches.append((draw[:, :-1].copy(), draw[:, 1:].copy()))
    return batches

'''
Generates ``tests/golden/model_golden.npz``: outputs of the REFERENCE's own ``composer/models/transformer.py``
(imported unmodified from /root/reference) executed with ``tests/golden/tf_shim.py`` standing in for TensorFlow.

    python tests/golden/make_model_golden.py

Per case of ``model_cases.py``:
  logits, presents         ``Transformer.call(x)`` (transformer.py:696-833), training=False
  step_loss, step_accuracy the scalars the reference's own ``Transformer.train`` loop (transformer.py:836-960) logged
  grad/<name>              the gradients its GradientTape produced at the first step (captured at apply_gradients)
  trained/<name>           every variable after the loop (``train_steps`` Adam updates)
  decode_ids, decode_logits  greedy decoding through ``past=`` (transformer.py:735-770, 423-437): the prompt once, then
                           one id per call with the presents of the previous call
'''

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from reference_shims import load_reference_transformer   # noqa: E402
import model_cases                                         # noqa: E402
from oracle import transformer_oracle as oracle           # noqa: E402  (initialiser recipe only)


def build_reference_model(reference, case, weights):
    model = reference.Transformer(case['vocab_size'], case['embedding_size'], case['window_size'],
                                  case['decoder_layers_count'], case['attention_head_count'], False, 0.0, 0.02,
                                  0.0, 0.0, 1e-5, case['scale'], case['use_layer_normalization'])
    model(np.zeros((1, 2), dtype=np.int64))                # first call builds the variables
    variables = dict(model.named_variables())
    # (without LayerNorm the ln_1 / ln_2 layers are never called, hence never built: they own no variables)
    missing = sorted(set(weights) - set(variables))
    assert not set(variables) - set(weights) and all('/ln_' in name for name in missing), missing
    with torch.no_grad():
        for name, value in weights.items():
            if name in variables:
                variables[name].as_subclass(torch.Tensor).copy_(torch.as_tensor(value, dtype=torch.float64))
    return model, variables


def main():
    loaded = load_reference_transformer()
    if loaded is None:
        raise SystemExit('the reference tree is not available')
    reference, shim = loaded
    out = {}
    for index, (name, case) in enumerate(model_cases.CASES.items()):
        seed = 10 + index
        cfg = model_cases.case_config(oracle, case)
        weights = model_cases.case_weights(oracle, cfg, seed)
        batches = model_cases.case_batches(case, seed)
        x0 = batches[0][0]

        model, variables = build_reference_model(reference, case, weights)
        logits, presents = model(x0)
        out[name + '/logits'] = logits.numpy()
        out[name + '/presents'] = np.stack([p.numpy() for p in presents])          # [L, 2, B, H, T, d_h]

        # greedy decode through past=
        prompt = x0[:, :case['prompt']]
        steps = case['window_size'] - case['prompt'] + 1
        ids, step_logits = [], []
        logits, past = model(prompt)
        for _ in range(steps):
            last = logits.numpy()[:, -1, :]
            step_logits.append(last)
            chosen = last.argmax(axis=-1)
            ids.append(chosen)
            if len(ids) < steps:
                logits, past = model(chosen[:, None], past=past)
        out[name + '/decode_ids'] = np.stack(ids, axis=1)
        out[name + '/decode_logits'] = np.stack(step_logits, axis=1)

        # the reference's own training loop
        first_grads = {}
        original_apply = shim.Adam.apply_gradients

        def recording_apply(self, grads_and_vars):
            pairs = list(grads_and_vars)
            if not first_grads:
                by_id = {id(v): n for n, v in variables.items()}
                for grad, variable in pairs:
                    first_grads[by_id[id(variable)]] = None if grad is None else grad.numpy().copy()
            return original_apply(self, pairs)

        shim.Adam.apply_gradients = recording_apply
        del shim.SCALARS[:]
        try:
            with tempfile.TemporaryDirectory() as logdir:
                model.train(batches, (case['batch'], case['length']), logdir, epochs=2, learning_rate=1e-3,
                            show_progress_bar=False)
        finally:
            shim.Adam.apply_gradients = original_apply
        out[name + '/step_loss'] = np.array([v for tag, v, _ in shim.SCALARS if tag == 'loss'])
        out[name + '/step_accuracy'] = np.array([v for tag, v, _ in shim.SCALARS if tag == 'accuracy'])
        assert len(out[name + '/step_loss']) == case['train_steps']
        reduced = case.get('reduced', False)
        for variable_name, grad in first_grads.items():
            if grad is None:
                continue
            if reduced and grad.size > 4096:         # large tensors: their L2 norm and 64 strided samples
                out[name + '/grad_norm/' + variable_name] = np.asarray(np.linalg.norm(grad))
                out[name + '/grad_sample/' + variable_name] = grad.reshape(-1)[::max(1, grad.size // 64)][:64].copy()
            else:
                out[name + '/grad/' + variable_name] = grad
        for variable_name, variable in variables.items():
            value = variable.numpy().copy()
            if reduced and value.size > 4096:
                delta = value - np.asarray(weights[variable_name], dtype=np.float64)
                out[name + '/update_norm/' + variable_name] = np.asarray(np.linalg.norm(delta))
            else:
                out[name + '/trained/' + variable_name] = value
        if reduced:
            out[name + '/logits'] = out[name + '/logits'].astype(np.float32)
            out[name + '/decode_logits'] = out[name + '/decode_logits'].astype(np.float32)
            del out[name + '/presents']
        print(name, 'loss', out[name + '/step_loss'], 'decode', out[name + '/decode_ids'][0, :8])

    path = os.path.join(HERE, 'model_golden.npz')
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 and ('/trained/' in k or '/grad/' in k or k.endswith('/presents')) else v)
                                 for k, v in out.items()})
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()

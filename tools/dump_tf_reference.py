'''Pins the model oracle to the real reference: run this where TensorFlow 2.x and a checkout of the reference exist.

    python tools/dump_tf_reference.py --reference /path/to/composer --out tests/golden/tf_reference.npz

TensorFlow is not installable in the image this repository is developed in (no wheel, no network), so the
oracle in ``oracle/transformer_oracle.py`` is checked against the reference's *code*, not against its outputs
("parity unpinned").  This script closes that gap from the other side: it builds the reference's own
``composer.models.Transformer`` (transformer.py:599-960) under TensorFlow at a small configuration, loads the
weights this repository's initialiser draws (same Keras variable names), runs

  * one forward pass without dropout (``training=False``, transformer.py:696-833) -> logits,
  * the loss of transformer.py:888/918 on the shifted labels,
  * one Adam step (transformer.py:887, 914-921) -> the updated weights,
  * a greedy KV-cache decode through ``past=`` (transformer.py:423-437; cli.py:663-676 with argmax),

and stores inputs and outputs in one ``.npz``.  ``tests/test_oracle.py::test_oracle_matches_tf_reference`` compares the
oracle with that file when it is present and is skipped otherwise.  The script has NOT been executed in this
repository's image (no TensorFlow); it only uses the reference's public constructor and ``call`` signature.
'''
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument('--reference', required=True, help='checkout of galacticglum/composer')
    parser.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden', 'tf_reference.npz'))
    parser.add_argument('--vocab', type=int, default=390)
    parser.add_argument('--embedding', type=int, default=64)
    parser.add_argument('--window', type=int, default=48)
    parser.add_argument('--layers', type=int, default=2)
    parser.add_argument('--heads', type=int, default=4)
    parser.add_argument('--batch', type=int, default=3)
    parser.add_argument('--decode-steps', type=int, default=16)
    parser.add_argument('--seed', type=int, default=0)
    args = parser.parse_args()

    import tensorflow as tf                                   # noqa: E402 (the point of this script)
    sys.path.insert(0, args.reference)
    from composer.models.transformer import Transformer       # the reference's class, unmodified

    from oracle import transformer_oracle as oracle            # only for the initialiser and the variable names
    cfg = oracle.OracleConfig(vocab_size=args.vocab, embedding_size=args.embedding, window_size=args.window,
                              decoder_layers_count=args.layers, attention_head_count=args.heads,
                              attention_dropout_rate=0.0, residual_dropout_rate=0.0)
    params = oracle.init_parameters(cfg, seed=args.seed)
    perturb = np.random.default_rng(args.seed + 2)             # zeros / ones would hide a swapped bias or gamma
    for key in params:
        if key.endswith('/bias') or key.endswith('/beta') or key.endswith('/gamma'):
            params[key] = (params[key] + 0.1 * perturb.standard_normal(params[key].shape)).astype(np.float32)

    model = Transformer(args.vocab, args.embedding, args.window, args.layers, args.heads, False, 0, 0.02, 0.0, 0.0,
                        1e-5, True, True)
    rng = np.random.default_rng(args.seed + 1)
    draw = rng.integers(0, args.vocab, size=(args.batch, args.window + 1)).astype(np.int32)
    x, y = draw[:, :-1], draw[:, 1:]
    model(tf.constant(x), training=False)                      # builds the variables

    # Map the Keras variables onto the oracle's names.  The reference names its variables
    # '<model>/wte/weight:0', '<model>/h_1/attn/c_attn/weight:0', ... (transformer.py:110-118, 186-192, 550-566,
    # 670-694); match on the suffix so that the model-level prefix does not matter.
    assigned = set()
    for variable in model.variables:
        name = variable.name.split(':')[0]
        matches = [key for key in params if name.endswith(key)]
        if len(matches) != 1:
            raise SystemExit('cannot map TF variable %r onto the oracle names (matches: %r)' % (variable.name, matches))
        key = matches[0]
        value = params[key].reshape(variable.shape.as_list())
        variable.assign(value)
        assigned.add(key)
    missing = set(params) - assigned
    if missing:
        raise SystemExit('oracle parameters without a TF variable: %s' % sorted(missing))

    out = {'config': np.array([args.vocab, args.embedding, args.window, args.layers, args.heads], dtype=np.int64),
           'seed': np.array(args.seed), 'x': x, 'y': y}
    for key, value in params.items():
        out['param/' + key] = value

    # forward + loss (transformer.py:916-918)
    logits = model(tf.constant(x), training=False)[0]
    loss_fn = tf.keras.losses.SparseCategoricalCrossentropy(from_logits=True)
    out['logits'] = logits.numpy()
    out['loss'] = np.array(loss_fn(y, logits).numpy())

    # greedy cached decode from the first token of every row (cli.py:663-676 with argmax instead of a draw)
    ids = x[:, :1]
    past, generated = None, []
    for _ in range(args.decode_steps):
        result = model(tf.constant(ids), past=past, training=False, use_cache=True)
        step_logits, past = result[0], result[1]
        ids = tf.argmax(step_logits[:, -1, :], axis=-1).numpy().astype(np.int32)[:, None]
        generated.append(ids[:, 0])
        out.setdefault('decode_logits', []).append(step_logits[:, -1, :].numpy())
    out['decode_ids'] = np.stack(generated, axis=1)
    out['decode_logits'] = np.stack(out['decode_logits'], axis=1)

    # one Adam step (transformer.py:887, 914-921), dropout off so that it is deterministic
    optimizer = tf.keras.optimizers.Adam(learning_rate=1e-3)
    with tf.GradientTape() as tape:
        step_logits = model(tf.constant(x), training=False)[0]
        step_loss = loss_fn(y, step_logits)
    gradients = tape.gradient(step_loss, model.trainable_variables)
    optimizer.apply_gradients(zip(gradients, model.trainable_variables))
    for variable, gradient in zip(model.trainable_variables, gradients):
        name = variable.name.split(':')[0]
        key = [k for k in params if name.endswith(k)][0]
        dense = tf.convert_to_tensor(gradient).numpy()           # the embedding gradient is IndexedSlices
        out['grad/' + key] = dense.reshape(params[key].shape)
        out['adam/' + key] = variable.numpy().reshape(params[key].shape)

    np.savez_compressed(args.out, **out)
    print('wrote %s (%d arrays, TensorFlow %s)' % (args.out, len(out), tf.__version__))


if __name__ == '__main__':
    main()

'''
Converts between the reference's TensorFlow checkpoints and this package's ``ckpt-N.npz`` files, on the host,
without TensorFlow (composer_b200/tf_checkpoint.py).

    python tools/convert_checkpoint.py to-npz  <reference logdir | checkpoint prefix> <out logdir>
    python tools/convert_checkpoint.py to-tf   <logdir with ckpt-N.npz>              <out logdir>
    python tools/convert_checkpoint.py list    <reference logdir | checkpoint prefix>

``to-npz`` writes what ``composer_b200 train --restoredir`` / ``generate`` read; ``to-tf`` writes ``ckpt-1.index`` /
``ckpt-1.data-00000-of-00001`` / ``checkpoint`` under the reference's object-graph names.  No device is needed.
'''
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from composer_b200 import tf_checkpoint  # noqa: E402


def _prefix(path):
    if os.path.isdir(path):
        prefix = tf_checkpoint.latest_checkpoint(path)
        if prefix is None:
            raise SystemExit('no `checkpoint` state file in %s' % path)
        return prefix
    return path[:-len('.index')] if path.endswith('.index') else path


def main(argv):
    if len(argv) < 2:
        raise SystemExit(__doc__)
    command = argv[0]
    if command == 'list':
        for name, array in sorted(tf_checkpoint.read_bundle(_prefix(argv[1])).items()):
            print('%-90s %-8s %s' % (name, array.dtype, tuple(array.shape)))
        return
    out = argv[2]
    os.makedirs(out, exist_ok=True)
    if command == 'to-npz':
        bundle = tf_checkpoint.read_bundle(_prefix(argv[1]))
        names = sorted({k[len('model/'):-len(tf_checkpoint.SUFFIX)] for k in bundle
                        if k.startswith('model/') and k.endswith(tf_checkpoint.SUFFIX) and '.OPTIMIZER_SLOT' not in k})

        def keras_name(path):
            parts = path.split('/')
            if parts[0] == 'decoder_blocks':
                parts = ['h_%d' % (int(parts[1]) + 1)] + parts[2:]
            return '/'.join(parts)

        keras = [keras_name(n) for n in names]
        shapes = {keras_name(n): bundle['model/' + n + tf_checkpoint.SUFFIX].shape for n in names}
        weights, adam_m, adam_v, counters = tf_checkpoint.to_arrays(bundle, keras, shapes)
        arrays = {'variables/' + k: v for k, v in weights.items()}
        arrays['step'] = np.asarray(counters.get('step', 1), dtype=np.int64)
        arrays['epoch'] = np.asarray(counters.get('epoch', 1), dtype=np.int64)
        arrays['optimizer/iterations'] = np.asarray(counters.get('iterations', 0), dtype=np.int64)
        # (the Adam slots of a .npz are flat arenas in the engine's layout: restore them through
        #  Transformer.load_from_checkpoint on the TensorFlow files directly when training is to be resumed)
        np.savez(os.path.join(out, 'ckpt-1.npz'), **arrays)
        with open(os.path.join(out, 'checkpoint.json'), 'w') as handle:
            json.dump({'all': ['ckpt-1.npz'], 'latest': 'ckpt-1.npz'}, handle)
        print('wrote %s (%d variables)' % (os.path.join(out, 'ckpt-1.npz'), len(weights)))
    elif command == 'to-tf':
        with open(os.path.join(argv[1], 'checkpoint.json')) as handle:
            latest = json.load(handle)['latest']
        data = np.load(os.path.join(argv[1], latest))
        weights = {k[len('variables/'):]: data[k] for k in data.files if k.startswith('variables/')}
        counters = {'step': int(data['step']), 'epoch': int(data['epoch']),
                    'iterations': int(data['optimizer/iterations']) if 'optimizer/iterations' in data.files else 0}
        prefix = os.path.join(out, 'ckpt-1')
        tf_checkpoint.write_bundle(prefix, tf_checkpoint.from_arrays(weights, None, None, counters))
        with open(os.path.join(out, 'checkpoint'), 'w') as handle:
            handle.write('model_checkpoint_path: "ckpt-1"\nall_model_checkpoint_paths: "ckpt-1"\n')
        print('wrote %s.index / .data-00000-of-00001 (%d variables)' % (prefix, len(weights)))
    else:
        raise SystemExit(__doc__)


if __name__ == '__main__':
    main(sys.argv[1:])

'''
Imports the reference's tokenizer (``/root/reference/composer/dataset/sequence.py``)
in this container.  Two in-memory shims are needed: a stub
``pretty_midi`` (only used by MIDI file I/O) and ``np.int`` (removed from numpy).
Returns ``None`` when the reference tree is absent (e.g. on the GPU box).
'''

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('COMPOSER_REFERENCE_ROOT', '/root/reference')


def load_reference_sequence():
    path = os.path.join(REFERENCE_ROOT, 'composer', 'dataset', 'sequence.py')
    if not os.path.exists(path):
        return None

    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    if 'pretty_midi' not in sys.modules:
        stub = types.ModuleType('pretty_midi')
        for name in ('PrettyMIDI', 'Instrument', 'Note', 'ControlChange'):
            setattr(stub, name, type(name, (), {}))
        sys.modules['pretty_midi'] = stub
    if 'composer' not in sys.modules:
        # ``import composer`` would pull in the click CLI; only composer.exceptions is needed here.
        package = types.ModuleType('composer')
        package.__path__ = []
        sys.modules['composer'] = package
        exc_spec = importlib.util.spec_from_file_location(
            'composer.exceptions', os.path.join(REFERENCE_ROOT, 'composer', 'exceptions.py'))
        exceptions = importlib.util.module_from_spec(exc_spec)
        exc_spec.loader.exec_module(exceptions)
        sys.modules['composer.exceptions'] = exceptions
        package.exceptions = exceptions
    spec = importlib.util.spec_from_file_location('_reference_sequence', path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module

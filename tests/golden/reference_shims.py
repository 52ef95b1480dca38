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


def load_reference_transformer():
    '''
    Imports the reference's own ``composer/models/transformer.py`` (unmodified, from ``/root/reference``) with
    ``tests/golden/tf_shim.py`` standing in for TensorFlow.  Returns ``(module, tf_shim)`` or ``None`` when the
    reference tree is absent (e.g. on the GPU box).  ``composer/__init__.py`` ends with ``from composer.cli import
    cli`` (click, the whole CLI): that one submodule is pre-registered as an empty stub, everything else
    (``composer``, ``composer.models``, ``composer.dataset.sequence``, ``composer.utils``) is the reference's file.
    '''

    if not os.path.exists(os.path.join(REFERENCE_ROOT, 'composer', 'models', 'transformer.py')):
        return None

    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    if not hasattr(np, 'float'):
        np.float = float
    import tf_shim
    tensorflow = sys.modules.get('tensorflow')
    if tensorflow is None or not getattr(tensorflow, '__shim__', False):
        tf_shim.install()
    if 'pretty_midi' not in sys.modules:
        stub = types.ModuleType('pretty_midi')
        for name in ('PrettyMIDI', 'Instrument', 'Note', 'ControlChange'):
            setattr(stub, name, type(name, (), {}))
        sys.modules['pretty_midi'] = stub
    # drop the partial in-memory ``composer`` package load_reference_sequence() may have registered
    for name in [n for n in sys.modules if n == 'composer' or n.startswith('composer.')]:
        del sys.modules[name]
    cli_stub = types.ModuleType('composer.cli')
    cli_stub.cli = None
    sys.modules['composer.cli'] = cli_stub
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        module = importlib.import_module('composer.models.transformer')
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return module, tf_shim

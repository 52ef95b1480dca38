'''
composer_b200: a B200-native implementation of Composer's Transformer path
(training on event tokens + temperature-sampled generation) behind the
reference's own CLI, model registry and configuration files.
'''

from enum import Enum, unique


@unique
class ModelSaveFrequencyMode(Enum):
    '''Units of ``--save-freq`` (reference: composer/__init__.py:4-16).'''

    EPOCH = 'epoch'
    GLOBAL_STEP = 'step'


def __getattr__(name):
    # ``composer_b200.cli`` (the click group) is resolved lazily so that
    # importing the package does not pull in click/torch.
    if name == 'cli':
        from composer_b200.cli import cli
        return cli

    raise AttributeError(name)

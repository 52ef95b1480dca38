'''
Models of the B200 build.  Only the Transformer path is provided; ``MusicRNN``
stays importable so that the registry (``composer_b200.cli.ModelType``) keeps the
reference's entries, but constructing it raises.
'''

from composer_b200.models.base import BaseModel


def __getattr__(name):
    # Resolved lazily: importing the package must not need torch / CUDA.
    if name == 'Transformer':
        from composer_b200.models.transformer import Transformer
        return Transformer
    if name == 'MusicRNN':
        return _MusicRNN
    raise AttributeError(name)


class _MusicRNN(BaseModel):
    '''Placeholder for the reference's LSTM model (composer/models/music_rnn.py), which is out of scope.'''

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('The MusicRNN model is not provided by the B200 build; use model type "transformer".')

    def train(self, *args, **kwargs):
        raise NotImplementedError()

    def load_from_checkpoint(self, restoredir):
        raise NotImplementedError()

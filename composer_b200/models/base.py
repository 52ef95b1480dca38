'''
The model contract used by the command-line interface.

Mirrors the reference's ``BaseModel`` (composer/models/__init__.py:12-90) minus
its Keras base class: ``train`` is abstract, ``load_from_checkpoint`` restores
the latest checkpoint of a log directory.
'''

from abc import ABC, abstractmethod

from composer_b200 import ModelSaveFrequencyMode


class BaseModel(ABC):
    '''A generic model interface class for use with the command-line interface.'''

    @abstractmethod
    def train(self, dataset, input_shape, logdir, restoredir=None, epochs=None,
              learning_rate=1e-3, save_frequency_mode=ModelSaveFrequencyMode.EPOCH,
              save_frequency=1, max_checkpoints=1, show_progress_bar=True):
        '''
        Fit the model to ``dataset``: an iterable of batched ``(features, labels)``
        integer pairs of shape ``(batch_size, window_size)``.  Argument meaning is
        that of the reference (composer/models/__init__.py:18-64).
        '''

        raise NotImplementedError()

    @abstractmethod
    def load_from_checkpoint(self, restoredir):
        '''Loads the latest checkpoint found in ``restoredir`` (composer/models/__init__.py:66-90).'''

        raise NotImplementedError()

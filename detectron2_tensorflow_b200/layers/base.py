"""`Layer` base-class contract of the reference (lib/layers/base.py:11-41), minus TensorFlow.

The reference's `Layer.__call__` only opens a `tf.name_scope(self.scope)` around `call`;
here it opens an NVTX range of the same name so profiles keep the reference's op labels.
"""
from abc import ABCMeta, abstractmethod

import torch

_TRAINING_PHASE = False


def set_training_phase(training):
    """lib/utils/tf_utils.py set_training_phase equivalent (global train/eval flag)."""
    global _TRAINING_PHASE
    _TRAINING_PHASE = bool(training)


def get_training_phase():
    return _TRAINING_PHASE


class Layer(object, metaclass=ABCMeta):

    def __init__(self, dtype=torch.float32, scope=None, **kwargs):
        self.dtype = dtype
        self.scope = self._set_scope(scope)
        for name, value in kwargs.items():
            setattr(self, name, value)
        if not hasattr(self, 'training'):
            self.training = get_training_phase()

    def _set_scope(self, scope=None):
        return self.__class__.__name__ if scope is None else scope

    def __call__(self, *args, **kwargs):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.scope)
            try:
                return self.call(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return self.call(*args, **kwargs)

    @abstractmethod
    def call(self):
        raise NotImplementedError

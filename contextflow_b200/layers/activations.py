"""Invertible activations.  Only the Sigmoid flow is on the path (inside the variational / argmax / prob encoders,
reference layers/activations.py:228-238); it is evaluated inside the fused encoder kernel.  The other activations of the
reference are never instantiated by create_model (model.py:105,137 are commented out) and are named here only so that
`from layers import *` resolves the names model.py mentions."""
import torch

from .flowlayer import FlowLayer

__all__ = ['FlowActivationLayer', 'Sigmoid', 'Softplus', 'SmoothLeakyRelu', 'SplineActivation', 'LearnableLeakyRelu']


class FlowActivationLayer(FlowLayer):
    def forward(self, input, context=None):
        raise NotImplementedError

    def reverse(self, input, context=None):
        raise NotImplementedError

    def logdet(self, input, context=None):
        raise NotImplementedError


class Sigmoid(FlowActivationLayer):
    def __init__(self, temperature=1, eps=0.0):
        super().__init__()
        self.eps = eps
        self.register_buffer('temperature', torch.Tensor([temperature]))


class Softplus(FlowActivationLayer):
    def __init__(self, eps=1e-7):
        super().__init__()
        self.eps = eps


def _off_path(name):
    class _Unused(FlowActivationLayer):
        def __init__(self, *a, **kw):
            raise NotImplementedError(f'{name} is never built by create_model and is outside the accelerated path')
    _Unused.__name__ = name
    return _Unused


SmoothLeakyRelu = _off_path('SmoothLeakyRelu')
SplineActivation = _off_path('SplineActivation')
LearnableLeakyRelu = _off_path('LearnableLeakyRelu')

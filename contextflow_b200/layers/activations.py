"""Invertible activations (reference layers/activations.py).  On the hot path the Sigmoid flow sits inside the variational / argmax /
prob encoders (dequantize.py:104,149,186) and is evaluated by the fused encoder kernel; used on their own, Sigmoid and Softplus
(activations.py:228-264) run the standalone CUDA kernels (forward, reverse, and the backward under autograd).  The other activations of
the reference are never instantiated by create_model (model.py:105,137 are commented out) and are named here only so that
`from layers import *` resolves the names model.py mentions."""
import torch
from torch.autograd import Function

from .. import ops
from .flowlayer import FlowLayer

__all__ = ['FlowActivationLayer', 'Sigmoid', 'Softplus', 'SmoothLeakyRelu', 'SplineActivation', 'LearnableLeakyRelu']


class _ActivationFn(Function):
    @staticmethod
    def forward(ctx, x, kind, temperature):
        ctx.save_for_backward(x)
        ctx.kind, ctx.temperature = kind, temperature
        return ops.activation_fwd(x, kind, temperature)

    @staticmethod
    def backward(ctx, dz, dldj):
        (x,) = ctx.saved_tensors
        return ops.activation_bwd(x, None if dz is None else dz.contiguous(), None if dldj is None else dldj.contiguous(), ctx.kind, ctx.temperature), None, None


class FlowActivationLayer(FlowLayer):
    kind = None

    def _temperature(self):
        return None

    def forward(self, input, context=None):
        if self.kind is None:
            raise NotImplementedError
        t = self._temperature()
        if torch.is_grad_enabled() and input.requires_grad:
            return _ActivationFn.apply(input, self.kind, t)
        return ops.activation_fwd(input, self.kind, t)

    def reverse(self, input, context=None):
        raise NotImplementedError

    def logdet(self, input, context=None):
        return self.forward(input, context)[1]


class Sigmoid(FlowActivationLayer):
    """activations.py:228-244: z = sigmoid(T x); ldj = sum_last(log T - softplus(-T x) - softplus(T x))."""
    kind = 'sigmoid'

    def __init__(self, temperature=1, eps=0.0):
        super().__init__()
        self.eps = eps
        self.register_buffer('temperature', torch.Tensor([temperature]))

    def _temperature(self):
        return self.temperature

    def reverse(self, z, context=None):
        lo, hi = torch.aminmax(z)                                   # activations.py:241 asserts on the host as well
        assert lo.item() >= 0 and hi.item() <= 1, 'input must be in [0,1]'
        return ops.activation_inv(z, 'sigmoid', self.eps, self.temperature)


class Softplus(FlowActivationLayer):
    """activations.py:247-264: z = softplus(x); ldj = sum_last logsigmoid(x); reverse x = z + log1p(-exp(-max(z, eps)))."""
    kind = 'softplus'

    def __init__(self, eps=1e-7):
        super().__init__()
        self.eps = eps

    def reverse(self, z, context=None):
        return ops.activation_inv(z, 'softplus', self.eps)


def _off_path(name):
    class _Unused(FlowActivationLayer):
        def __init__(self, *a, **kw):
            raise NotImplementedError(f'{name} is never built by create_model and is outside the accelerated path')
    _Unused.__name__ = name
    return _Unused


SmoothLeakyRelu = _off_path('SmoothLeakyRelu')
SplineActivation = _off_path('SplineActivation')
LearnableLeakyRelu = _off_path('LearnableLeakyRelu')

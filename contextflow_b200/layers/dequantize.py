"""Dequantisation layers (reference layers/dequantize.py): image Dequantization (:8-23) and the categorical context
surjections -- uniform (:26-70), variational (:73-126), eye (:129-142), prob (:145-166), argmax (:170-276).
Each surjection's forward((embedding, context)) -> (c, ldj) is one launch of the fused encoder kernel."""
import numpy as np
import torch

from .. import ops
from .activations import Sigmoid, Softplus
from .flowlayer import PreprocessingFlowLayer

__all__ = ['Dequantization', 'UniformCatDequantization', 'VariationalCatDequantization', 'EyeSampling', 'ProbSampling',
           'ArgmaxCatDequantization']


class Dequantization(PreprocessingFlowLayer):
    def __init__(self, dist):
        super().__init__()
        self.dist = dist           # support on [0,1]^d

    def forward(self, input, context=None):
        noise = self.dist.draw(input.size(0), device=input.device)
        return ops.add(input, noise), input.new_zeros(input.shape[0])

    def reverse(self, input, context=None):
        return ops.floor(input)                                     # dequantize.py:19-20

    def logdet(self, input, context=None):
        raise NotImplementedError


class _CatSurjection(PreprocessingFlowLayer):
    """forward((x, context)): x is the embedded context (int64 one-hot / raw ints, or float embedding)."""
    kind = None

    def forward(self, input):
        from ._encoder_desc import run_surjection
        x, context = input
        return run_surjection(self, x, context)

    def reverse(self, z, context=None):
        raise NotImplementedError('inverse path is outside this round (SURVEY §8f-3)')

    def logdet(self, x, context=None):
        raise NotImplementedError


class UniformCatDequantization(_CatSurjection):
    kind = 'uniform'

    def __init__(self, num_cats=[1]):
        super().__init__()
        self.D = len(num_cats)
        self.register_buffer('qbins', torch.tensor(num_cats, dtype=torch.float))
        self.register_buffer('ldj_per_dim', -torch.log(torch.tensor(num_cats, dtype=torch.float)))


class VariationalCatDequantization(_CatSurjection):
    kind = 'vardeq'

    def __init__(self, encoder, num_cats=[1]):
        super().__init__()
        self.D = len(num_cats)
        self.register_buffer('qbins', torch.tensor(num_cats, dtype=torch.float))
        self.register_buffer('ldj_per_dim', -torch.log(torch.tensor(num_cats, dtype=torch.float)))
        self.encoder = encoder
        self.sigmoid = Sigmoid()


class EyeSampling(_CatSurjection):
    kind = 'eyesample'

    def forward(self, input):
        x, context = input
        return x, torch.zeros(x.shape[0], device=x.device)


class ProbSampling(_CatSurjection):
    kind = 'probsample'

    def __init__(self, encoder):
        super().__init__()
        self.encoder = encoder
        self.sigmoid = Sigmoid()


class ArgmaxCatDequantization(_CatSurjection):
    kind = 'argmax'

    def __init__(self, encoder, num_cats=[1]):
        super().__init__()
        self.encoder = encoder
        self.num_bits = self.cats2bits(num_cats)
        self.sigmoid = Sigmoid()
        self.softplus = Softplus()

    @staticmethod
    def cats2bits(num_cats):
        if isinstance(num_cats, (list, tuple)):
            return [int(np.ceil(np.log2(cat))) for cat in num_cats]
        return int(np.ceil(np.log2(num_cats)))

"""Squeeze / UnSqueeze: space-to-depth index permutation, bit exact (reference layers/squeeze.py)."""
from .. import ops, training
from .flowlayer import FlowLayer


class Squeeze(FlowLayer):
    def __init__(self, patch_size=(2, 2)):
        super().__init__()
        self.p = patch_size

    def forward(self, input, context=None):
        if training.wants_grad(input):
            return training.SqueezeFn.apply(input, self.p[0], self.p[1]), self.logdet(input, context)
        return ops.squeeze(input, self.p[0], self.p[1]), self.logdet(input, context)

    def reverse(self, input, context=None):
        return ops.unsqueeze(input, self.p[0], self.p[1])

    def logdet(self, input, context=None):
        return input.new_zeros(len(input))


class UnSqueeze(FlowLayer):
    def __init__(self, patch_size=(2, 2)):
        super().__init__()
        self.p = patch_size

    def forward(self, input, context=None):
        return ops.unsqueeze(input, self.p[0], self.p[1]), self.logdet(input, context)

    def reverse(self, input, context=None):
        return ops.squeeze(input, self.p[0], self.p[1])

    def logdet(self, input, context=None):
        return input.new_zeros(len(input))

"""PermuteAxes: the ATM stack swaps channel and time axes (reference layers/permute_axes.py:5-21, model.py:149-151)."""
from collections.abc import Iterable

from .. import ops, training
from .flowlayer import FlowLayer


class PermuteAxes(FlowLayer):
    def __init__(self, permutation):
        super().__init__()
        assert isinstance(permutation, Iterable), 'permutation must be an Iterable'
        assert permutation[0] == 0, 'First element of permutation must be 0 (such that batch dimension stays intact)'
        self.permutation = tuple(permutation)
        self.inverse_permutation = sorted(range(len(self.permutation)), key=self.permutation.__getitem__)

    def _permute(self, x, perm):
        if tuple(perm) == tuple(range(len(perm))):
            return x.clone()
        if tuple(perm) != (0, 2, 1, 3):
            raise NotImplementedError(f'only the (0,2,1,3) axis swap has a kernel; got {tuple(perm)}')
        if training.wants_grad(x):
            return training.PermuteFn.apply(x)
        return ops.permute_chw(x)

    def forward(self, input, context=None):
        return self._permute(input, self.permutation), self.logdet(input, context)

    def reverse(self, input, context=None):
        return self._permute(input, self.inverse_permutation)

    def logdet(self, input, context=None):
        return input.new_zeros(len(input))

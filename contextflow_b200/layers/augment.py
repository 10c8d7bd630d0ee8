"""Augment: append noise channels, ldj = -log q(noise) of shape (B,1) (reference layers/augment.py:7-18)."""
import torch

from .. import ops
from .flowlayer import FlowLayer


class Augment(FlowLayer):
    def __init__(self, aug_distribution, aug_size, split_dim=1):
        super().__init__()
        self.distribution = aug_distribution
        self.aug_size = aug_size
        self.split_dim = split_dim

    def forward(self, input, context=None):
        if self.split_dim != 1:
            raise NotImplementedError('Augment concatenates along the channel axis on the hot path')
        noise = self.distribution.draw(input.size(0), device=input.device)
        if input.dim() == 2:                                  # (B, D) inputs inside context encoders
            y, ldj = ops.augment(input[:, :, None, None], noise.reshape(noise.shape[0], -1, 1, 1))
            return y.flatten(1), ldj.unsqueeze(-1)
        y, ldj = ops.augment(input, noise)
        return y, ldj.unsqueeze(-1)

    def reverse(self, input, context=None):
        keep = input.shape[self.split_dim] - self.aug_size             # augment.py:20-23: drop the noise channels
        if input.dim() == 4 and self.split_dim == 1:
            return ops.slice_channels(input, 0, keep)
        return input[:, :keep].contiguous()

    def logdet(self, input, context=None):
        raise NotImplementedError

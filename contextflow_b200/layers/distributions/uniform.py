"""UniformDistribution: the image dequantisation noise (reference layers/distributions/uniform.py)."""
import numpy as np
import torch
import torch.nn as nn

from ... import rng


__all__ = ['UniformDistribution']

class UniformDistribution(nn.Module):
    def __init__(self, size, scale=1.0):
        super().__init__()
        self.size = size
        self.scale = scale
        self.dim = int(np.prod(size))
        self.register_buffer('empty', torch.zeros(1))

    def forward(self, input, context=None):
        return self.log_prob(input, context)

    def log_prob(self, input, context=None):
        raise NotImplementedError('UniformDistribution.log_prob is not on the log-density forward path')

    def draw(self, n_samples, device=None):
        dev = self.empty.device if device is None else device
        x = rng.rand((n_samples, *self.size), dev, host_draw=True)      # uniform.py:32 (CPU draw in 'host' mode)
        return x if self.scale == 1.0 else x / self.scale

    def sample(self, n_samples, context=None):
        x = self.draw(n_samples)
        return x, torch.zeros(x.shape[0], device=x.device)

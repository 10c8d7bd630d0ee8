"""Normalization: y = x / scale + translation with a scalar scale (reference layers/normalize.py:6-49)."""
import torch
from torch import Tensor

from .. import ops
from .flowlayer import PreprocessingFlowLayer


def _as_vec(v):
    if isinstance(v, Tensor):
        return v
    if isinstance(v, (list, tuple)):
        return torch.Tensor(v)
    return torch.Tensor([v])


class Normalization(PreprocessingFlowLayer):
    def __init__(self, translation, scale, learnable=False):
        super().__init__()
        translation, scale = _as_vec(translation), _as_vec(scale)
        if learnable:
            self.translation = torch.nn.Parameter(translation)
            self.scale = torch.nn.Parameter(scale)
        else:
            self.register_buffer('translation', translation)
            self.register_buffer('scale', scale)
        self._host = None

    def host_constants(self):
        """(scale, translation, log(scale) in float32) as python floats; one device read, then cached."""
        key = (self.scale.data_ptr(), self.scale._version, self.translation._version)
        if self._host is None or self._host[0] != key:
            if self.scale.numel() != 1 or self.translation.numel() != 1:
                raise NotImplementedError('per-channel Normalization is not on the hot path (model.py:98-99 uses scalars)')
            s = self.scale.detach().float().cpu()
            self._host = (key, float(s), float(self.translation.detach().float().cpu()), float(torch.log(s).sum()))
        return self._host[1:]

    def forward(self, input, context=None):
        s, t, _ = self.host_constants()
        return ops.normalize(input, s, t), self.logdet(input, context)

    def reverse(self, input, context=None):
        s, t, _ = self.host_constants()
        return ops.normalize_inv(input, s, t)                        # (x - t) * s   (normalize.py:37-41)

    def logdet_value(self, C, D):
        """The per-sample ldj as a python float: float32 arithmetic in the reference's order, C * (-1 * D * log(scale))
        (normalize.py:42-47)."""
        _, _, logs = self.host_constants()
        return float(torch.tensor(C, dtype=torch.float32) * (torch.tensor(-1 * D, dtype=torch.float32) * torch.tensor(logs, dtype=torch.float32)))

    def logdet(self, input, context=None):
        B, C = input.shape[:2]
        D = input.numel() / B / C
        return torch.full((B,), self.logdet_value(C, D), device=input.device, dtype=torch.float32)

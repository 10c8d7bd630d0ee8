"""Categorical context embeddings (reference layers/rtdl/nn/_embeddings.py:76-283): each returns (embedding, context)."""
import math
from typing import List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from .... import _cabi, ops

__all__ = ['OneHotEncoder', 'EyeEncoder', 'CatEmbeddings']


def _init_embedding(weight: Tensor, d: int, init: str = 'uniform') -> None:
    """_initialize_embeddings of the reference: U(-1/sqrt(d), 1/sqrt(d)), zeros, or orthogonal."""
    if init == 'uniform':
        bound = 1 / math.sqrt(d)
        nn.init.uniform_(weight, a=-bound, b=bound)
    elif init == 'zeros':
        nn.init.zeros_(weight)
    elif init == 'orthogonal':
        nn.init.orthogonal_(weight)
    else:
        raise ValueError(f'unknown embedding init {init!r}')


class EyeEncoder(nn.Module):
    """Identity 'embedding': the integer context itself."""

    def forward(self, x: Tensor):
        if x.ndim != 2:
            raise ValueError('The input must have two dimensions')
        return x, x


class OneHotEncoder(nn.Module):
    def __init__(self, cardinalities: List[int]) -> None:
        super().__init__()
        self.register_buffer('cardinalities', torch.tensor(cardinalities))
        self._cards = [int(c) for c in cardinalities]

    def forward(self, x: Tensor):
        if x.ndim != 2:
            raise ValueError('The input must have two dimensions')
        d = _cabi.EncDesc()
        d.emb, d.type, d.n_ctx, d.C = _cabi.EMB['onehot'], _cabi.ENC['eyesample'], len(self._cards), sum(self._cards)
        for i, c in enumerate(self._cards):
            d.card[i] = c
        onehot, _ = ops.ctx_encode(x, None, d)
        return onehot.long(), x


class CatEmbeddings(nn.Module):
    def __init__(self, _cardinalities_and_maybe_dimensions: Union[List[int], List[Tuple[int, int]]],
                 d_embedding: Optional[int] = None, *, stack: bool = False, bias: bool = False, init: str = 'uniform') -> None:
        spec = _cardinalities_and_maybe_dimensions
        if not spec:
            raise ValueError('The first argument must be non-empty')
        pairs = isinstance(spec[0], tuple) and d_embedding is None
        plain = isinstance(spec[0], int) and d_embedding is not None
        if not (pairs or plain):
            raise ValueError('Invalid arguments: pass (cardinality, size) tuples, or cardinalities together with d_embedding')
        if stack and d_embedding is None:
            raise ValueError('stack can be True only when d_embedding is not None')
        if bias or stack:
            raise NotImplementedError('CatEmbeddings(bias=True / stack=True) is not used on the flow path (model.py:46,75-83)')
        super().__init__()
        spec_ = list(spec) if pairs else [(c, d_embedding) for c in spec]
        self._embeddings = nn.ModuleList(nn.Embedding(c, d) for c, d in spec_)
        self._biases = None
        self.stack, self.init = stack, init
        self.reset_parameters()

    def reset_parameters(self) -> None:
        for m in self._embeddings:
            _init_embedding(m.weight, m.weight.shape[-1], init=self.init)

    def get_embeddings(self, feature_idx: int) -> Tensor:
        if feature_idx < 0 or feature_idx >= len(self._embeddings):
            raise ValueError(f'feature_idx must be in the range(0, {len(self._embeddings)}). The provided value is {feature_idx}.')
        return self._embeddings[feature_idx].weight

    def tables(self):
        return [m.weight for m in self._embeddings]

    def uniform_width(self):
        w = {m.weight.shape[1] for m in self._embeddings}
        return w.pop() if len(w) == 1 else None

    def forward(self, x: Tensor):
        if x.ndim != 2:
            raise ValueError('x must have two dimensions')
        if x.shape[1] != len(self._embeddings):
            raise ValueError(f'x has {x.shape[1]} columns, but it must have {len(self._embeddings)} columns.')
        if self.uniform_width() is None:
            raise NotImplementedError('per-feature embedding widths differ: not produced by ContextEncoder')
        return ops.embed_lookup(x, self.tables()), x

"""The path's random draws (RNG contract).

The reference draws, in layer order: Dequantization noise with torch.rand on the CPU generator then .to(device)
(layers/distributions/uniform.py:32); Augment / encoder noise with torch.randn / torch.rand on the device
(layers/distributions/gaussian.py:69,265, layers/dequantize.py:57).  The layers here draw through this module in
the same order with the same shapes.  Tests replace the source with a replayed tape so that reference, oracle
and CUDA path consume identical numbers.
"""
from __future__ import annotations

import contextlib
import torch

_source = None            # object with rand(shape, device=, dtype=) / randn(...), or None for torch's generators
_dequant_on_host = False  # True reproduces uniform.py:32 literally (CPU generator + H2D copy every batch)


def set_dequant_mode(mode: str):
    """'device' (default): draw the image dequantisation noise with torch.rand on the GPU; 'host': on the CPU generator
    and copy, exactly as the reference does.  Same distribution, different generator stream."""
    global _dequant_on_host
    if mode not in ('device', 'host'):
        raise ValueError(mode)
    _dequant_on_host = mode == 'host'


@contextlib.contextmanager
def use_source(src):
    global _source
    prev, _source = _source, src
    try:
        yield src
    finally:
        _source = prev


def rand(shape, device, dtype=torch.float32, host_draw=False):
    if _source is not None:
        return _source.rand(tuple(shape), device=device, dtype=dtype)
    if host_draw and _dequant_on_host:
        return torch.rand(tuple(shape)).to(device=device, dtype=dtype)
    return torch.rand(tuple(shape), device=device, dtype=dtype)


def multinomial(probs, n):
    """n categorical draws (int64, on probs.device) with replacement."""
    if _source is not None and hasattr(_source, 'multinomial'):
        return _source.multinomial(probs, n)
    return torch.multinomial(probs, n, replacement=True)


def randn(shape, device, dtype=torch.float32):
    if _source is not None:
        return _source.randn(tuple(shape), device=device, dtype=dtype)
    return torch.randn(tuple(shape), device=device, dtype=dtype)

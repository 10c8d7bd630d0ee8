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
_recorder = None          # DrawRecorder: keeps every draw of a forward, in the reference's order (parity checks at large batches)
_dequant_on_host = False  # True reproduces uniform.py:32 literally (CPU generator + H2D copy every batch)
_capture_host_draws = []   # device buffers standing in for host draws inside the graph being captured (see rand)
_encoder_draws_batched = True   # False: every context encoder draws its own (B, width) block, i.e. the reference's generator stream


def set_reference_streams(flag: bool = True):
    """flag=True: consume torch's generators EXACTLY as the reference does -- image dequantisation noise from the CPU generator
    (uniform.py:32) and one device draw per context encoder in layer order (gaussian.py:265, dequantize.py:57) -- so that, under the
    same torch.manual_seed, this path and the reference's torch layers on the same GPU see identical noise (tests/test_gpu_boundary.py
    compares the unmodified experiment loops that way).  Default (False): dequantisation noise drawn on the device and all encoder
    noise of a forward cut from one draw -- same distributions, fewer launches, no per-batch host RNG.  Env: CFPP_RNG=reference."""
    global _dequant_on_host, _encoder_draws_batched
    _dequant_on_host, _encoder_draws_batched = bool(flag), not flag


def set_dequant_mode(mode: str):
    """'device' (default): draw the image dequantisation noise with torch.rand on the GPU; 'host': on the CPU generator
    and copy, exactly as the reference does.  Same distribution, different generator stream."""
    global _dequant_on_host
    if mode not in ('device', 'host'):
        raise ValueError(mode)
    _dequant_on_host = mode == 'host'


import os as _os
if _os.environ.get('CFPP_RNG') == 'reference':
    set_reference_streams(True)


@contextlib.contextmanager
def use_source(src):
    global _source
    prev, _source = _source, src
    try:
        yield src
    finally:
        _source = prev


class DrawRecorder:
    """Records the draws torch's own generators produce during a forward, as [(kind, tensor)] in the reference's draw order (SURVEY
    App. C-7), WITHOUT changing how they are drawn: the recorded forward consumes exactly the Philox stream an unrecorded one (or a
    CUDA-graph replay under the same seed) does.  `rows(idx)` replays a row subsample into the CPU oracle, which is how a batch of
    8192 or 131072 samples is compared with the reference arithmetic at the sizes bench.py measures."""

    def __init__(self):
        self.log = []
        self.paused = False

    def add(self, kind, t):
        if not self.paused:
            self.log.append((kind, t))

    def rows(self, idx):
        return _RowReplay(self.log, idx)


class _RowReplay:
    def __init__(self, log, idx):
        self.log, self.idx, self.pos = log, idx, 0

    def _next(self, kind, shape):
        k, t = self.log[self.pos]; self.pos += 1
        out = t.detach()[self.idx.to(t.device)].cpu()
        assert k == kind and tuple(out.shape) == tuple(shape), f'draw {self.pos - 1}: recorded {k}{tuple(out.shape)}, requested {kind}{tuple(shape)}'
        return out

    def rand(self, shape, device=None, dtype=None):
        return self._next('rand', shape)

    def randn(self, shape, device=None, dtype=None):
        return self._next('randn', shape)


@contextlib.contextmanager
def record_draws():
    global _recorder
    prev, _recorder = _recorder, DrawRecorder()
    try:
        yield _recorder
    finally:
        _recorder = prev


def rand(shape, device, dtype=torch.float32, host_draw=False):
    if _source is not None:
        return _source.rand(tuple(shape), device=device, dtype=dtype)
    if host_draw and _dequant_on_host:
        if torch.cuda.is_current_stream_capturing():
            # a CPU draw cannot live inside a CUDA graph: the capture records a persistent device buffer that GraphedLogProb refills
            # from the CPU generator before every replay (same stream of numbers as the reference's uniform.py:32)
            out = torch.empty(tuple(shape), device=device, dtype=dtype)
            _capture_host_draws.append(out)
            return out
        out = torch.rand(tuple(shape)).to(device=device, dtype=dtype)
    else:
        out = torch.rand(tuple(shape), device=device, dtype=dtype)
    if _recorder is not None:
        _recorder.add('rand', out)
    return out


def multinomial(probs, n):
    """n categorical draws (int64, on probs.device) with replacement."""
    if _source is not None and hasattr(_source, 'multinomial'):
        return _source.multinomial(probs, n)
    return torch.multinomial(probs, n, replacement=True)


def randn(shape, device, dtype=torch.float32):
    if _source is not None:
        return _source.randn(tuple(shape), device=device, dtype=dtype)
    out = torch.randn(tuple(shape), device=device, dtype=dtype)
    if _recorder is not None:
        _recorder.add('randn', out)
    return out

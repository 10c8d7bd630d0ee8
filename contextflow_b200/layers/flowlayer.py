"""The plugin boundary: FlowLayer.forward(x, context) -> (z, ldj) (reference layers/flowlayer.py:7-24)."""
from abc import ABCMeta, abstractmethod

import torch
import torch.nn as nn


class FlowLayer(nn.Module, metaclass=ABCMeta):
    @abstractmethod
    def forward(self, input, context=None):
        """-> (output, ldj)"""

    @abstractmethod
    def reverse(self, input, context=None):
        """-> input of forward"""

    @abstractmethod
    def logdet(self, input, context=None):
        """-> ldj"""


class ModifiedGradFlowLayer(FlowLayer):
    pass


class PreprocessingFlowLayer(FlowLayer):
    pass


def mark_expensive(func):
    func._expensive_computation = True
    return func


def inference_only(t: torch.Tensor):
    """Layers without backward kernels (context-conditioned layers, encoders, the ViT conditioner) refuse autograd loudly; the
    context-free conv stack trains through contextflow_b200/training.py."""
    if torch.is_grad_enabled() and t.requires_grad:
        raise NotImplementedError('this layer has no backward kernel yet (specialist / ViT layers): wrap the call in torch.no_grad(); '
                                  'context-free conv stacks train through contextflow_b200.training (DESIGN.md §8 f-1)')


LIVE_CAPTURE = False     # set by graphed.GraphedTrainStep while it captures a training step (see PackCache.get)


class live_capture:
    """While a TRAINING step is being captured into a CUDA graph, quantities derived from trainable parameters (log|det NN|, NN^-1,
    mixture tables, masked weights, repacked operands) must be recomputed by kernels recorded IN the graph: the optimizer rewrites the
    parameters between replays, and a cache hit at capture time would bake the warm-up's values into every replay."""

    def __enter__(self):
        global LIVE_CAPTURE
        self._prev, LIVE_CAPTURE = LIVE_CAPTURE, True

    def __exit__(self, *exc):
        global LIVE_CAPTURE
        LIVE_CAPTURE = self._prev


def derived_is_live(tensors):
    return LIVE_CAPTURE and any(getattr(t, 'requires_grad', False) for t in tensors)


class PackCache:
    """One-time weight repacking (K-major, padded) keyed on the source tensors' version counters and storage."""

    def __init__(self):
        self._store = {}

    def __deepcopy__(self, memo):                    # packed operands / ctypes descriptors belong to the tensors they were built from
        return PackCache()

    def get(self, name, tensors, build):
        if derived_is_live(tensors):                 # recorded in the graph, reads the live parameters on every replay; never cached
            with torch.no_grad():
                return build()
        key = tuple((t.data_ptr(), t._version, t.device) for t in tensors)
        hit = self._store.get(name)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, build())
            self._store[name] = hit
        return hit[1]

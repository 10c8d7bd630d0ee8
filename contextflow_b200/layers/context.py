"""Running a layer's `context_net(context) -> (c, logp_c)`.

create_model attaches to every specialist layer its own ContextEncoder = nn.Sequential(embedding, surjection)
(reference model.py:30-90, built from this package's classes).  ContextPlan recognises that shape and evaluates it as ONE
launch of the fused encoder kernel straight from the int64 context; any other callable is simply called."""
import torch.nn as nn


class ContextPlan:
    def __init__(self):
        self._fused = None
        self._owner = None
        self.preset = None          # (c, logp_c) computed ahead by an EncoderBatch launch of the enclosing FlowSequential

    def __deepcopy__(self, memo):                    # copy.deepcopy(model): the copy re-recognises its own encoder lazily
        return ContextPlan()

    def fused_for(self, context_net):
        from ._encoder_desc import FusedEncoder
        if self._owner is not context_net:
            self._owner = context_net
            self._fused = FusedEncoder.recognise(context_net)
        return self._fused

    def run(self, context_net, context):
        if self.preset is not None:
            out, self.preset = self.preset, None
            return out
        if self.fused_for(context_net) is not None:
            return self._fused(context)
        return context_net(context)

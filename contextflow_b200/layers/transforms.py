"""LogitTransform (reference layers/transforms.py:6-18)."""
from .. import ops
from .flowlayer import PreprocessingFlowLayer


class LogitTransform(PreprocessingFlowLayer):
    def forward(self, input, context=None):
        return ops.logit(input)

    def reverse(self, input, context=None):
        return ops.sigmoid(input)                                   # transforms.py:14-15

    def logdet(self, input, context=None):
        return ops.logit(input)[1]

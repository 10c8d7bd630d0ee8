from .gaussian import *   # noqa: F401,F403
from .uniform import *    # noqa: F401,F403


__all__ = ['StandardNormal', 'GaussianMixtureDistribution', 'ConditionalGaussianDistribution', 'UniformDistribution',
           'StudentMixtureDistribution']


class StudentMixtureDistribution:  # --dist tdist: outside the hot path
    def __init__(self, *a, **kw):
        raise NotImplementedError('StudentMixtureDistribution (--dist tdist) is outside the accelerated path')

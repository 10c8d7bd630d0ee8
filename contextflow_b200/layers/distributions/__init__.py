from .gaussian import *   # noqa: F401,F403
from .uniform import *    # noqa: F401,F403
from .student import *    # noqa: F401,F403


__all__ = ['StandardNormal', 'GaussianMixtureDistribution', 'ConditionalGaussianDistribution', 'UniformDistribution',
           'StudentMixtureDistribution']

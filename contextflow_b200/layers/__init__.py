"""Drop-in replacement for the reference's `layers` package (contextflow/layers/__init__.py:1-14): the same public class
names, constructor signatures, parameter / buffer names and shapes, with every forward evaluated by libcfpp (sm_100a)."""
from .dequantize import *            # noqa: F401,F403
from .normalize import Normalization  # noqa: F401
from .augment import Augment         # noqa: F401
from .distributions import *         # noqa: F401,F403
from .splitprior import SplitPrior   # noqa: F401
from .flowsequential import *        # noqa: F401,F403
from .conv1x1 import *               # noqa: F401,F403
from .permute_axes import PermuteAxes  # noqa: F401
from .activations import *           # noqa: F401,F403
from .actnorm import *               # noqa: F401,F403
from .squeeze import Squeeze, UnSqueeze  # noqa: F401
from .transforms import LogitTransform   # noqa: F401
from .coupling import *              # noqa: F401,F403
from .flowlayer import FlowLayer, PreprocessingFlowLayer  # noqa: F401
from .simple_vit import SimpleViT    # noqa: F401

"""algp_b200 -- B200-native GP inference and information-gain scoring with the
API of sumitsk/algp's models.py / utils.py / agent.py hot path.

Importing this package loads libalgp_b200.so (hand-written sm_100a CUDA behind
the C ABI of include/algp_b200.h) and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401  (loads the library or raises)
from .models import GPR, ExactGPModel  # noqa: F401
from .utils import CONST, entropy_from_cov, predictive_distribution, to_numpy, to_torch  # noqa: F401
from .agent import Agent, HotPath, patch  # noqa: F401
from . import tracing  # noqa: F401  (ALGP_TRACE=1 enables per-call NVTX ranges + CUDA-event timings)

__version__ = "0.1.0"

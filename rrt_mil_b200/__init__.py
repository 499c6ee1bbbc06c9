"""rrt_mil_b200 -- B200-native drop-in for DearCaat/RRT-MIL's ``RRTEncoder`` hot path.

``RRTEncoder`` keeps the reference's plug-in surface (constructor keywords, parameter names,
``forward(x:[1,N,D]) -> [1,N,D]``, ``final_dim``); the forward runs hand-written sm_100a CUDA kernels
through the C ABI in ``include/rrt_b200.h`` (``librrt_b200.so``).  There is no CPU / eager fallback.
"""
from .encoder import RRTEncoder, initialize_weights  # noqa: F401
from .mil import RRTMIL, DAttention  # noqa: F401
from . import cabi  # noqa: F401

__all__ = ["RRTEncoder", "RRTMIL", "DAttention", "initialize_weights", "cabi"]

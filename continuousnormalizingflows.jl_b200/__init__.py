"""continuousnormalizingflows.jl_b200 -- B200-native hot path of
ContinuousNormalizingFlows.jl behind the reference's own API names.

Importing this package loads ``libicnf_b200.so`` (hand-written sm_100a CUDA
kernels behind a C ABI).  It fails loudly if the library is missing: there is no
CPU or PyTorch fallback for any numeric result.
"""

from ._lib import ICNFError, LIB_PATH, lib  # noqa: F401  (loads the shared library)
from .api import (  # noqa: F401
    B200MatrixMode, Chain, ComputeMode, Dense, ICNF, MatrixMode, Mode, PlanarLayer, SolverStats, TestMode, TrainMode,
    augmented_f, base_sol, create_group, generate, group_info, group_join_id, group_set_global_norm, group_unique_id, inference, loss,
    loss_and_gradient, measure_fp32_peak, setup,
)
from .dist import CondICNFDist, ICNFDist  # noqa: F401
from .parallel import all_reduce_sum, dp_loss_and_gradient, group_join, shard_bounds  # noqa: F401
from .mlj import Adam, CondICNFModel, ICNFModel, WeightDecay, make_opt_callback  # noqa: F401

__all__ = [
    "ICNF", "inference", "generate", "loss", "loss_and_gradient", "setup", "augmented_f", "base_sol",
    "TestMode", "TrainMode", "B200MatrixMode", "Dense", "PlanarLayer", "Chain", "ICNFDist", "CondICNFDist",
    "ICNFModel", "CondICNFModel", "ICNFError", "SolverStats",
]

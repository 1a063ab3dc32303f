"""rlap_b200 — B200-native randomized Schur-complement graph augmentor (rLap).

Drop-in for the hot path of kvignesh1420/rlap: `rlap_b200.ops.approximate_cholesky` has the
signature and output contract of `rlap.ops.approximate_cholesky` (rlap/ops.py:7-58).
"""
from . import ops  # noqa: F401
from .ops import (Graph, approximate_cholesky, approximate_cholesky_batched, identity, manual_seed, prepare,  # noqa: F401
                  schur_views)

VERSION = "0.1.0"


def register_torch_op():
    """register torch.ops.extension_cpp.approximate_cholesky / identity with the reference's schema"""
    from . import torch_op
    return torch_op.register()

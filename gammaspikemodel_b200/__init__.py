"""gammaspikemodel_b200 — B200 (sm_100a) implementation of GammaSpikeModel's per-branch compute
(clock-rate map, spike / rate prior log-densities, stumped-tree-prior tree and stub terms) behind the
C ABI of include/gsm_b200.h.  The CUDA library is required: there is no CPU path."""
from . import _abi as abi  # noqa: F401
from .engine import GsmEngine, GsmError  # noqa: F401

__all__ = ["abi", "GsmEngine", "GsmError"]

"""diral_b200 -- Blackwell-native vectorised V2V resource-allocation environment.

The per-time-slot body of the DIRAL "test simulator" (reference envs/test_env.py, network.py,
vehicle.py) as hand-written sm_100a CUDA kernels over a batch of independent environments, behind
the reference's own ``TestEnv`` surface.  See DESIGN.md and include/diral_env.h.
"""
from .env import BatchedTestEnv, METRIC_FIELDS, METRIC_LEN, Network, TestEnv, cfg_from_kwargs  # noqa: F401
from ._lib import DiralError  # noqa: F401

__all__ = ["TestEnv", "BatchedTestEnv", "Network", "DiralError", "cfg_from_kwargs", "METRIC_LEN", "METRIC_FIELDS"]

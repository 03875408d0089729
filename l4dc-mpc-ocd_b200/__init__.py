"""B200-native batched MPC engine for the planner hot path of avikj/L4DC-MPC-OCD.

``engine``          torch-tensor front-end of the C ABI (libocd_b200.so; CUDA only, no CPU fallback)
``interact_drive``  drop-in mirror of the reference's planner / car / world / MPC_ORD interface
``experiments``     the reference's scenario constructors and the run_mpc_ord driver
"""
from . import _native  # noqa: F401  (raises ImportError when libocd_b200.so has not been built)
from .engine import (  # noqa: F401
    Engine, HostContext, PlannerParams, Scenario, MATH_FAST, MATH_PRECISE, device_count,
)

__all__ = ["Engine", "HostContext", "PlannerParams", "Scenario", "MATH_FAST", "MATH_PRECISE", "device_count"]

"""rsruckig_b200 — B200-native batched Ruckig (state-to-state, jerk-limited trajectory calculation).

The package holds the CUDA kernels + C ABI (csrc/, include/ruckig_b200.h) and a thin host-side mirror of the
reference's public API (api.py). Nothing here computes on the CPU.
"""
from .api import (BatchRollout, BatchRuckig, MultiBatchRuckig, CalculatorError, ControlInterface, ControlSigns, Direction, DurationDiscretization,
                  IgnoreErrorHandler, InputParameter, OutputParameter, Profile, ReachedLimits, Ruckig, RuckigError,
                  RuckigResult, Synchronization, ThrowErrorHandler, Trajectory, ValidationError)

__all__ = ["BatchRollout", "BatchRuckig", "MultiBatchRuckig", "CalculatorError", "ControlInterface", "ControlSigns", "Direction", "DurationDiscretization",
           "IgnoreErrorHandler", "InputParameter", "OutputParameter", "Profile", "ReachedLimits", "Ruckig", "RuckigError",
           "RuckigResult", "Synchronization", "ThrowErrorHandler", "Trajectory", "ValidationError"]

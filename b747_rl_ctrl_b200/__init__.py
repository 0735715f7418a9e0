"""b747_rl_ctrl_b200 -- B200-native batched B747 pitch-control environment.

Hot path (CUDA, sm_100a): csrc/ behind the C ABI in include/b747.h.
Host-side mirrors of the reference interface: core.model.Model, core.controller.Controller,
env.ctrl_env.ControllerEnv, vec_env.B747VecEnv.
"""
from ._lib import B747Error, F32, F64  # noqa: F401

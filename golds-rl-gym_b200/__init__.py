"""golds-rl-gym_b200 -- B200-native (sm_100a) swarm-environment hot path of allentran/golds-rl-gym.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI),
the ctypes binding and the host-side mirror of the reference's interfaces:

  envs.multiagent          SwarmEnv / BatchedSwarmEnv          (fed_gym/envs/multiagent.py)
  agents.state_processors  SwarmStateProcessor                 (fed_gym/agents/state_processors.py)
  agents.paac              SwarmRunner statics, GridRunners    (fed_gym/agents/paac/{emulator_runner,runners}.py)

The directory name carries a hyphen (the reference repo's name); import it with
``importlib.import_module("golds-rl-gym_b200")`` or through the root-level alias module
``golds_rl_gym_b200``.
"""
from . import _native
from ._native import SwarmNativeError, build, build_torch_ext, load, load_torch_ops

__all__ = ["_native", "SwarmNativeError", "build", "build_torch_ext", "load", "load_torch_ops"]

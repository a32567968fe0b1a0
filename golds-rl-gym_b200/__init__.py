"""golds-rl-gym_b200 -- B200-native (sm_100a) swarm-environment hot path of allentran/golds-rl-gym.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI),
the ctypes binding and the host-side mirror of the reference's interfaces:

  envs.multiagent                 SwarmEnv / BatchedSwarmEnv / make    (fed_gym/envs/multiagent.py, fed_gym/__init__.py)
  agents.state_processors         SwarmStateProcessor                  (fed_gym/agents/state_processors.py)
  agents.paac.emulator_runner     SwarmRunner statics                  (fed_gym/agents/paac/emulator_runner.py)
  agents.paac.runners             GridRunners (six shared variables)   (fed_gym/agents/paac/runners.py)
  agents.paac.paac                GridPAACLearner (device-resident)    (fed_gym/agents/paac/paac.py, actor_learner.py)
  agents.paac.policy_v_network    ConvSingleAgentPolicyNetwork (torch) (fed_gym/agents/paac/policy_v_network.py)
  agents.paac.policy_monitor      SwarmPolicyMonitor                   (fed_gym/agents/paac/policy_monitor.py)
  agents.paac.environment_creator SwarmEnvironmentCreator              (fed_gym/agents/paac/environment_creator.py)
  sharding                        env-batch sharding over GPUs         (fed_gym/agents/paac/runners.py:18-19)

The directory name carries a hyphen (the reference repo's name); import it with
``importlib.import_module("golds-rl-gym_b200")`` or through the root-level alias module
``golds_rl_gym_b200``.
"""
from . import _native
from ._native import SwarmNativeError, build, build_torch_ext, load, load_torch_ops

__all__ = ["_native", "SwarmNativeError", "build", "build_torch_ext", "load", "load_torch_ops"]

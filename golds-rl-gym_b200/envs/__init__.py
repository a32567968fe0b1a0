from .multiagent import BatchedSwarmEnv, InjectedDraws, SwarmEnv, make  # noqa: F401

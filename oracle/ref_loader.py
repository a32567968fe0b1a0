"""Import the *real* reference swarm env / rasteriser by file path.

TEST INFRASTRUCTURE.  ``/root/reference`` is mounted read-only in the build
container and does not exist on the GPU box; there the byte-identical copies
staged by ``oracle/make_ref.py`` under ``oracle/_ref/`` (git-ignored, shipped like
a built .so) are loaded instead.  Everything that calls :func:`load_reference`
must still be skippable (``reference_available()``).

The reference only needs ``gym.Env`` as a base class whose ``step/reset``
delegate to ``_step/_reset`` (gym==0.9.4 semantics, requirements.txt:6), so a
6-line stub stands in for gym.  ``fed_gym/__init__.py`` is never executed
(it would import the gym registry, pandas and the finance envs).

Reference entry points loaded:
  fed_gym/envs/multiagent.py:7-115            SwarmEnv
  fed_gym/agents/state_processors.py:15-42    SwarmStateProcessor
  fed_gym/agents/paac/emulator_runner.py:98-118  SwarmRunner.get_local_states /
                                              transform_actions_for_env
"""
import importlib
import importlib.util
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("SWARM_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/fed_gym/envs/multiagent.py") else _STAGED)

_cache = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fed_gym", "envs", "multiagent.py"))


def _gym_stub():
    gym = types.ModuleType("gym")

    class Env(object):
        def step(self, action):
            return self._step(action)

        def reset(self):
            return self._reset()

    gym.Env = Env
    return gym


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (multiagent_module, state_processors_module)."""
    if "core" in _cache:
        return _cache["core"]
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    saved = sys.modules.get("gym")
    sys.modules["gym"] = _gym_stub()
    try:
        ma = _load("_ref_multiagent", "fed_gym/envs/multiagent.py")
    finally:
        if saved is None:
            sys.modules.pop("gym", None)
        else:
            sys.modules["gym"] = saved
    sp = _load("_ref_state_processors", "fed_gym/agents/state_processors.py")
    _cache["core"] = (ma, sp)
    return ma, sp


def load_reference_runner():
    """Returns the reference ``SwarmRunner`` class (static helpers only are usable:
    its module imports tensorflow at top level, which is stubbed out here)."""
    if "runner" in _cache:
        return _cache["runner"]
    ma, sp = load_reference()
    names = ["tensorflow", "fed_gym", "fed_gym.agents", "fed_gym.agents.paac",
             "fed_gym.agents.a3c", "fed_gym.agents.a3c.worker", "fed_gym.agents.state_processors"]
    saved = {n: sys.modules.get(n) for n in names}
    try:
        for n in names:
            sys.modules[n] = types.ModuleType(n)
        sys.modules["fed_gym"].__path__ = [os.path.join(REFERENCE_ROOT, "fed_gym")]
        sys.modules["fed_gym.agents"].__path__ = [os.path.join(REFERENCE_ROOT, "fed_gym", "agents")]
        sys.modules["fed_gym.agents.paac"].__path__ = [os.path.join(REFERENCE_ROOT, "fed_gym", "agents", "paac")]
        sys.modules["fed_gym.agents.a3c"].__path__ = []
        sys.modules["fed_gym.agents.a3c.worker"].sigmoid = lambda x: x
        sys.modules["fed_gym.agents.state_processors"] = sp     # the real module, loaded by path
        mod = importlib.import_module("fed_gym.agents.paac.emulator_runner")
        runner = mod.SwarmRunner
    finally:
        for n in names + ["fed_gym.agents.paac.emulator_runner"]:
            sys.modules.pop(n, None)
        for n, m in saved.items():
            if m is not None:
                sys.modules[n] = m
    _cache["runner"] = runner
    return runner


def make_reference_env(n_locusts=80, seed=None):
    """A fresh reference SwarmEnv with the swarm size set the way the reference allows
    (class attribute, tests/env_tests.py mutate it the same way)."""
    ma, _ = load_reference()
    cls = type("SwarmEnvN%d" % n_locusts, (ma.SwarmEnv,), {"N_LOCUSTS": n_locusts})
    return cls(seed=seed)


def reference_reset_injected(env, x0, xa0, burn_actions, agent_noise, particle_noise):
    """Run the reference's reset with the random draws injected instead of drawn.

    Equivalent to multiagent.py:46-63 with lines 51-56 replaced by assignments.
    ``agent_noise``/``particle_noise`` need rows 0..N_BURN_IN (row 10 is the frozen
    post-reset row); they are padded to the reference's 138 rows.
    """
    import numpy as np
    env.t = 0
    nb = env.N_BURN_IN
    rows = 128 + nb

    def pad(tbl):
        out = np.zeros((rows,) + tbl.shape[1:], dtype=np.float64)
        out[: tbl.shape[0]] = tbl
        return out

    env.agent_noise = pad(np.asarray(agent_noise, dtype=np.float64))
    env.particle_noise = pad(np.asarray(particle_noise, dtype=np.float64))
    env.states = [np.array(x0, dtype=np.float64), np.array(xa0, dtype=np.float64)]
    for ii in range(nb):
        env.step(np.asarray(burn_actions[ii], dtype=np.float64))
        env.t += 1
    return env.states

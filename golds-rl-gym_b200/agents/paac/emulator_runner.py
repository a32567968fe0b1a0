"""Host-side mirror of fed_gym/agents/paac/emulator_runner.py:82-118 (SwarmRunner statics).

The per-worker process loop (_run, :120-151) has no counterpart: the whole env batch is stepped by
one kernel launch (see runners.GridRunners).  The static helpers keep their names and meaning but
act on device tensors for the whole batch.
"""
import ctypes

import torch

from ... import _native as nat


class SwarmRunner(object):
    STATE_IDX = 0
    HISTORY_IDX = 1
    AGENT_POSITIONS_IDX = 2
    REWARD_IDX = 3
    DONE_IDX = 4
    ACTIONS_IDX = 5

    MAX_MOVE_NORM = 1

    @staticmethod
    def get_local_states(state, agent_positions, out=None):
        """emulator_runner.py:98-111 for a batch: state (E,G,G,2) f32 + agent_positions (E,A,2) u8 ->
        (E,A,G,G,3) f32 (grid channels + one-hot at the agent's cell).  A single env may be passed
        as (G,G,2)/(A,2) and comes back as (A,G,G,3)."""
        lib = nat.load()
        single = state.dim() == 3
        if single:
            state, agent_positions = state[None], agent_positions[None]
        state = state.contiguous()
        agent_positions = agent_positions.contiguous()
        E, G = state.shape[0], state.shape[1]
        A = agent_positions.shape[1]
        if out is None:
            out = torch.empty(E, A, G, G, 3, dtype=torch.float32, device=state.device)
        p = nat.SwarmParams(n_envs=E, n_locusts=1, n_agents=A, grid_size=G, n_burn_in=0, max_episode_steps=0,
                            math_mode=0, tuning=0, noise=0, gravity=0, wind=0, F=0, L=1, dt=0, box_width=3.0,
                            box_height=3.0, seed=0, env_id_offset=0)
        with torch.cuda.device(state.device):      # the C side launches on the CURRENT device
            nat.check(lib.swarm_expand_obs(ctypes.byref(p), ctypes.c_void_p(state.data_ptr()),
                                           ctypes.c_void_p(agent_positions.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                           ctypes.c_void_p(torch.cuda.current_stream(state.device).cuda_stream)),
                      "swarm_expand_obs")
        return out[0] if single else out

    @staticmethod
    def transform_actions_for_env(actions):
        """emulator_runner.py:113-118: rows with |a| >= MAX_MOVE_NORM are normalised IN PLACE; the same
        tensor is returned (the learner stores the clipped actions, paac.py:309,316)."""
        lib = nat.load()
        if actions.dtype != torch.float32 or not actions.is_contiguous() or actions.shape[-1] != 2:
            raise ValueError("actions must be a contiguous float32 (...,2) CUDA tensor")
        with torch.cuda.device(actions.device):    # the C side launches on the CURRENT device
            nat.check(lib.swarm_clip_actions(ctypes.c_void_p(actions.data_ptr()), actions.numel() // 2,
                                             float(SwarmRunner.MAX_MOVE_NORM),
                                             ctypes.c_void_p(torch.cuda.current_stream(actions.device).cuda_stream)),
                      "swarm_clip_actions")
        return actions

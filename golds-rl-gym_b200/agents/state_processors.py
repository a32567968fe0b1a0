"""Host-side mirror of fed_gym/agents/state_processors.py:15-42 (SwarmStateProcessor).

process_state runs the rasteriser kernel (swarm_rasterize in include/swarm_b200.h) on the
given state; numpy in / numpy out like the reference.  The batched, device-resident form is
``BatchedSwarmEnv.observe()`` / the fused rasterise inside ``BatchedSwarmEnv.step``.
"""
import ctypes

import numpy as np
import torch

from .. import _native as nat


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class SwarmStateProcessor(object):
    def __init__(self, scales=1., grid_size=20):
        self.scales = scales
        self.grid_size = grid_size
        self.positions = None
        self.WIDTH = 3.
        self.HEIGHT = 3.
        self._lib = nat.load()

    def _params(self, n_points, n_agents):
        return nat.SwarmParams(n_envs=1, n_locusts=n_points, n_agents=n_agents, grid_size=self.grid_size,
                               n_burn_in=10, max_episode_steps=0, math_mode=0, tuning=0, noise=1e-4,
                               gravity=-1.0, wind=1.0, F=0.5, L=10.0, dt=0.05, box_width=self.WIDTH,
                               box_height=self.HEIGHT, seed=0, env_id_offset=0)

    def _run(self, x, xa):
        dev = torch.device("cuda", torch.cuda.current_device())
        x_t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        A = 0 if xa is None else int(np.shape(xa)[0])
        xa_t = torch.as_tensor(np.ascontiguousarray(xa, dtype=np.float64)).to(dev) if A else None
        G = self.grid_size
        grid = torch.empty(G, G, 2, dtype=torch.float32, device=dev)
        pos = torch.empty(max(A, 1), 2, dtype=torch.uint8, device=dev)
        box = torch.empty(4, dtype=torch.float64, device=dev)
        p = self._params(x_t.shape[0], A)
        nat.check(self._lib.swarm_rasterize(ctypes.byref(p), _ptr(x_t), _ptr(xa_t), _ptr(grid), _ptr(pos),
                                            _ptr(box), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                  "swarm_rasterize")
        return grid, pos[:A], box

    def _get_bounding_box(self, x):
        """state_processors.py:25-27: [[mean_x - W/2, mean_x + W/2], [0, 2H]] of the given points."""
        _, _, box = self._run(x, None)
        b = box.cpu().numpy()
        return [[b[0], b[1]], [0, 2 * self.HEIGHT]]

    def process_state(self, state):
        """state_processors.py:29-42: (G,G,2) occupancy grid; side effect: self.positions (A,2) uint8."""
        grid, pos, _ = self._run(state[0], state[1])
        self.positions = pos.cpu().numpy()
        return grid.double().cpu().numpy()

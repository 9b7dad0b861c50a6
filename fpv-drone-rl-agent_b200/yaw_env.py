"""Host-side mirror of the reference's yaw task (simulation/yaw.py ``DroneEnv``) on libquadx_b200.so.

yaw.py cannot be imported in the reference itself (``from hover import DroneEnv`` -- hover.py defines only
``QuadXHoverEnv``, yaw.py:8) and calls four methods that exist nowhere (sphere detector, angular velocity,
add_sphere, reward).  What it does write down -- spaces, action scaling, one Aviary.step per env step, the 4-deep
action history, truncation / termination, the observation order -- is implemented as written; the missing pieces are
declared stand-ins."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import QX_TASK_YAW, QxConfig, default_config
from .hover_env import Box, QuadXSim


class QuadXYawVecEnv:
    """Batched yaw env with the VecEnv call shape of QuadXHoverVecEnv: obs f32 [N,12], actions f32 [N,1]."""

    def __init__(self, num_envs: int, seed: int = 0, device=None, cfg: QxConfig | None = None, env_id0: int = 0, **cfg_overrides):
        cfg = cfg if cfg is not None else default_config(QX_TASK_YAW)
        cfg.update(auto_reset=1, **cfg_overrides)
        self.sim = QuadXSim(num_envs, cfg, seed=seed, env_id0=env_id0, device=device, task=QX_TASK_YAW)
        self.num_envs, self.device = num_envs, self.sim.device
        self.action_space = Box(-np.ones(1), np.ones(1), np.float32)  # yaw.py:37
        self.observation_space = Box(-np.ones(12), np.ones(12), np.float32)  # yaw.py:41-45
        n, d = num_envs, self.device
        self.obs = torch.zeros(n, 12, device=d)
        self.rewards = torch.zeros(n, device=d)
        self.terminated = torch.zeros(n, dtype=torch.uint8, device=d)
        self.truncated = torch.zeros(n, dtype=torch.uint8, device=d)
        self.terminal_obs = torch.zeros(n, 12, device=d)

    def reset(self) -> torch.Tensor:
        self.sim.reset(self.obs)
        return self.obs

    def step(self, actions):
        a = torch.as_tensor(actions, dtype=torch.float32, device=self.device).reshape(self.num_envs, 1).contiguous()
        self.sim.step(a, self.obs, self.rewards, self.terminated, self.truncated, self.terminal_obs)
        return self.obs, self.rewards, (self.terminated | self.truncated).bool(), None

    def close(self) -> None:
        self.sim.close()


class DroneEnv:
    """Single-env facade with the gymnasium signatures of yaw.py:32-149."""

    def __init__(self, render: bool = False, seed: int = 0, device=None, **cfg_overrides):
        cfg = default_config(QX_TASK_YAW)
        cfg.update(auto_reset=0, render=int(render), **cfg_overrides)
        self.sim = QuadXSim(1, cfg, seed=seed, device=device, task=QX_TASK_YAW)
        self.action_space = Box(-np.ones(1), np.ones(1), np.float32)
        self.observation_space = Box(-np.ones(12), np.ones(12), np.float32)
        self.hardcoded_roll, self.hardcoded_pitch, self.hardcoded_throttle = 0.0, 0.0, -1.0  # yaw.py:48-50
        self.info: dict = {}

    def reset(self, seed=None, options=None):
        return self.sim.reset_host()[0], {}  # yaw.py:99-100

    def step(self, action):
        obs, rew, te, tr, _ = self.sim.step_host(np.asarray(action, np.float32).reshape(1, 1))
        self.info = {"out_of_bounds": True} if te[0] else {}  # yaw.py:135,141-143
        return obs[0], float(rew[0]), bool(te[0]), bool(tr[0]), self.info

    def close(self):
        self.sim.close()
